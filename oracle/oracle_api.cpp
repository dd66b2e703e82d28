// TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// C driver around the restated AD3 solver. Mirrors, per sample, what
// lvmhelpers/sparsemap.py:13-44 does through the Cython factor wrappers:
//   eta = [x (double), 0_n]; optional init_active_set_from_scores(init);
//   solve_qp(eta, t, max_iter); (aset, p) = get_sparse_solution();
//   backward: dist_jacobian_vec(dp) -> d_eta_u[:n] (and d_eta_v).
//
// Built twice from this one file:
//   * default            -> oracle/_build/liboracle.so, our restated factors;
//   * -DSMLVM_REF_FACTORS -> oracle/_ref/libref_factors.so, the reference's
//     own bernoulli.h / budget.h / sequence.h / sequence_binary.h, included
//     from /root/reference/lvmhelpers where they lie (oracle/ref/Makefile).
#include <cstdint>
#include <cstring>
#include <memory>

#ifdef SMLVM_REF_FACTORS
#include "bernoulli.h"
#include "budget.h"
#include "sequence.h"
#include "sequence_binary.h"
namespace AD3 {
typedef FactorBernoulli FBern;
typedef FactorBudget FBudget;
typedef FactorSequenceBinary FSeqBin;
typedef FactorSequence FSeq;
}
#else
#include "factors_restated.h"
namespace AD3 {
typedef RBernoulli FBern;
typedef RBudget FBudget;
typedef RSequenceBinary FSeqBin;
typedef RSequence FSeq;
}
#endif

using namespace AD3;

enum FactorKind { KIND_BERNOULLI = 0, KIND_BUDGET = 1, KIND_SEQUENCE_BINARY = 2 };

static GenericFactor* make_factor(int kind, int n, int budget)
{
    if (kind == KIND_BERNOULLI) { auto* f = new FBern(); f->Initialize(n); return f; }
    if (kind == KIND_BUDGET) { auto* f = new FBudget(); f->Initialize(n, budget); return f; }
    if (kind == KIND_SEQUENCE_BINARY) { auto* f = new FSeqBin(); f->Initialize(n); return f; }
    return nullptr;
}

extern "C" {

// One sample, forward.
//   x[n]            on-scores (off-scores are 0, sparsemap.py:22)
//   t[n-1] or NULL  1->1 transition scores (sequence-binary only)
//   init[2n] or NULL  scores for init_active_set_from_scores (sparsemap.py:26)
// Outputs (caller-allocated, capacity `cap` vertices):
//   p[cap] (double), aset[cap*n] (int32 0/1), S[cap*cap] (inverse_A[1:,1:],
//   row stride = *count), *count = |A|, *iters, *status (AD3::SolveStatus),
//   u[2n] = marginals M p of the returned distribution.
// Returns 0, or -1 on bad arguments / capacity overflow.
int smlvm_oracle_sparsemap_fwd(int kind, int n, int budget, const double* x,
                               const double* t, const double* init,
                               int max_iter, int cap, double* p, int32_t* aset,
                               double* S, int* count, int* iters, int* status,
                               double* u)
{
    std::unique_ptr<GenericFactor> f(make_factor(kind, n, budget));
    if (!f || n < 1 || max_iter < 0) return -1;
    f->SetVerbosity(1);
    f->SetQPMaxIter(max_iter);
    vector<double> eta(2 * n, 0.0);
    for (int i = 0; i < n; ++i) eta[i] = x[i];
    vector<double> add;
    if (kind == KIND_SEQUENCE_BINARY) {
        add.assign(n - 1, 0.0);
        if (t) for (int i = 0; i + 1 < n; ++i) add[i] = t[i];
    }
    if (init) {
        vector<double> init_u(init, init + 2 * n);
        vector<double> init_v(add.size(), 0.0);
        f->InitActiveSetFromScores(init_u, init_v);
    }
    vector<double> post_u, post_v;
    f->SolveQP(eta, add, &post_u, &post_v);

    const vector<Configuration>& A = f->GetQPActiveSet();
    const vector<double>& dist = f->GetQPDistribution();
    const vector<double>& invA = f->GetQPInvA();
    int m = static_cast<int>(A.size());
    *count = m;
    *iters = f->last_iterations();
    *status = f->last_status();
    if (m > cap) return -1;
    for (int j = 0; j < m; ++j) {
        p[j] = dist[j];
        const vector<int>& y = *static_cast<const vector<int>*>(A[j]);
        for (int i = 0; i < n; ++i) aset[j * n + i] = y[i];
    }
    if (S) {
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j)
                S[i * m + j] = invA[(i + 1) * (m + 1) + (j + 1)];
    }
    if (u) {
        for (int i = 0; i < 2 * n; ++i) u[i] = 0.0;
        for (int j = 0; j < m; ++j) {
            const vector<int>& y = *static_cast<const vector<int>*>(A[j]);
            for (int i = 0; i < n; ++i) u[y[i] == 1 ? i : n + i] += dist[j];
        }
    }
    return 0;
}

// One sample, forward + backward in one go (the reference keeps `ctx.f` alive
// between the two, sparsemap.py:19,42; here we re-solve deterministically).
//   dp[|A|] -> dx[n] = d_eta_u[:n], dt[n-1] = d_eta_v (sequence only).
int smlvm_oracle_sparsemap_fwd_bwd(int kind, int n, int budget, const double* x,
                                   const double* t, const double* init,
                                   int max_iter, const double* dp, int dp_len,
                                   double* dx, double* dt)
{
    std::unique_ptr<GenericFactor> f(make_factor(kind, n, budget));
    if (!f || n < 1) return -1;
    f->SetQPMaxIter(max_iter);
    vector<double> eta(2 * n, 0.0);
    for (int i = 0; i < n; ++i) eta[i] = x[i];
    vector<double> add;
    if (kind == KIND_SEQUENCE_BINARY) {
        add.assign(n - 1, 0.0);
        if (t) for (int i = 0; i + 1 < n; ++i) add[i] = t[i];
    }
    if (init) {
        vector<double> init_u(init, init + 2 * n);
        vector<double> init_v(add.size(), 0.0);
        f->InitActiveSetFromScores(init_u, init_v);
    }
    vector<double> post_u, post_v;
    f->SolveQP(eta, add, &post_u, &post_v);
    int m = static_cast<int>(f->GetQPActiveSet().size());
    if (dp_len != m) return -2;
    vector<double> v(dp, dp + m);
    vector<double> out_u(2 * n, 0.0), out_v(add.size(), 0.0);
    f->DistJacobianVec(v, &out_u, &out_v);
    for (int i = 0; i < n; ++i) dx[i] = out_u[i];
    if (dt) for (size_t i = 0; i < out_v.size(); ++i) dt[i] = out_v[i];
    return 0;
}

// Factor MAP oracle alone (a-9): eta[2n] (+ t[n-1]) -> y[n], value.
int smlvm_oracle_factor_map(int kind, int n, int budget, const double* eta2n,
                            const double* t, int32_t* y, double* value)
{
    std::unique_ptr<GenericFactor> f(make_factor(kind, n, budget));
    if (!f) return -1;
    vector<double> eta(eta2n, eta2n + 2 * n);
    vector<double> add;
    if (kind == KIND_SEQUENCE_BINARY) {
        add.assign(n - 1, 0.0);
        if (t) for (int i = 0; i + 1 < n; ++i) add[i] = t[i];
    }
    Configuration cfg = f->CreateConfiguration();
    f->Maximize(eta, add, cfg, value);
    const vector<int>& yy = *static_cast<const vector<int>*>(cfg);
    for (int i = 0; i < n; ++i) y[i] = yy[i];
    f->DeleteConfiguration(cfg);
    return 0;
}

// ---- general linear chain (lvmhelpers/sequence.h, bound as PFactorSequence.initialize(list[int])) ----
// L positions with num_states[i] states; x[V] node scores (V = sum num_states), add[E] edge scores in the order of
// FactorSequence::Initialize (sequence.h:264-291).  Vertices come back as state sequences, aset[cap * L].
static FSeq* make_seq(int L, const int32_t* num_states)
{
    auto* f = new FSeq();
    f->Initialize(vector<int>(num_states, num_states + L));
    return f;
}

int smlvm_oracle_seq_fwd(int L, const int32_t* num_states, const double* x, const double* add_in, int max_iter, int cap,
                         double* p, int32_t* aset, double* S, int* count, int* iters, int* status)
{
    std::unique_ptr<FSeq> f(make_seq(L, num_states));
    int V = 0;
    for (int i = 0; i < L; ++i) V += num_states[i];
    f->SetVerbosity(1);
    f->SetQPMaxIter(max_iter);
    vector<double> eta(x, x + V), add(add_in, add_in + f->GetNumAdditionals());
    vector<double> post_u, post_v;
    f->SolveQP(eta, add, &post_u, &post_v);
    const vector<Configuration>& A = f->GetQPActiveSet();
    const vector<double>& dist = f->GetQPDistribution();
    const vector<double>& invA = f->GetQPInvA();
    const int m = static_cast<int>(A.size());
    *count = m;
    *iters = f->last_iterations();
    *status = f->last_status();
    if (m > cap) return -1;
    for (int j = 0; j < m; ++j) {
        p[j] = dist[j];
        const vector<int>& y = *static_cast<const vector<int>*>(A[j]);
        for (int i = 0; i < L; ++i) aset[j * L + i] = y[i];
    }
    if (S)
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) S[i * m + j] = invA[(i + 1) * (m + 1) + (j + 1)];
    return 0;
}

int smlvm_oracle_seq_fwd_bwd(int L, const int32_t* num_states, const double* x, const double* add_in, int max_iter,
                             const double* dp, int dp_len, double* dx, double* dt)
{
    std::unique_ptr<FSeq> f(make_seq(L, num_states));
    int V = 0;
    for (int i = 0; i < L; ++i) V += num_states[i];
    f->SetQPMaxIter(max_iter);
    vector<double> eta(x, x + V), add(add_in, add_in + f->GetNumAdditionals());
    vector<double> post_u, post_v;
    f->SolveQP(eta, add, &post_u, &post_v);
    const int m = static_cast<int>(f->GetQPActiveSet().size());
    if (dp_len != m) return -2;
    vector<double> v(dp, dp + m), out_u(V, 0.0), out_v(add.size(), 0.0);
    f->DistJacobianVec(v, &out_u, &out_v);
    for (int i = 0; i < V; ++i) dx[i] = out_u[i];
    for (size_t i = 0; i < out_v.size(); ++i) dt[i] = out_v[i];
    return 0;
}

int smlvm_oracle_seq_map(int L, const int32_t* num_states, const double* x, const double* add_in, int32_t* y,
                         double* value)
{
    std::unique_ptr<FSeq> f(make_seq(L, num_states));
    int V = 0;
    for (int i = 0; i < L; ++i) V += num_states[i];
    vector<double> eta(x, x + V), add(add_in, add_in + f->GetNumAdditionals());
    Configuration cfg = f->CreateConfiguration();
    f->Maximize(eta, add, cfg, value);
    const vector<int>& yy = *static_cast<const vector<int>*>(cfg);
    for (int i = 0; i < L; ++i) y[i] = yy[i];
    f->DeleteConfiguration(cfg);
    return 0;
}

// Batched forward in a tight C++ loop (CPU-baseline timing leg: separates the
// solver cost from the reference's per-sample Python loop, structmarg.py:181).
// Writes only support sizes and iteration counts; returns sum of |A_b|.
long smlvm_oracle_sparsemap_batch(int kind, int n, int budget, const float* x,
                                  long rows, int max_iter, int* counts,
                                  int* iters)
{
    long total = 0;
    vector<double> xd(n);
    int cap = n + 2;
    vector<double> p(cap);
    vector<int32_t> aset(static_cast<size_t>(cap) * n);
    for (long b = 0; b < rows; ++b) {
        for (int i = 0; i < n; ++i) xd[i] = x[b * n + i];
        int c = 0, it = 0, st = 0;
        smlvm_oracle_sparsemap_fwd(kind, n, budget, xd.data(), nullptr, nullptr,
                                   max_iter, cap, p.data(), aset.data(),
                                   nullptr, &c, &it, &st, nullptr);
        if (counts) counts[b] = c;
        if (iters) iters[b] = it;
        total += c;
    }
    return total;
}

} // extern "C"
