// SparseMAP active-set solver for sm_100a: one CTA per sample.
//
// replaces, per sample, SparseMAP.run_sparsemap (lvmhelpers/sparsemap.py:13-36): building a
// factor graph with 2n binary variables, `init_active_set_from_scores`, `solve_qp` (libad3
// GenericFactor::SolveQP) and `get_sparse_solution`, for the factors of
// lvmhelpers/bernoulli.h, budget.h and sequence_binary.h; and SparseMAP.jv
// (sparsemap.py:38-44, dist_jacobian_vec) for the backward.
//
// The algorithm is AD3's active-set method, followed step by step so that pivots, tie rules
// and hence (p, aset) agree with it:
//   state  A (active vertices, ordered), p (distribution), inv = inverse of [[0,1'],[1,M'M]]
//   loop   if A changed: b = [1, score(y_j)],  [tau; z] = inv b; line search p -> z, first
//          blocking constraint leaves A (block down-date of inv);
//          else: u = M z, y^ = MAP(eta - u); stop if value <= tau + 1e-9; else y^ joins A
//          (block update of inv, Schur complement s = <y^,y^> - r' inv r).
// Everything lives in shared memory in fp64, in SOLVER ORDER (index a + 1 of inv = a-th vertex of the
// active set, as in libad3): every sum walks memory directly, an insertion appends a row and a
// column, a removal (rare: ~1 per sample for Bernoulli, ~10 for the budget) closes the gap while it
// down-dates.  Nothing is ever zero-filled: loops are bounded by |A|.  Vertices are bitmasks (Gram entries are n - popc(a ^ b)), MAP oracles as block-wide primitives
// (ballot for Bernoulli, rank-select for the budget, 2-state Viterbi for the sequence).
// No FMA contraction in the inverse updates: the pivoting sequence is sensitive to last bits.
#include <math_constants.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace smlvm {

// Development aid (make timing -> tools/_dev/libsmlvm_timing.so, never shipped): cycle totals of selected solver
// phases (SMLVM_TIMING_MASK: one bit per phase; a recorded phase collects everything since the previous recorded
// point, so few live counters -- no spills), accumulated by thread 0 of every CTA; smlvm_debug_phase_cycles().
#ifdef SMLVM_SMAP_TIMING
#ifndef SMLVM_TIMING_MASK
#define SMLVM_TIMING_MASK 0xffff
#endif
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_BEGIN() long long ph_last = clock64(); unsigned long long ph_acc[16] = {0}
#define PHASE(k) do { if (((SMLVM_TIMING_MASK >> (k)) & 1) && threadIdx.x == 0) { const long long ph_now = clock64(); ph_acc[k] += (unsigned long long)(ph_now - ph_last); ph_last = ph_now; } } while (0)
#define PHASE_END() do { if (threadIdx.x == 0) for (int k = 0; k < 16; ++k) if ((SMLVM_TIMING_MASK >> k) & 1) atomicAdd(&g_phase_cycles[k], ph_acc[k]); } while (0)
#else
#define PHASE_BEGIN()
#define PHASE(k)
#define PHASE_END()
#endif

struct SolverSmem {
    double* inv;    // [LD*LD]  index 0 = constraint row, vertex a -> a + 1
    double* score;  // [cap]    score(y_a) under eta
    double* p;      // [cap]
    double* z;      // [cap]
    double* d;      // [LD]
    double* r;      // [LD]
    double* eta;    // [n]      on-scores (off-scores are 0)
    double* t;      // [n]      transition scores (n-1 used)
    double* son;    // [n]      oracle input, "on" half
    double* soff;   // [n]      oracle input, "off" half
    double* red;    // [40]     reduction scratch
    double* terms;  // [terms_doubles(cap)] Schur-complement terms in libad3's summation order (double buffer)
    uint32_t* bits; // [cap*W]
    uint32_t* ynew; // [W]
    unsigned char* bp;  // [2*n] Viterbi back-pointers
    int* misc;      // [8]
    // general chain only (SMLVM_FACTOR_SEQUENCE), behind everything else:
    double* te;     // [E]      edge scores in the order of FactorSequence::Initialize (sequence.h:264-291)
    double* val;    // [2][32]  Viterbi values of the previous / current position
    int* lay;       // [2L+3]   offset of the states of position i (L+1 entries), first edge of position i (L+2)
    int* ys;        // [L]      state sequence of the vertex in ynew
};

// The general chain (lvmhelpers/sequence.h; PFactorSequence.initialize(list[int])): L positions, position i with
// S_i states, one variable (bit) per (position, state) -- a vertex is a one-hot pattern per position, n = sum S_i bits
// -- and E = S_0 + sum S_{i-1} S_i + S_{L-1} additional (edge) scores.  lay = [off_0..off_L | ebase_0..ebase_{L+1}].
struct SeqArgs {
    const int32_t* lay;  // device
    int L, E;
};
__host__ __device__ inline size_t seq_extra_bytes(int L, int E)
{
    return L > 0 ? (((size_t)(E + 64) * 8 + (size_t)(3 * L + 3) * 4 + 15) & ~(size_t)15) : 0;
}

// The Schur terms of one insertion ((|A|+1)(|A|+2)/2 of them) are produced in rounds of whole matrix rows
// into one half of a double buffer of kTermsChunk doubles per half while thread 0 folds the other half
// (a row has at most n + 2 <= kTermsChunk terms).  Up to |A| = 18 a single round, single buffer.
constexpr int kTermsChunk = 192;
static_assert(kTermsChunk >= SMLVM_SPARSEMAP_MAX_BITS + 2 && kTermsChunk % 2 == 0, "a matrix row must fit one round");
__host__ __device__ inline int terms_doubles(int cap)
{
    const int full = (cap + 1) * (cap + 2) / 2;
    return full <= kTermsChunk ? ((full + 1) & ~1) : 2 * kTermsChunk;
}

// `cap` = largest active set this launch can hold (a capacity tier, <= n + 1)
__host__ __device__ inline size_t solver_smem_bytes(int n, int cap, size_t extra = 0)
{
    const int LD = (cap + 1) | 1, W = (n + 31) / 32;
    const int LV = LD > n ? LD : n;  // d doubles as the budget oracle's gain vector [n]
    const int ne = (n + 1) & ~1, ce = terms_doubles(cap);
    size_t dbl = (size_t)LD * LD + 3 * cap + 2 * LV + 2 * n + 2 * ne + 40 + (size_t)ce;
    size_t b = dbl * sizeof(double);
    b += (size_t)8 * sizeof(int);
    b += (size_t)(cap * W + W) * sizeof(uint32_t);
    b += (size_t)2 * n + 16;
    return ((b + 15) & ~(size_t)15) + extra;
}

__device__ inline SolverSmem carve(unsigned char* raw, int n, int cap, int L = 0, int E = 0)
{
    const int LD = (cap + 1) | 1, W = (n + 31) / 32;
    const int ne = (n + 1) & ~1, ce = terms_doubles(cap);
    SolverSmem s;
    double* d = reinterpret_cast<double*>(raw);
    // the three arrays that are folded in sequence come first: 16-byte aligned, son/soff contiguous
    s.terms = d; d += ce;
    s.son = d; d += ne;
    s.soff = d; d += ne;
    s.inv = d; d += (size_t)LD * LD;
    s.score = d; d += cap;
    s.p = d; d += cap;
    s.z = d; d += cap;
    const int LV = LD > n ? LD : n;
    s.d = d; d += LV;
    s.r = d; d += LV;
    s.eta = d; d += n;
    s.t = d; d += n;
    s.red = d; d += 40;
    int* ip = reinterpret_cast<int*>(d);
    s.misc = ip; ip += 8;
    uint32_t* up = reinterpret_cast<uint32_t*>(ip);
    s.bits = up; up += cap * W;
    s.ynew = up; up += W;
    s.bp = reinterpret_cast<unsigned char*>(up);
    s.te = reinterpret_cast<double*>(raw + solver_smem_bytes(n, cap));
    s.val = s.te + E;
    s.lay = reinterpret_cast<int*>(s.val + 64);
    s.ys = s.lay + 2 * L + 3;
    return s;
}

__device__ __forceinline__ int block_sum_int(int v, double* red)
{
    return (int)(block_sum((double)v, red) + 0.5);
}

// ---- factor MAP oracles: read (son, soff[, t]), write the vertex to sm.ynew, return value ----

// lvmhelpers/bernoulli.h:14-35.  Thread i reads the scores it wrote itself; one barrier.
__device__ double map_bernoulli(const SolverSmem& sm, int n)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int n32 = (n + 31) & ~31;
    double part = 0.0;
    for (int i = threadIdx.x; i < n32; i += blockDim.x) {
        const bool valid = i < n;
        const double on = valid ? sm.son[i] : 0.0, off = valid ? sm.soff[i] : 0.0;
        const bool y = valid && (on > off);
        if (valid) part += y ? on : off;
        const unsigned b = __ballot_sync(kFull, y);
        if (lane == 0) sm.ynew[i >> 5] = b;
    }
    if (warp * 32 < n32) part = group_sum<32>(part);
    if (lane == 0) sm.red[warp] = part;
    __syncthreads();
    double tot = 0.0;
    const int wn = min(nw, n32 >> 5);
    for (int w = 0; w < wn; ++w) tot += sm.red[w];
    return tot;
}

// lvmhelpers/budget.h:24-85.  Uses sm.d as the gain vector.
__device__ double map_budget(const SolverSmem& sm, int n, int budget)
{
    const int lane = threadIdx.x & 31;
    const int n32 = (n + 31) & ~31;
    double base = 0.0, sum = 0.0;
    int cnt = 0;
    for (int i = threadIdx.x; i < n32; i += blockDim.x) {
        const bool valid = i < n;
        const double gain = valid ? __dsub_rn(sm.son[i], sm.soff[i]) : -1.0;
        if (valid) {
            sm.d[i] = gain;
            base += sm.soff[i];
        }
        const bool y = valid && gain > 0.0;
        if (y) { sum += gain; ++cnt; }
        const unsigned b = __ballot_sync(kFull, y);
        if (lane == 0) sm.ynew[i >> 5] = b;
    }
    // (one barrier-with-count when every thread owns at most one coordinate; also orders the sm.d / sm.ynew writes)
    const int active = n <= (int)blockDim.x ? __syncthreads_count(cnt) : block_sum_int(cnt, sm.red);
    if (active > budget) {
        // keep the `budget` largest gains, ties towards the lower index (sort of (-gain, i))
        sum = 0.0;
        for (int i = threadIdx.x; i < n32; i += blockDim.x) {
            const bool valid = i < n;
            bool y = false;
            if (valid) {
                const double gi = sm.d[i];
                int rank = 0;
                int j = 0;
                for (; j + 8 <= n; j += 8) {   // eight gains loaded ahead of their compares (this loop ran at one
                    double gj[8];              // shared-memory latency per element: 17 % of the budget-64 solve)
#pragma unroll
                    for (int u = 0; u < 8; ++u) gj[u] = sm.d[j + u];
#pragma unroll
                    for (int u = 0; u < 8; ++u) rank += (gj[u] > gi || (gj[u] == gi && j + u < i)) ? 1 : 0;
                }
                for (; j < n; ++j) {
                    const double gj = sm.d[j];
                    rank += (gj > gi || (gj == gi && j < i)) ? 1 : 0;
                }
                y = rank < budget && !(gi < 0.0);
                if (y) sum += gi;
            }
            const unsigned b = __ballot_sync(kFull, y);
            if (lane == 0) sm.ynew[i >> 5] = b;
        }
    }
    const double tot = block_sum(base + sum, sm.red);
    return tot;
}

// lvmhelpers/sequence.h:74-151 with the scores of lvmhelpers/sequence_binary.h:10-37.
// `edge` == nullptr means all transition scores are 0 (init_active_set_from_scores).
__device__ double map_sequence(const SolverSmem& sm, int n, const double* edge)
{
    const int W = (n + 31) / 32;
    if (threadIdx.x == 0) {
        double v0 = sm.soff[0], v1 = sm.son[0];
        for (int i = 0; i + 1 < n; ++i) {
            const double e = edge ? edge[i] : 0.0;
            // into state 0: l=0 gives v0, l=1 gives v1 (edge 0); into state 1: v0, v1 + e
            const int b0 = (v1 > v0) ? 1 : 0;
            const double c1 = __dadd_rn(v1, e);
            const int b1 = (c1 > v0) ? 1 : 0;
            const double n0 = __dadd_rn(b0 ? v1 : v0, sm.soff[i + 1]);
            const double n1 = __dadd_rn(b1 ? c1 : v0, sm.son[i + 1]);
            sm.bp[2 * (i + 1)] = (unsigned char)b0;
            sm.bp[2 * (i + 1) + 1] = (unsigned char)b1;
            v0 = n0;
            v1 = n1;
        }
        int state = (v1 > v0) ? 1 : 0;
        sm.red[39] = state ? v1 : v0;
        for (int q = 0; q < W; ++q) sm.ynew[q] = 0u;
        for (int i = n - 1; i >= 0; --i) {
            if (state) sm.ynew[i >> 5] |= 1u << (i & 31);
            if (i > 0) state = sm.bp[2 * i + state];
        }
    }
    __syncthreads();
    const double v = sm.red[39];
    __syncthreads();
    return v;
}

// lvmhelpers/sequence.h:74-151, general chain.  Warp 0: lane k owns state k of the current position (S_i <= 32),
// the values of the previous position sit in sm.val, back-pointers in sm.bp
// (one per variable).  Additions in the reference's order: values[i][l] + edge, then + node; among equal
// predecessors the lower state wins (`best < 0 || val > best_value`, sequence.h:112,136).
__device__ double map_chain(const SolverSmem& sm, int n, int L)
{
    const int W = (n + 31) / 32;
    const int* off = sm.lay;
    const int* eb = sm.lay + L + 1;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double* val = sm.val;             // [2][32]
        int S0 = off[1] - off[0];
        if (lane < S0) val[lane] = __dadd_rn(sm.son[off[0] + lane], sm.te[eb[0] + lane]);
        __syncwarp();
        int cur = 0;
        for (int i = 0; i + 1 < L; ++i) {
            const int Sp = off[i + 1] - off[i], Sc = off[i + 2] - off[i + 1];
            if (lane < Sc) {
                double best_value = 0.0;
                int best = -1;
                for (int l = 0; l < Sp; ++l) {
                    const double v = __dadd_rn(val[cur * 32 + l], sm.te[eb[i + 1] + l * Sc + lane]);
                    if (best < 0 || v > best_value) { best_value = v; best = l; }
                }
                val[(cur ^ 1) * 32 + lane] = __dadd_rn(best_value, sm.son[off[i + 1] + lane]);
                sm.bp[off[i + 1] + lane] = (unsigned char)best;
            }
            cur ^= 1;
            __syncwarp();
        }
        if (lane == 0) {
            const int Sl = off[L] - off[L - 1];
            double best_value = 0.0;
            int best = -1;
            for (int l = 0; l < Sl; ++l) {
                const double v = __dadd_rn(val[cur * 32 + l], sm.te[eb[L] + l]);
                if (best < 0 || v > best_value) { best_value = v; best = l; }
            }
            sm.red[39] = best_value;
            for (int q = 0; q < W; ++q) sm.ynew[q] = 0u;
            int state = best;
            for (int i = L - 1; i >= 0; --i) {
                sm.ys[i] = state;
                const int b = off[i] + state;
                sm.ynew[b >> 5] |= 1u << (b & 31);
                if (i > 0) state = sm.bp[off[i] + state];
            }
        }
    }
    __syncthreads();
    const double v = sm.red[39];
    __syncthreads();
    return v;
}

template <int KIND>
__device__ __forceinline__ double factor_map(const SolverSmem& sm, int n, int budget,
                                             const double* edge, int L = 0)
{
    if (KIND == SMLVM_FACTOR_BERNOULLI) return map_bernoulli(sm, n);
    if (KIND == SMLVM_FACTOR_BUDGET) return map_budget(sm, n, budget);
    if (KIND == SMLVM_FACTOR_SEQUENCE) return map_chain(sm, n, L);
    return map_sequence(sm, n, edge);
}

// ---- order-sensitive sums, done the way libad3 does them -------------------------------------
// At every oracle call the reduced scores of all coordinates that are already "solved" tie to
// within rounding, so WHICH vertex the MAP oracle returns is decided by the last bits of z.  To
// follow the reference's pivoting the sums that feed z must be rounded in its order.  Row-wise
// sums (z, tau, d, u) are sequential per thread anyway; the two whole-matrix / whole-vector sums
// below are chains: their terms are formed in parallel into shared memory and folded by ONE thread.

// acc = (((acc -+ t_0) -+ t_1) -+ ...) over `total` doubles of a 16-byte aligned shared array, by ONE
// thread.  The next eight terms are loaded (128-bit) while the current eight are folded, so the chain
// advances at DADD latency instead of load + DADD latency.
template <bool SUB>
__device__ __forceinline__ double fold_serial(const double* terms, int total, double s)
{
    const double2* t2 = reinterpret_cast<const double2*>(terms);
    int e = 0;
    if (total >= 8) {
        double2 c0 = t2[0], c1 = t2[1], c2 = t2[2], c3 = t2[3];
        for (e = 8; e + 8 <= total; e += 8) {
            const double2 n0 = t2[(e >> 1)], n1 = t2[(e >> 1) + 1], n2 = t2[(e >> 1) + 2], n3 = t2[(e >> 1) + 3];
#define SMLVM_FOLD(v) s = SUB ? __dsub_rn(s, (v)) : __dadd_rn(s, (v))
            SMLVM_FOLD(c0.x); SMLVM_FOLD(c0.y); SMLVM_FOLD(c1.x); SMLVM_FOLD(c1.y);
            SMLVM_FOLD(c2.x); SMLVM_FOLD(c2.y); SMLVM_FOLD(c3.x); SMLVM_FOLD(c3.y);
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        }
        SMLVM_FOLD(c0.x); SMLVM_FOLD(c0.y); SMLVM_FOLD(c1.x); SMLVM_FOLD(c1.y);
        SMLVM_FOLD(c2.x); SMLVM_FOLD(c2.y); SMLVM_FOLD(c3.x); SMLVM_FOLD(c3.y);
    }
    for (; e < total; ++e) SMLVM_FOLD(terms[e]);
#undef SMLVM_FOLD
    return s;
}

// score of the vertex in sm.ynew under (eta, 0[, t]): bernoulli.h:38-54 / sequence.h:154-179 -- a
// sequential fp64 sum in bit order (adding 0.0 for a clear bit is exact, so only set bits count).
// One WARP: the set bits' scores are compacted in bit order into `list` (son/soff, free between two
// oracle calls; for the sequence every set bit is followed by its transition score or an exact 0.0),
// then lane 0 folds the dense list.  Result valid in lane 0.
template <int KIND>
__device__ double evaluate_warp(const SolverSmem& sm, int n, int L = 0)
{
    const int lane = threadIdx.x & 31;
    if (KIND == SMLVM_FACTOR_SEQUENCE) {
        // sequence.h:154-179: node then edge at every position (sm.ys = the states of ynew, written by the oracle
        // that produced it), the stop edge last -- one sequential fp64 sum
        double acc = 0.0;
        if (lane == 0) {
            const int* off = sm.lay;
            const int* eb = sm.lay + L + 1;
            int prev = 0;
            for (int i = 0; i < L; ++i) {
                const int st = sm.ys[i], Sc = off[i + 1] - off[i];
                acc = __dadd_rn(acc, sm.eta[off[i] + st]);
                acc = __dadd_rn(acc, sm.te[eb[i] + prev * Sc + st]);
                prev = st;
            }
            acc = __dadd_rn(acc, sm.te[eb[L] + prev]);
        }
        return acc;
    }
    const int W = (n + 31) >> 5;
    double* list = sm.son;  // [2 * even(n)] with soff
    int base = 0;
    uint32_t carry = 0u;  // bit 31 of the previous word (the left neighbour of bit 0)
    for (int q = 0; q < W; ++q) {
        const uint32_t word = sm.ynew[q];
        const int i = q * 32 + lane;
        const bool on = (word >> lane) & 1u;
        const int rank = base + __popc(word & ((1u << lane) - 1u));
        if (KIND == SMLVM_FACTOR_SEQUENCE_BINARY) {
            const uint32_t pairs = word & ((word << 1) | carry);  // bit i set with bit i-1 set
            if (on) {
                list[2 * rank] = sm.eta[i];
                list[2 * rank + 1] = ((pairs >> lane) & 1u) ? sm.t[i - 1] : 0.0;
            }
        } else if (on) {
            list[rank] = sm.eta[i];
        }
        base += __popc(word);
        carry = word >> 31;
    }
    __syncwarp();
    double acc = 0.0;
    if (lane == 0)
        acc = fold_serial<false>(list, KIND == SMLVM_FACTOR_SEQUENCE_BINARY ? 2 * base : base, 0.0);
    return acc;
}

// Schur complement s = r0 - sum_{i<=j} c_ij r_i r_j inv_ij in libad3's order (rows in solver
// order, c_ii = 1, c_ij = 2; libad3 skips r_i == 0 / r_j == 0, whose terms are exact zeros).
// Two steps: the terms are formed in parallel into one half of sm.terms (row-major upper triangle,
// rounds of whole rows), then ONE thread folds them in sequence -- a pure DSUB dependency chain fed by
// 128-bit shared loads issued a batch ahead (ncu r01: a shuffle-fed chain cost ~45 cycles per term and
// kept the other 7 warps at the barrier 60 % of the time; this one runs at ~11).
// rows [a0, a1) of the upper triangle -> out (row-major, packed), by warps w0 .. nw-1 of the CTA
__device__ __forceinline__ void schur_terms_fill(const SolverSmem& sm, double* dst, int m, int LD, int a0, int a1,
                                                 int w0)
{
    const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) - w0, nw = (blockDim.x >> 5) - w0;
    if (warp < 0) return;
    for (int a = a0 + warp; a < a1; a += nw) {
        const double ri = sm.r[a];
        // rows a0..a-1 hold (m+1-a0) + ... + (m+2-a) terms
        const int off = (a - a0) * (m + 1) - ((a - a0) * (a0 + a - 1)) / 2;
        double* out = dst + off - a;  // + b
        for (int b = a + lane; b <= m; b += 32) {
            const double v = sm.inv[a * LD + b];
            out[b] = (b == a) ? __dmul_rn(__dmul_rn(ri, ri), v)
                              : __dmul_rn(__dmul_rn(__dmul_rn(2.0, ri), sm.r[b]), v);
        }
    }
}
// rows [a0, a1) that go into one round of the terms buffer: row a has m + 1 - a terms, so chunk / (m + 1 - a0) rows
// from a0 on hold at most `chunk` terms (at least one row: a row has <= cap + 1 <= chunk terms).  Closed form: this
// runs on the fold thread's critical path once per round (a search loop here cost 8 % of the kernel's instructions).
__device__ __forceinline__ int schur_round_end(int m, int a0, int chunk, int& count)
{
    const int rest = ((m + 1 - a0) * (m + 2 - a0)) / 2;
    if (rest <= chunk) {  // everything that is left fits (|A| <= 18: the only round)
        count = rest;
        return m + 1;
    }
    const int len = m + 1 - a0;
    const int rows = chunk / len;
    const int a1 = a0 + rows;  // < m + 1 because rest > chunk
    count = rows * len - (rows * (rows - 1)) / 2;
    return a1;
}

// CountCommonValues: equal positions of two vertices.  Independent bits (on / off per position): n - Hamming
// distance; one-hot chain: the positions where both have the same state bit set.
template <int KIND>
__device__ __forceinline__ int common_values(const uint32_t* a, const uint32_t* b, int W, int n)
{
    int c = 0;
    for (int q = 0; q < W; ++q) c += (KIND == SMLVM_FACTOR_SEQUENCE) ? __popc(a[q] & b[q]) : __popc(a[q] ^ b[q]);
    return (KIND == SMLVM_FACTOR_SEQUENCE) ? c : n - c;
}

// A sample that outgrows its capacity tier is handed to the next one WITH its solver state (active set, p, z, scores,
// inverse, iteration count, tau): the next tier resumes at the insertion that did not fit instead of re-solving from
// scratch (the samples that overflow are the long ones: before this, the second tier cost 5 of 33.5 ms for 2.7 % of
// the Bernoulli samples at n = 128, 27 of 120 ms for 6.7 % of budget-64).  Records are bump-allocated from whatever
// the caller's workspace holds beyond its mandatory part; a record that does not fit means "re-solve" (offset -1).
struct TierHandoff {
    unsigned char* pool;            // record storage (NULL: always re-solve)
    long long pool_bytes;
    unsigned long long* cursor;     // bytes handed out
    const long long* rec_in;        // [work] record offset of the worklist entries of THIS tier (NULL: first tier)
    long long* rec_out;             // [..]   record offsets for the next tier's worklist
};

template <int KIND>
__global__ void __launch_bounds__(256, 4) sparsemap_fwd_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                     const double* __restrict__ init_scores, int budget,
                                     int max_iter, int64_t rows, int n, int cap, int last_tier,
                                     const int32_t* __restrict__ wl_in, const int32_t* __restrict__ wl_in_count,
                                     int32_t* __restrict__ wl_out, int32_t* __restrict__ wl_out_count,
                                     float* __restrict__ p_pad,
                                     uint32_t* __restrict__ aset_bits, int32_t* __restrict__ counts,
                                     int32_t* __restrict__ kept, int32_t* __restrict__ iters,
                                     int32_t* __restrict__ status, double* __restrict__ S_arena,
                                     long long S_capacity, long long* __restrict__ S_offsets,
                                     unsigned long long* __restrict__ S_cursor, SeqArgs seq, TierHandoff ho)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cap_out = n + 1;                    // layout of the padded outputs (ABI)
    const int LD = (cap + 1) | 1, W = (n + 31) / 32;  // odd LD: row-wise reads spread over banks
    const SolverSmem sm = carve(smem_raw, n, cap, seq.L, seq.E);
    const int tid = threadIdx.x, T = blockDim.x;
    // positions of a vertex = common(y, y): the n bits, or the L positions of the one-hot chain
    const int npos = KIND == SMLVM_FACTOR_SEQUENCE ? seq.L : n;
    if (KIND == SMLVM_FACTOR_SEQUENCE)
        for (int i = tid; i < 2 * seq.L + 3; i += T) sm.lay[i] = __ldg(seq.lay + i);

    // the first capacity tier takes every sample; later tiers only the samples an earlier tier could not hold,
    // from the worklist that tier appended them to (a small persistent grid: no empty CTAs)
    // (a later tier's CTAs pull the next entry with an atomic ticket: its few samples are the long ones, a static
    // assignment would leave most CTAs idle behind the unlucky one)
    const int64_t work = wl_in != nullptr ? (int64_t)__ldg(wl_in_count) : rows;
    int* ticket = wl_in != nullptr ? const_cast<int*>(wl_in_count) + 1 : nullptr;
    for (int64_t wi = blockIdx.x;; wi += gridDim.x) {
        __syncthreads();
        if (wl_in != nullptr) {
            if (threadIdx.x == 0) sm.misc[7] = atomicAdd(ticket, 1);
            __syncthreads();
            wi = sm.misc[7];
        }
        if (wi >= work) break;
        const int64_t row = wl_in != nullptr ? (int64_t)__ldg(wl_in + wi) : wi;
        PHASE_BEGIN();
        // ---- load the sample -----------------------------------------------------------
        for (int i = tid; i < n; i += T) {
            sm.eta[i] = (double)x[row * n + i];
            sm.t[i] = (KIND == SMLVM_FACTOR_SEQUENCE_BINARY && t != nullptr && i + 1 < n)
                          ? (double)t[row * (int64_t)(n - 1) + i] : 0.0;
            if (init_scores != nullptr) {
                sm.son[i] = init_scores[row * 2 * (int64_t)n + i];
                sm.soff[i] = init_scores[row * 2 * (int64_t)n + n + i];
            } else {
                sm.son[i] = (double)x[row * n + i];
                sm.soff[i] = 0.0;
            }
        }
        if (KIND == SMLVM_FACTOR_SEQUENCE)
            for (int e = tid; e < seq.E; e += T) sm.te[e] = t != nullptr ? (double)t[row * (int64_t)seq.E + e] : 0.0;
        __syncthreads();

        int m = 1;          // |A|
        bool changed = true;
        double tau = 0.0;
        int st = SMLVM_SOLVE_MAX_ITER;
        int it = 0, it0 = 0;
        const int eval_warp = T > 32 ? 1 : 0;  // folds the new vertex's score while thread 0 folds the Schur terms
        const long long rec = (wl_in != nullptr && ho.rec_in != nullptr) ? __ldg(ho.rec_in + wi) : -1;
        if (rec >= 0) {
            // ---- resume: the state an earlier tier left when the active set outgrew it -----------------
            const double* rd = reinterpret_cast<const double*>(ho.pool + rec);
            m = (int)rd[0];
            it0 = (int)rd[1];
            tau = rd[2];
            changed = false;
            const double* rp = rd + 4;
            for (int a = tid; a < m; a += T) {
                sm.p[a] = rp[a];
                sm.z[a] = rp[m + a];
                sm.score[a] = rp[2 * m + a];
            }
            const double* ri = rp + 3 * m;
            for (int e = tid; e < (m + 1) * (m + 1); e += T) {
                const int a = e / (m + 1), c = e - a * (m + 1);
                sm.inv[a * LD + c] = ri[e];
            }
            const uint32_t* rb = reinterpret_cast<const uint32_t*>(ri + (m + 1) * (m + 1));
            for (int e = tid; e < m * W; e += T) sm.bits[e] = rb[e];
        } else {
        // ---- seed: single MAP vertex, weight 1, inv = [[-n, 1], [1, 0]] ---------------------
        // (init_active_set_from_scores, sparsemap.py:25-27, or the LP solution under eta)
        factor_map<KIND>(sm, n, budget, init_scores != nullptr ? nullptr : sm.t, seq.L);
        __syncthreads();
        if (tid < 32) {
            const double sc0 = evaluate_warp<KIND>(sm, n, seq.L);
            if (tid == 0) {
                sm.score[0] = sc0;
                sm.p[0] = 1.0;
                sm.z[0] = 1.0;
                sm.inv[0] = -(double)npos;
                sm.inv[1] = 1.0;
                sm.inv[LD] = 1.0;
                sm.inv[LD + 1] = 0.0;
            }
        }
        for (int q = tid; q < W; q += T) sm.bits[q] = sm.ynew[q];
        }
        __syncthreads();

        // Barrier discipline: every phase below ends with ONE barrier; what a phase reads was written
        // before the previous barrier, what it writes is not read by anyone inside the same phase.
        for (it = it0; it < max_iter; ++it) {
            if (changed) {
                // ---- [tau; z] = inv * [1; score]; a blocking constraint exists iff some z_a < p_a is negative:
                //      the test rides on the barrier that publishes z, so the common pass (none) computes no
                //      ratios and needs no second barrier ------------------------------------------------------
                int neg = 0;
                for (int a = tid; a <= m; a += T) {
                    const double* row = sm.inv + a * LD;
                    double acc = row[0];  // column 0 times b_0 = 1
#pragma unroll 8
                    for (int j = 1; j <= m; ++j) acc = __dadd_rn(acc, __dmul_rn(row[j], sm.score[j - 1]));
                    if (a == 0) {
                        sm.red[38] = acc;
                    } else {
                        sm.z[a - 1] = acc;
                        neg |= (acc < sm.p[a - 1] && acc < 0.0) ? 1 : 0;
                    }
                }
                const bool exist_blocking = __syncthreads_or(neg) != 0;
                PHASE(0);
                tau = sm.red[38];
                if (!exist_blocking) {
                    // p is next read by a later blocking test (barriers in between): no barrier here
                    for (int a = tid; a < m; a += T) sm.p[a] = sm.z[a];
                    changed = false;
                } else {
                    // ---- line search: first minimal ratio p/(p-z) over z < p -----------------
                    if (tid < 32) {
                        double best = CUDART_INF;
                        int besti = 0x7fffffff;
                        // four candidates per lane and pass: their divisions are independent and overlap
                        for (int a0 = tid; a0 < m; a0 += 128) {
                            double ratio[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int a = a0 + 32 * u;
                                ratio[u] = CUDART_INF;
                                if (a < m) {
                                    const double zz = sm.z[a], pp = sm.p[a];
                                    if (zz < pp) ratio[u] = __ddiv_rn(pp, __dsub_rn(pp, zz));
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u)    // ascending index: the first minimum wins
                                if (ratio[u] < best) { best = ratio[u]; besti = a0 + 32 * u; }
                        }
                        // lexicographic minimum of (ratio, index) over the warp: order-preserving 64-bit key of
                        // the double (sign flip), reduced as two 32-bit halves, then the lowest index holding it
                        const long long bb = __double_as_longlong(__dadd_rn(best, 0.0));  // (-0.0 -> +0.0: numeric ties)
                        const unsigned long long key = (unsigned long long)bb ^ ((unsigned long long)(bb >> 63) | 0x8000000000000000ull);
                        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
                        const bool cand = besti != 0x7fffffff;
                        const unsigned mhi = __reduce_min_sync(kFull, cand ? hi : 0xffffffffu);
                        const bool e1 = cand && hi == mhi;
                        const unsigned mlo = __reduce_min_sync(kFull, e1 ? lo : 0xffffffffu);
                        const bool e2 = e1 && lo == mlo;
                        const unsigned mi = __reduce_min_sync(kFull, e2 ? (unsigned)besti : 0x7fffffffu);
                        const double gbest = __shfl_sync(kFull, best, mi & 31);  // lane (index mod 32) owns that candidate
                        if (tid == 0) {
                            sm.misc[1] = (int)mi;
                            sm.red[37] = gbest;
                        }
                    }
                    __syncthreads();
                    const int blocking = sm.misc[1];
                    double alpha = sm.red[37];
                    if (alpha > 1.0) alpha = 1.0;
                    const int rem = blocking + 1;
                    for (int a = tid; a < m; a += T) {
                        if (alpha == 1.0) {
                            sm.p[a] = sm.z[a];
                        } else {
                            const double v = __dadd_rn(__dmul_rn(__dsub_rn(1.0, alpha), sm.p[a]),
                                                       __dmul_rn(alpha, sm.z[a]));
                            sm.z[a] = v;
                            sm.p[a] = v;
                        }
                    }
                    // ---- down-date inv: d = -(1/inv_rr) inv[r,:],  inv -= inv_rr d d', then drop
                    //      row and column r (the rest slides up / left: solver order is memory order)
                    const double invs = sm.inv[rem * LD + rem];
                    const double sval = __ddiv_rn(1.0, invs);
                    for (int i = tid; i <= m; i += T)
                        sm.d[i] = (i == rem) ? 0.0 : __dmul_rn(-sval, sm.inv[rem * LD + i]);
                    __syncthreads();
                    // pass 1, a warp per row: down-date and close the column gap inside the row
                    // (32 columns at a time: read, then write one place to the left beyond r)
                    for (int i = tid >> 5; i <= m; i += T >> 5) {
                        if (i == rem) continue;
                        const double di = __dmul_rn(invs, sm.d[i]);
                        double* row = sm.inv + i * LD;
                        for (int j0 = 0; j0 <= m; j0 += 32) {
                            const int j = j0 + (tid & 31);
                            double v = 0.0;
                            if (j <= m) v = __dsub_rn(row[j], __dmul_rn(di, sm.d[j]));
                            __syncwarp();
                            if (j <= m && j != rem) row[j - (j > rem ? 1 : 0)] = v;
                            __syncwarp();
                        }
                    }
                    __syncthreads();
                    // pass 2, a thread per column: rows beyond r move up; the last threads slide the
                    // per-vertex vectors and bitmasks down over the removed vertex
                    for (int j = tid; j < m; j += T)
                        for (int i = rem; i < m; ++i) sm.inv[i * LD + j] = sm.inv[(i + 1) * LD + j];
                    {
                        const int v = T - 1 - tid;  // 0: score, 1: p, 2: z, 3..: one bitmask word each
                        if (v < 3) {
                            double* vec = v == 0 ? sm.score : (v == 1 ? sm.p : sm.z);
                            for (int a = blocking; a + 1 < m; ++a) vec[a] = vec[a + 1];
                        } else if (v < 3 + W) {
                            uint32_t* col = sm.bits + (v - 3);
                            for (int a = blocking; a + 1 < m; ++a) col[a * W] = col[(a + 1) * W];
                        }
                    }
                    --m;
                    changed = true;
                    __syncthreads();
                }
            } else {
                // ---- u = M z, oracle scores eta - u ----------------------------------------------
                for (int i = tid; i < n; i += T) {
                    double uon = 0.0, uoff = 0.0;
                    const uint32_t* bw = sm.bits + (i >> 5);
                    const int sh = i & 31;
#pragma unroll 8
                    for (int a = 0; a < m; ++a) {
                        const double zz = sm.z[a];
                        if ((bw[a * W] >> sh) & 1u) uon = __dadd_rn(uon, zz);
                        else uoff = __dadd_rn(uoff, zz);
                    }
                    sm.son[i] = __dsub_rn(sm.eta[i], uon);
                    if (KIND != SMLVM_FACTOR_SEQUENCE) sm.soff[i] = __dsub_rn(0.0, uoff);   // (the chain has no "off" half)
                }
                // Bernoulli / budget oracles read, per thread, the scores that thread wrote
                if (KIND == SMLVM_FACTOR_SEQUENCE_BINARY || KIND == SMLVM_FACTOR_SEQUENCE) __syncthreads();
                PHASE(3);
                const double value = factor_map<KIND>(sm, n, budget, sm.t, seq.L);  // ends behind a barrier
                PHASE(4);
                if (value <= tau + 1e-9) {
                    st = SMLVM_SOLVE_CONVERGED;
                    ++it;
                    break;
                }
                // ---- insertion: r = [1, common(y_j, y^)]; y^ already active iff common == n ---
                // (libad3 clears its cache in that case; we keep the solution)
                int same = 0;
                if (tid == 0) sm.r[0] = 1.0;
                for (int a = tid; a < m; a += T) {
                    const int c = common_values<KIND>(sm.bits + a * W, sm.ynew, W, n);
                    sm.r[a + 1] = (double)c;
                    same |= (c == npos) ? 1 : 0;
                }
                same = __syncthreads_or(same);
                PHASE(5);
                if (same) {
                    st = SMLVM_SOLVE_REPEATED_CONFIG;
                    ++it;
                    break;
                }
                // ---- d = inv r, s = n - r'd ---------------------------------------------------
                // the two order-sensitive sums: Schur terms formed by the whole CTA, folded by thread 0;
                // the score of the new vertex compacted and folded by another warp; meanwhile the
                // threads at the far end form d_j = sum_i inv[i][j] r_i.
                int cnt_cur = 0;
                int a_end = schur_round_end(m, 0, kTermsChunk, cnt_cur);
                schur_terms_fill(sm, sm.terms, m, LD, 0, a_end, 0);
                __syncthreads();
                PHASE(6);
                double sfold = (double)npos;
                const int fill_w0 = T > 32 ? 1 : 0;  // warp 0 folds: the others fill the next round meanwhile
                for (int half = 0, first = 1;; half ^= 1, first = 0) {
                    int cnt_nxt = 0, a_nxt = a_end;
                    if (a_end <= m) {
                        a_nxt = schur_round_end(m, a_end, kTermsChunk, cnt_nxt);
                        schur_terms_fill(sm, sm.terms + (half ^ 1) * kTermsChunk, m, LD, a_end, a_nxt, fill_w0);
                    }
                    PHASE(13);
                    if (tid == 0) sfold = fold_serial<true>(sm.terms + half * kTermsChunk, cnt_cur, sfold);
                    PHASE(12);
                    if (first) {
                        if ((tid >> 5) == eval_warp) {
                            const double scn = evaluate_warp<KIND>(sm, n, seq.L);
                            if ((tid & 31) == 0) sm.red[35] = scn;
                        }
                        for (int j = T - 1 - tid; j <= m; j += T) {
                            double acc = __dmul_rn(sm.inv[j], sm.r[0]);
#pragma unroll 8
                            for (int b = 1; b <= m; ++b) {
                                const double ri = sm.r[b];
                                if (ri != 0.0) acc = __dadd_rn(acc, __dmul_rn(sm.inv[b * LD + j], ri));
                            }
                            sm.d[j] = acc;
                        }
                    }
                    if (a_end > m) break;  // that was the last round
                    __syncthreads();
                    a_end = a_nxt;
                    cnt_cur = cnt_nxt;
                }
                if (tid == 0) sm.red[36] = sfold;
                __syncthreads();
                PHASE(7);
                const double sval = sm.red[36];
                if (sval > -1e-9 && sval < 1e-9) {
                    st = SMLVM_SOLVE_SINGULAR;
                    ++it;
                    break;
                }
                if (m == cap) {  // the active set outgrows this tier
                    st = last_tier ? SMLVM_SOLVE_SINGULAR : -1;
                    ++it;
                    break;
                }
                const double invs = __ddiv_rn(1.0, sval);
                const int pn = m + 1;
                // ---- block update of inv: old part += d d'/s, new row/column = -d/s, corner 1/s ----
                for (int i = tid >> 5; i <= m; i += T >> 5) {
                    const double di = __dmul_rn(invs, sm.d[i]);
                    double* row = sm.inv + i * LD;
                    for (int j = tid & 31; j <= m; j += 32) row[j] = __dadd_rn(row[j], __dmul_rn(di, sm.d[j]));
                }
                for (int i = tid; i <= m; i += T) {
                    const double v = __dmul_rn(-invs, sm.d[i]);
                    sm.inv[i * LD + pn] = v;
                    sm.inv[pn * LD + i] = v;
                }
                for (int q = tid; q < W; q += T) sm.bits[m * W + q] = sm.ynew[q];
                if (tid == 0) {
                    sm.inv[pn * LD + pn] = invs;
                    sm.score[m] = sm.red[35];
                    sm.p[m] = 0.0;
                    sm.z[m] = 0.0;
                }
                ++m;
                changed = true;
                __syncthreads();
                PHASE(8);
            }
        }
        if (it > max_iter) it = max_iter;

        // ---- outputs (solver order) -----------------------------------------------------
        __syncthreads();
        if (st == -1) {  // hand the sample to the next capacity tier, with its state if the pool has room
            const long long nd = 4 + 3 * (long long)m + (long long)(m + 1) * (m + 1);
            const long long bytes = (nd * 8 + (long long)m * W * 4 + 15) & ~15LL;
            if (tid == 0) {
                const int idx = atomicAdd(wl_out_count, 1);
                wl_out[idx] = (int32_t)row;
                long long off = -1;
                if (ho.pool != nullptr && ho.rec_out != nullptr) {
                    const long long o = (long long)atomicAdd(ho.cursor, (unsigned long long)bytes);
                    if (o + bytes <= ho.pool_bytes) off = o;
                }
                if (ho.rec_out != nullptr) ho.rec_out[idx] = off;
                reinterpret_cast<long long*>(sm.red)[34] = off;
            }
            __syncthreads();
            const long long off = reinterpret_cast<const long long*>(sm.red)[34];
            if (off >= 0) {
                double* rd = reinterpret_cast<double*>(ho.pool + off);
                if (tid == 0) {
                    rd[0] = (double)m;
                    rd[1] = (double)(it - 1);   // the iteration whose insertion did not fit is run again
                    rd[2] = tau;
                    rd[3] = 0.0;
                }
                double* rp = rd + 4;
                for (int a = tid; a < m; a += T) {
                    rp[a] = sm.p[a];
                    rp[m + a] = sm.z[a];
                    rp[2 * m + a] = sm.score[a];
                }
                double* ri = rp + 3 * m;
                for (int e = tid; e < (m + 1) * (m + 1); e += T) {
                    const int a = e / (m + 1), c = e - a * (m + 1);
                    ri[e] = sm.inv[a * LD + c];
                }
                uint32_t* rb = reinterpret_cast<uint32_t*>(ri + (m + 1) * (m + 1));
                for (int e = tid; e < m * W; e += T) rb[e] = sm.bits[e];
            }
            continue;
        }
        int kp = 0;
        for (int a = tid; a < cap_out; a += T) {
            float pf = 0.f;
            if (a < m) {
                pf = (float)sm.p[a];
                kp += pf > 0.f ? 1 : 0;
            }
            p_pad[row * cap_out + a] = pf;
        }
        for (int e = tid; e < cap_out * W; e += T) {
            const int a = e / W, q = e - a * W;
            aset_bits[(row * cap_out + a) * W + q] = (a < m) ? sm.bits[a * W + q] : 0u;
        }
        if (S_arena != nullptr) {
            // |A|^2 doubles of this sample, bump-allocated from the caller's arena; a sample that does not fit
            // gets offset -1 (the cursor still advances: it ends at the size the arena should have had)
            if (tid == 0) {
                const long long off = (long long)atomicAdd(S_cursor, (unsigned long long)(m * m));
                const long long ok = (off + (long long)m * m <= S_capacity) ? off : -1;
                S_offsets[row] = ok;
                reinterpret_cast<long long*>(sm.red)[34] = ok;
            }
            __syncthreads();
            const long long off = reinterpret_cast<const long long*>(sm.red)[34];
            if (off >= 0) {
                double* S = S_arena + off;
                for (int e = tid; e < m * m; e += T) {
                    const int a = e / m, c = e - a * m;
                    S[e] = sm.inv[(a + 1) * LD + c + 1];
                }
            }
        }
        if (kept != nullptr) {
            const int tot = block_sum_int(kp, sm.red);
            if (tid == 0) kept[row] = tot;
        }
        if (tid == 0) {
            counts[row] = m;
            if (iters != nullptr) iters[row] = it;
            if (status != nullptr) status[row] = st;
        }
        PHASE(11);
        PHASE_END();
    }
}


// Ragged expansion: one warp per sample.
__global__ void __launch_bounds__(256)
sparsemap_expand_kernel(const float* __restrict__ p_pad, const uint32_t* __restrict__ aset_bits,
                        const int32_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                        int only_positive, float* __restrict__ p_out, float* __restrict__ aset_out,
                        int32_t* __restrict__ idx_out, int64_t rows, int n)
{
    const int cap = n + 1, W = (n + 31) / 32;
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * 8;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += warps_total) {
        const int m = __ldg(counts + row);
        int64_t out = __ldg(offsets + row);
        for (int a = 0; a < m; ++a) {
            const float pa = __ldg(p_pad + row * cap + a);
            if (only_positive && !(pa > 0.f)) continue;
            if (lane == 0) {
                if (p_out != nullptr) p_out[out] = pa;
                if (idx_out != nullptr) idx_out[out] = (int32_t)row;
            }
            if (aset_out != nullptr) {
                const uint32_t* bw = aset_bits + (row * cap + a) * W;
                for (int i = lane; i < n; i += 32)
                    stg_stream(aset_out + out * n + i, (float)((__ldg(bw + (i >> 5)) >> (i & 31)) & 1u));
            }
            ++out;
        }
    }
}

// Backward: w = S dp (dp scattered to solver order, 0 for dropped vertices), dx = M_on w,
// dt_i = sum_a w_a [y_a,i = y_a,i+1 = 1].  One CTA (128 threads) per sample.
__global__ void __launch_bounds__(128)
sparsemap_bwd_kernel(const float* __restrict__ dp, const int64_t* __restrict__ offsets,
                     int only_positive, const float* __restrict__ p_pad,
                     const uint32_t* __restrict__ aset_bits, const int32_t* __restrict__ counts,
                     const double* __restrict__ S_arena, const long long* __restrict__ S_offsets,
                     float* __restrict__ dx, float* __restrict__ dt, int64_t rows, int n, SeqArgs seq)
{
    extern __shared__ double sh[];
    const int cap = n + 1, W = (n + 31) / 32;
    double* g = sh;        // [cap] upstream gradient in solver order
    double* wv = sh + cap; // [cap]
    const int tid = threadIdx.x, T = blockDim.x;
    const int dt_len = seq.L > 0 ? seq.E : n - 1;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int m = __ldg(counts + row);
        const int64_t base = __ldg(offsets + row);
        __syncthreads();
        if (tid == 0) {
            int rank = 0;
            for (int a = 0; a < m; ++a) {
                const bool keep = !only_positive || (__ldg(p_pad + row * cap + a) > 0.f);
                g[a] = keep ? (double)__ldg(dp + base + rank) : 0.0;
                rank += keep ? 1 : 0;
            }
        }
        __syncthreads();
        const long long soff = __ldg(S_offsets + row);
        if (soff < 0) {  // this sample's inverse did not fit the arena of the forward call: fail loudly
            for (int i = tid; i < n; i += T) dx[row * n + i] = __int_as_float(0x7fc00000);
            if (dt != nullptr)
                for (int i = tid; i < dt_len; i += T) dt[row * (int64_t)dt_len + i] = __int_as_float(0x7fc00000);
            continue;
        }
        const double* S = S_arena + soff;
        for (int a = tid; a < m; a += T) {
            double acc = 0.0;
            for (int c = 0; c < m; ++c) acc += S[(int64_t)c * m + a] * g[c];  // S symmetric
            wv[a] = acc;
        }
        __syncthreads();
        for (int i = tid; i < n; i += T) {
            double on = 0.0, edge = 0.0;
            for (int a = 0; a < m; ++a) {
                const uint32_t* bw = aset_bits + (row * cap + a) * W;
                const bool yi = (__ldg(bw + (i >> 5)) >> (i & 31)) & 1u;
                if (yi) {
                    on += wv[a];
                    if (dt != nullptr && i + 1 < n &&
                        ((__ldg(bw + ((i + 1) >> 5)) >> ((i + 1) & 31)) & 1u))
                        edge += wv[a];
                }
            }
            dx[row * n + i] = (float)on;
            if (dt != nullptr && seq.L == 0 && i + 1 < n) dt[row * (int64_t)(n - 1) + i] = (float)edge;
        }
        if (dt != nullptr && seq.L > 0) {
            // general chain: d eta_v[e] = sum_a w_a [vertex a runs over edge e].  A thread per edge of the layout
            // (position i, previous state l, state k): both state bits set (start / stop edges: one of them).
            const int* off = seq.lay;
            const int* eb = seq.lay + seq.L + 1;
            for (int e = tid; e < seq.E; e += T) {
                int i = 0;
                while (i < seq.L && __ldg(eb + i + 1) <= e) ++i;
                const int Sc = i < seq.L ? __ldg(off + i + 1) - __ldg(off + i) : 1;
                const int l = (e - __ldg(eb + i)) / Sc, k = (e - __ldg(eb + i)) - l * Sc;
                const int bp = i > 0 ? __ldg(off + i - 1) + l : -1, bc = i < seq.L ? __ldg(off + i) + k : -1;
                double acc = 0.0;
                for (int a = 0; a < m; ++a) {
                    const uint32_t* bw = aset_bits + (row * cap + a) * W;
                    const bool hp = bp < 0 || ((__ldg(bw + (bp >> 5)) >> (bp & 31)) & 1u);
                    const bool hc = bc < 0 || ((__ldg(bw + (bc >> 5)) >> (bc & 31)) & 1u);
                    if (hp && hc) acc += wv[a];
                }
                dt[row * (int64_t)seq.E + e] = (float)acc;
            }
        }
    }
}

// Threads per CTA (= per sample).  Serial phases dominate, so small CTAs: more samples per SM.
static int solver_threads(int n, int kind, int64_t rows = INT64_MAX)
{
    static int ovr = -1;
    if (ovr < 0) {
        const char* e = getenv("SMLVM_SMAP_THREADS");
        ovr = e ? atoi(e) : 0;
    }
    if (ovr >= 32 && ovr <= 256 && ovr % 32 == 0) return ovr;
    // a batch that leaves most SMs idle (the experiments' B = 64) is pure latency: widest CTA
    // (64 samples, n = 32: 0.19 -> 0.15 ms per solve)
    if (rows <= 2 * (int64_t)device_sm_count()) return 256;
    if (n <= 32) return 64;
    if (n <= 64) return 128;
    return kind == SMLVM_FACTOR_BUDGET ? 256 : 128;
}

// Capacity tiers of the forward solver for n bits (ascending, last = n + 1); returns their number.
// First tier: the capacity that lets 5 samples share an SM (4 for the budget factor, whose active
// sets run larger: measured at n = 128, bernoulli q99 = 57 vertices, budget-64 q99 = 83), second
// tier 2 per SM, last tier n + 1.
constexpr int kMaxTiers = 4;
static int plan_tiers(int n, int kind, int (&tiers)[kMaxTiers], size_t extra = 0)
{
    static int single = -1;
    static int forced[3] = {0, 0, 0};
    if (single < 0) {
        const char* e = getenv("SMLVM_SMAP_SINGLE_TIER");
        single = (e && atoi(e)) ? 1 : 0;
        const char* t = getenv("SMLVM_SMAP_TIERS");  // experiments: "56,106" or "54,72,110"
        if (t) sscanf(t, "%d,%d,%d", &forced[0], &forced[1], &forced[2]);
    }
    // largest capacity that still lets `ctas` CTAs share the 227 KB (+1 KB reserved each) of an SM
    auto cap_for = [&](int ctas) {
        const size_t budget_b = (size_t)(227 * 1024) / ctas - 1024;
        int c = n + 1;
        while (c > 8 && solver_smem_bytes(n, c, extra) > budget_b) --c;
        return c;
    };
    int ntier = 0;
    if (!single) {
        int c1 = cap_for(kind == SMLVM_FACTOR_BUDGET ? 4 : 6), c2 = cap_for(2);
        if (forced[0] > 0) c1 = forced[0];
        if (forced[1] > 0) c2 = forced[1];
        if (c1 < n + 1 && (5 * c1 >= 2 * n || forced[0] > 0)) tiers[ntier++] = c1;  // useful only if >= 0.4 n
        if (c2 < n + 1 && (ntier == 0 || c2 > tiers[ntier - 1])) tiers[ntier++] = c2;
        if (forced[2] > 0 && forced[2] < n + 1 && forced[2] > tiers[ntier - 1]) tiers[ntier++] = forced[2];
    }
    tiers[ntier++] = n + 1;
    return ntier;
}

}  // namespace smlvm

using namespace smlvm;

#ifdef SMLVM_SMAP_TIMING
// out_host[16]: cycles per solver phase summed over all CTAs since the last call (then reset)
extern "C" int smlvm_debug_phase_cycles(unsigned long long* out_host)
{
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpyFromSymbol(out_host, g_phase_cycles, sizeof(unsigned long long) * 16);
    if (e != cudaSuccess) return (int)e;
    unsigned long long zero[16] = {0};
    e = cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero));
    return (int)e;
}
#endif

extern "C" int32_t smlvm_sparsemap_capacity(int32_t n) { return n + 1; }

extern "C" int32_t smlvm_sparsemap_plan(int32_t n, int32_t kind, int32_t* caps, int32_t* smem_bytes,
                                        int32_t* ctas_per_sm)
{
    if (n < 1 || n > SMLVM_SPARSEMAP_MAX_BITS) return 0;
    if (kind < SMLVM_FACTOR_BERNOULLI || kind > SMLVM_FACTOR_SEQUENCE_BINARY) return 0;
    int tiers[kMaxTiers];
    const int nt = plan_tiers(n, kind, tiers);
    const int by_regs = 65536 / (64 * solver_threads(n, kind));  // __launch_bounds__(256, 4): <= 64 registers
    for (int i = 0; i < nt; ++i) {
        const size_t b = solver_smem_bytes(n, tiers[i]);
        if (caps) caps[i] = tiers[i];
        if (smem_bytes) smem_bytes[i] = (int32_t)b;
        if (ctas_per_sm) {
            int c = (int)((size_t)(227 * 1024) / (b + 1024));
            c = c < by_regs ? c : by_regs;
            ctas_per_sm[i] = c < 1 ? 1 : (c > 32 ? 32 : c);
        }
    }
    return nt;
}

static size_t ws_list_bytes(int64_t rows) { return ((size_t)rows * sizeof(int32_t) + 63) & ~(size_t)63; }
static size_t ws_rec_bytes(int64_t rows) { return ((size_t)rows * sizeof(long long) + 63) & ~(size_t)63; }

extern "C" size_t smlvm_sparsemap_workspace_bytes(int64_t rows)
{
    // per later tier (at most 3): one counter (64 B slot), a worklist of sample ids and of hand-off record offsets;
    // then the cursor of the record pool.  Whatever the caller passes BEYOND this is the pool (smlvm.h).
    if (rows < 0) return 0;
    return 3 * (64 + ws_list_bytes(rows) + ws_rec_bytes(rows)) + 64;
}

static int sparsemap_fwd_impl(const float* x, const float* t, const double* init_scores,
                              int32_t kind, int32_t budget, int32_t max_iter, int64_t rows,
                              int32_t n, float* p_pad, uint32_t* aset_bits, int32_t* counts,
                              int32_t* kept, int32_t* iters, int32_t* status, double* S_arena,
                              int64_t S_capacity, int64_t* S_offsets, uint64_t* S_cursor,
                              void* workspace, size_t workspace_bytes, void* stream, SeqArgs seq)
{
    if (rows < 0 || n < 1 || max_iter < 0 || budget < 0) return SMLVM_ERR_ARG;
    if (kind < SMLVM_FACTOR_BERNOULLI || kind > SMLVM_FACTOR_SEQUENCE) return SMLVM_ERR_ARG;
    const size_t extra = seq_extra_bytes(seq.L, seq.E);
    if (n > SMLVM_SPARSEMAP_MAX_BITS) return SMLVM_ERR_UNSUPPORTED;
    if (rows == 0) return SMLVM_OK;
    if (x == nullptr || p_pad == nullptr || aset_bits == nullptr || counts == nullptr)
        return SMLVM_ERR_ARG;
    if (S_arena != nullptr && (S_offsets == nullptr || S_cursor == nullptr || S_capacity < 0)) return SMLVM_ERR_ARG;
    if (rows > 0x7fffffff) return SMLVM_ERR_UNSUPPORTED;  // worklists hold int32 sample ids
    if (workspace == nullptr || workspace_bytes < smlvm_sparsemap_workspace_bytes(rows)) return SMLVM_ERR_WORKSPACE;
    const int threads = solver_threads(n, kind, rows);
    // Capacity tiers.  The (|A|+1)^2 fp64 inverse dominates shared memory: sized for the worst case
    // |A| = n + 1 it allows ONE sample per SM at n = 128, although typical active sets hold 40-60
    // vertices.  Samples are first solved with the capacity that lets 6 (budget: 4) CTAs share an SM, then 2
    // (54 / 72 and 110 vertices at n = 128); a sample whose active set outgrows its tier is appended to the next
    // tier's worklist and re-solved from scratch there by a small persistent grid.  Same stream, no host
    // synchronisation; an empty tier costs one launch of CTAs that read a zero and exit.
    int tiers[kMaxTiers];
    const int ntier = plan_tiers(n, kind, tiers, extra);
    if (solver_smem_bytes(n, tiers[0], extra) > (size_t)227 * 1024) return SMLVM_ERR_UNSUPPORTED;
    // first tier: one CTA per sample, the hardware scheduler evens out the data-dependent solve lengths
    // (measured best against persistent grids of 16/64/256 CTAs per SM at n = 32 and n = 128)
    int64_t cap_grid = 0x7fffffff;
    {
        static int gf = -1;
        if (gf < 0) {
            const char* e = getenv("SMLVM_SMAP_GRID_FACTOR");
            gf = e ? atoi(e) : 0;
        }
        if (gf > 0) cap_grid = (int64_t)device_sm_count() * gf;
    }
    const int grid = (int)(rows < cap_grid ? rows : cap_grid);
    cudaStream_t st = as_stream(stream);
    // The later tiers re-solve the few samples that outgrew the first one: most SMs idle, pure latency,
    // so they get the widest CTA like a small batch (SMLVM_SMAP_TAIL_THREADS=0: same CTA as the first tier).
    static int tail_threads = -1;
    if (tail_threads < 0) {
        const char* e = getenv("SMLVM_SMAP_TAIL_THREADS");
        const int v = e ? atoi(e) : 256;
        tail_threads = (v >= 32 && v <= 256 && v % 32 == 0) ? v : 0;
    }
    // workspace: [count_1 (64 B)] [list_1: rows x int32] [count_2 (64 B)] [list_2: rows x int32]
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    const size_t list_bytes = ws_list_bytes(rows), rec_bytes = ws_rec_bytes(rows);
    int32_t* wl_count[kMaxTiers] = {nullptr, nullptr, nullptr, nullptr};
    int32_t* wl_list[kMaxTiers] = {nullptr, nullptr, nullptr, nullptr};
    long long* wl_rec[kMaxTiers] = {nullptr, nullptr, nullptr, nullptr};
    for (int ti = 1; ti < ntier; ++ti) {
        unsigned char* base = ws + (size_t)(ti - 1) * (64 + list_bytes + rec_bytes);
        wl_count[ti] = reinterpret_cast<int32_t*>(base);
        wl_list[ti] = reinterpret_cast<int32_t*>(base + 64);
        wl_rec[ti] = reinterpret_cast<long long*>(base + 64 + list_bytes);
        cudaError_t e = cudaMemsetAsync(wl_count[ti], 0, 64, st);
        if (e != cudaSuccess) return (int)e;
    }
    // hand-off records: the pool is what the workspace holds beyond its mandatory part (none: samples are re-solved)
    const size_t mandatory = smlvm_sparsemap_workspace_bytes(rows);
    TierHandoff ho = {nullptr, 0, reinterpret_cast<unsigned long long*>(ws + mandatory - 64), nullptr, nullptr};
    static int resume = -1;
    if (resume < 0) {
        const char* e = getenv("SMLVM_SMAP_RESUME");
        resume = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (resume && ntier > 1 && workspace_bytes >= mandatory + 4096) {
        ho.pool = ws + mandatory;
        ho.pool_bytes = (long long)((workspace_bytes - mandatory) & ~(size_t)15);
        cudaError_t e = cudaMemsetAsync(ho.cursor, 0, 64, st);
        if (e != cudaSuccess) return (int)e;
    }
#define SMLVM_SMAP_LAUNCH(K)                                                                       \
    do {                                                                                           \
        cudaError_t e =                                                                            \
            cudaFuncSetAttribute(sparsemap_fwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)(solver_smem_bytes(n, n + 1, extra) < (size_t)227 * 1024             \
                                           ? solver_smem_bytes(n, n + 1, extra) : (size_t)227 * 1024));   \
        if (e != cudaSuccess) return (int)e;                                                       \
        e = cudaFuncSetAttribute(sparsemap_fwd_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                 (int)cudaSharedmemCarveoutMaxShared);                             \
        if (e != cudaSuccess) return (int)e;                                                       \
        for (int ti = 0; ti < ntier; ++ti) {                                                       \
            const int th = (ti == 0 || tail_threads == 0) ? threads : tail_threads;                \
            const size_t smem = solver_smem_bytes(n, tiers[ti], extra);                            \
            if (smem > (size_t)227 * 1024) return SMLVM_ERR_UNSUPPORTED;                           \
            int g = grid;                                                                          \
            if (ti > 0) {  /* persistent: as many CTAs as fit the device at this tier's footprint */ \
                int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));                          \
                per_sm = per_sm < 1 ? 1 : per_sm;                                                  \
                const int64_t fit = (int64_t)device_sm_count() * per_sm;                           \
                g = (int)(rows < fit ? rows : fit);                                                \
            }                                                                                      \
            TierHandoff hot = ho;                                                                  \
            hot.rec_in = wl_rec[ti];                                                               \
            hot.rec_out = ti + 1 < ntier ? wl_rec[ti + 1] : nullptr;                               \
            sparsemap_fwd_kernel<K><<<g, th, smem, st>>>(                                          \
                x, t, init_scores, budget, max_iter, rows, n, tiers[ti], ti == ntier - 1,          \
                wl_list[ti], wl_count[ti], ti + 1 < ntier ? wl_list[ti + 1] : nullptr,             \
                ti + 1 < ntier ? wl_count[ti + 1] : nullptr, p_pad, aset_bits, counts, kept, iters, \
                status, S_arena, (long long)S_capacity, reinterpret_cast<long long*>(S_offsets),   \
                reinterpret_cast<unsigned long long*>(S_cursor), seq, hot);                        \
        }                                                                                          \
    } while (0)
    if (kind == SMLVM_FACTOR_BERNOULLI) SMLVM_SMAP_LAUNCH(SMLVM_FACTOR_BERNOULLI);
    else if (kind == SMLVM_FACTOR_BUDGET) SMLVM_SMAP_LAUNCH(SMLVM_FACTOR_BUDGET);
    else if (kind == SMLVM_FACTOR_SEQUENCE) SMLVM_SMAP_LAUNCH(SMLVM_FACTOR_SEQUENCE);
    else SMLVM_SMAP_LAUNCH(SMLVM_FACTOR_SEQUENCE_BINARY);
#undef SMLVM_SMAP_LAUNCH
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_sparsemap_fwd(const float* x, const float* t, const double* init_scores,
                                   int32_t kind, int32_t budget, int32_t max_iter, int64_t rows,
                                   int32_t n, float* p_pad, uint32_t* aset_bits, int32_t* counts,
                                   int32_t* kept, int32_t* iters, int32_t* status, double* S_arena,
                                   int64_t S_capacity, int64_t* S_offsets, uint64_t* S_cursor,
                                   void* workspace, size_t workspace_bytes, void* stream)
{
    if (kind == SMLVM_FACTOR_SEQUENCE) return SMLVM_ERR_ARG;   // the general chain has its own entry point
    SeqArgs none = {nullptr, 0, 0};
    return sparsemap_fwd_impl(x, t, init_scores, kind, budget, max_iter, rows, n, p_pad, aset_bits, counts, kept, iters,
                              status, S_arena, S_capacity, S_offsets, S_cursor, workspace, workspace_bytes, stream, none);
}

extern "C" int smlvm_sparsemap_seq_fwd(const float* x, const float* edge_scores, const int32_t* layout, int32_t length,
                                       int32_t num_edges, int32_t max_iter, int64_t rows, int32_t n, float* p_pad,
                                       uint32_t* aset_bits, int32_t* counts, int32_t* kept, int32_t* iters,
                                       int32_t* status, double* S_arena, int64_t S_capacity, int64_t* S_offsets,
                                       uint64_t* S_cursor, void* workspace, size_t workspace_bytes, void* stream)
{
    if (layout == nullptr || length < 1 || num_edges < 2 || n < length) return SMLVM_ERR_ARG;
    SeqArgs seq = {layout, length, num_edges};
    return sparsemap_fwd_impl(x, edge_scores, nullptr, SMLVM_FACTOR_SEQUENCE, 0, max_iter, rows, n, p_pad, aset_bits, counts,
                              kept, iters, status, S_arena, S_capacity, S_offsets, S_cursor, workspace, workspace_bytes,
                              stream, seq);
}

extern "C" int smlvm_sparsemap_expand(const float* p_pad, const uint32_t* aset_bits,
                                      const int32_t* counts, const int64_t* offsets,
                                      int32_t only_positive, float* p_out, float* aset_out,
                                      int32_t* idx_out, int64_t rows, int32_t n, void* stream)
{
    if (rows < 0 || n < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (p_pad == nullptr || aset_bits == nullptr || counts == nullptr || offsets == nullptr)
        return SMLVM_ERR_ARG;
    int64_t need = (rows + 7) / 8;
    int64_t cap = (int64_t)device_sm_count() * 16;
    sparsemap_expand_kernel<<<(int)(need < cap ? need : cap), 256, 0, as_stream(stream)>>>(
        p_pad, aset_bits, counts, offsets, only_positive, p_out, aset_out, idx_out, rows, n);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

static int sparsemap_bwd_impl(const float* dp, const int64_t* offsets, int32_t only_positive,
                              const float* p_pad, const uint32_t* aset_bits,
                              const int32_t* counts, const double* S_arena, const int64_t* S_offsets,
                              float* dx, float* dt, int64_t rows, int32_t n, void* stream, SeqArgs seq)
{
    if (rows < 0 || n < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (dp == nullptr || offsets == nullptr || p_pad == nullptr || aset_bits == nullptr ||
        counts == nullptr || S_arena == nullptr || S_offsets == nullptr || dx == nullptr)
        return SMLVM_ERR_ARG;
    if (n > SMLVM_SPARSEMAP_MAX_BITS) return SMLVM_ERR_UNSUPPORTED;
    const size_t smem = 2 * (size_t)(n + 1) * sizeof(double);
    int64_t cap = (int64_t)device_sm_count() * 32;
    sparsemap_bwd_kernel<<<(int)(rows < cap ? rows : cap), 128, smem, as_stream(stream)>>>(
        dp, offsets, only_positive, p_pad, aset_bits, counts, S_arena,
        reinterpret_cast<const long long*>(S_offsets), dx, dt, rows, n, seq);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_sparsemap_bwd(const float* dp, const int64_t* offsets, int32_t only_positive,
                                   const float* p_pad, const uint32_t* aset_bits,
                                   const int32_t* counts, const double* S_arena, const int64_t* S_offsets,
                                   float* dx, float* dt, int64_t rows, int32_t n, void* stream)
{
    SeqArgs none = {nullptr, 0, 0};
    return sparsemap_bwd_impl(dp, offsets, only_positive, p_pad, aset_bits, counts, S_arena, S_offsets, dx, dt, rows, n,
                              stream, none);
}

extern "C" int smlvm_sparsemap_seq_bwd(const float* dp, const int64_t* offsets, int32_t only_positive,
                                       const float* p_pad, const uint32_t* aset_bits, const int32_t* counts,
                                       const double* S_arena, const int64_t* S_offsets, const int32_t* layout,
                                       int32_t length, int32_t num_edges, float* dx, float* dedge, int64_t rows,
                                       int32_t n, void* stream)
{
    if (layout == nullptr || length < 1 || num_edges < 2) return SMLVM_ERR_ARG;
    SeqArgs seq = {layout, length, num_edges};
    return sparsemap_bwd_impl(dp, offsets, only_positive, p_pad, aset_bits, counts, S_arena, S_offsets, dx, dedge, rows, n,
                              stream, seq);
}
