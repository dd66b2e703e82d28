// TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// Restatement of AD3's active-set QP solver `GenericFactor::SolveQP` and its
// helpers (third-party libad3 / LP-SparseMAP, absent from /root/reference;
// call sites lvmhelpers/sparsemap.py:27,29-30,42). PARITY UNPINNED -- see the
// header. Constants restated from the public source: convergence slack 1e-9,
// singularity test |s| < 1e-9, seed inverse [[-n, 1], [1, 0]].
//
// Deliberate deviations from libad3 (both are dead ends there):
//  * "vertex already active" branch: libad3 clears its cached active set and
//    returns, so `get_sparse_solution()` would be empty and the reference
//    would trip `assert supp > 0` (structmarg.py:192). We keep the current
//    (p, aset) and flag SOLVE_REPEATED_CONFIG.
//  * singular insertion: libad3 runs an Eigen eigendecomposition to evict a
//    vertex from the null space. We stop with the current (p, aset) and flag
//    SOLVE_SINGULAR. Tests count how often either flag fires.
#include "ad3/GenericFactor.h"

namespace AD3 {

void GenericFactor::ClearActiveSet()
{
    for (size_t j = 0; j < active_set_.size(); ++j)
        DeleteConfiguration(active_set_[j]);
    active_set_.clear();
    distribution_.clear();
    inverse_A_.clear();
}

void GenericFactor::SeedActiveSet(Configuration configuration)
{
    active_set_.push_back(configuration);
    distribution_.push_back(1.0);
    // inv([[0, 1], [1, c]]) = [[-c, 1], [1, 0]],  c = <m_y, m_y>.
    inverse_A_.resize(4);
    inverse_A_[0] = -static_cast<double>(
      CountCommonValues(configuration, configuration));
    inverse_A_[1] = 1;
    inverse_A_[2] = 1;
    inverse_A_[3] = 0;
}

void GenericFactor::InitActiveSetFromScores(const vector<double>& eta_u,
                                            const vector<double>& eta_v)
{
    ClearActiveSet();
    Configuration configuration = CreateConfiguration();
    double value;
    Maximize(eta_u, eta_v, configuration, &value);
    SeedActiveSet(configuration);
}

void GenericFactor::ComputeMarginalsFromSparseDistribution(
  const vector<Configuration>& active_set,
  const vector<double>& distribution,
  vector<double>* variable_posteriors,
  vector<double>* additional_posteriors)
{
    variable_posteriors->assign(variable_posteriors->size(), 0.0);
    additional_posteriors->assign(additional_posteriors->size(), 0.0);
    for (size_t i = 0; i < active_set.size(); ++i) {
        UpdateMarginalsFromConfiguration(active_set[i],
                                         distribution[i],
                                         variable_posteriors,
                                         additional_posteriors);
    }
}

// Block update of inv(A) when a row/column r is appended to A.
bool GenericFactor::InvertAfterInsertion(
  const vector<Configuration>& active_set,
  const Configuration& inserted_element)
{
    vector<double> inverse_A = inverse_A_;
    int size_A = static_cast<int>(active_set.size()) + 1;
    vector<double> r(size_A);

    r[0] = 1.0;
    for (size_t i = 0; i < active_set.size(); ++i) {
        r[i + 1] = static_cast<double>(
          CountCommonValues(active_set[i], inserted_element));
    }
    double r0 = static_cast<double>(
      CountCommonValues(inserted_element, inserted_element));

    // Schur complement s = r0 - r' inv(A) r (upper triangle, doubled).
    double s = r0;
    for (int i = 0; i < size_A; ++i) {
        if (r[i] == 0.0) continue;
        s -= r[i] * r[i] * inverse_A[i * size_A + i];
        for (int j = i + 1; j < size_A; ++j) {
            if (r[j] == 0.0) continue;
            s -= 2 * r[i] * r[j] * inverse_A[i * size_A + j];
        }
    }
    if (s > -1e-9 && s < 1e-9) return false;  // would become singular

    double invs = 1.0 / s;
    vector<double> d(size_A, 0.0);
    for (int i = 0; i < size_A; ++i) {
        if (r[i] == 0.0) continue;
        for (int j = 0; j < size_A; ++j)
            d[j] += inverse_A[i * size_A + j] * r[i];
    }

    int size_A_after = size_A + 1;
    inverse_A_.resize(size_A_after * size_A_after);
    for (int i = 0; i < size_A; ++i) {
        for (int j = 0; j < size_A; ++j) {
            inverse_A_[i * size_A_after + j] =
              inverse_A[i * size_A + j] + invs * d[i] * d[j];
        }
        inverse_A_[i * size_A_after + size_A] = -invs * d[i];
        inverse_A_[size_A * size_A_after + i] = -invs * d[i];
    }
    inverse_A_[size_A * size_A_after + size_A] = invs;
    return true;
}

// Block downdate of inv(A) when row/column `removed_index` leaves A.
void GenericFactor::InvertAfterRemoval(const vector<Configuration>& active_set,
                                       int removed_index)
{
    vector<double> inverse_A = inverse_A_;
    int size_A = static_cast<int>(active_set.size()) + 1;

    ++removed_index;  // index in A is offset by the constraint row
    double invs = inverse_A[removed_index * size_A + removed_index];
    double s = 1.0 / invs;
    vector<double> d(size_A - 1, 0.0);
    int k = 0;
    for (int i = 0; i < size_A; ++i) {
        if (i == removed_index) continue;
        d[k] = -s * inverse_A[removed_index * size_A + i];
        ++k;
    }

    int size_A_after = size_A - 1;
    inverse_A_.resize(size_A_after * size_A_after);
    k = 0;
    for (int i = 0; i < size_A; ++i) {
        if (i == removed_index) continue;
        int l = 0;
        for (int j = 0; j < size_A; ++j) {
            if (j == removed_index) continue;
            inverse_A_[k * size_A_after + l] =
              inverse_A[i * size_A + j] - invs * d[k] * d[l];
            ++l;
        }
        ++k;
    }
}

void GenericFactor::SolveQP(const vector<double>& variable_log_potentials,
                            const vector<double>& additional_log_potentials,
                            vector<double>* variable_posteriors,
                            vector<double>* additional_posteriors)
{
    variable_posteriors->resize(variable_log_potentials.size());
    additional_posteriors->resize(additional_log_potentials.size());
    last_status_ = SOLVE_MAX_ITER;
    last_iterations_ = 0;

    // Seed the active set with the LP solution (quadratic term dropped).
    if (active_set_.size() == 0) {
        distribution_.clear();
        Configuration configuration = CreateConfiguration();
        double value;
        Maximize(variable_log_potentials,
                 additional_log_potentials,
                 configuration,
                 &value);
        SeedActiveSet(configuration);
    }

    bool changed_active_set = true;
    vector<double> z;
    double tau = 0;
    for (int iter = 0; iter < num_max_iterations_QP_; ++iter) {
        last_iterations_ = iter + 1;
        bool same_as_before = true;
        if (changed_active_set) {
            // b = [1, score(y_1), ..., score(y_|A|)];  [tau; z] = inv(A) b.
            vector<double> b(active_set_.size() + 1, 0.0);
            b[0] = 1.0;
            for (size_t i = 0; i < active_set_.size(); ++i) {
                double score;
                Evaluate(variable_log_potentials,
                         additional_log_potentials,
                         active_set_[i],
                         &score);
                b[i + 1] = score;
            }
            z.resize(active_set_.size());
            int size_A = static_cast<int>(active_set_.size()) + 1;
            for (size_t i = 0; i < active_set_.size(); ++i) {
                z[i] = 0.0;
                for (int j = 0; j < size_A; ++j)
                    z[i] += inverse_A_[(i + 1) * size_A + j] * b[j];
            }
            tau = 0.0;
            for (int j = 0; j < size_A; ++j)
                tau += inverse_A_[j] * b[j];
            same_as_before = false;
        }

        if (same_as_before) {
            // u = M z; most violated constraint = MAP of (eta_u - u, eta_v).
            ComputeMarginalsFromSparseDistribution(
              active_set_, z, variable_posteriors, additional_posteriors);
            vector<double> scores = variable_log_potentials;
            for (size_t i = 0; i < scores.size(); ++i)
                scores[i] -= (*variable_posteriors)[i];
            Configuration configuration = CreateConfiguration();
            double value = 0.0;
            Maximize(scores, additional_log_potentials, configuration, &value);

            const double very_small_threshold = 1e-9;
            if (value <= tau + very_small_threshold) {
                DeleteConfiguration(configuration);
                last_status_ = SOLVE_CONVERGED;
                return;
            }
            for (size_t k = 0; k < active_set_.size(); ++k) {
                if (SameConfiguration(active_set_[k], configuration)) {
                    DeleteConfiguration(configuration);
                    last_status_ = SOLVE_REPEATED_CONFIG;  // see file header
                    return;
                }
            }
            if (!InvertAfterInsertion(active_set_, configuration)) {
                DeleteConfiguration(configuration);
                last_status_ = SOLVE_SINGULAR;  // see file header
                return;
            }
            z.push_back(0.0);
            distribution_.push_back(0.0);
            active_set_.push_back(configuration);
            changed_active_set = true;
        } else {
            // Line search towards z; find the first blocking constraint.
            int blocking = -1;
            bool exist_blocking = false;
            double alpha = 1.0;
            for (size_t i = 0; i < active_set_.size(); ++i) {
                if (z[i] >= distribution_[i]) continue;
                if (z[i] < 0) exist_blocking = true;
                double tmp = distribution_[i] / (distribution_[i] - z[i]);
                if (blocking < 0 || tmp < alpha) {
                    alpha = tmp;
                    blocking = static_cast<int>(i);
                }
            }

            if (!exist_blocking) {
                distribution_ = z;
                changed_active_set = false;
            } else {
                if (alpha > 1.0) alpha = 1.0;
                if (alpha == 1.0) {
                    distribution_ = z;
                } else {
                    for (size_t i = 0; i < active_set_.size(); ++i) {
                        z[i] = (1 - alpha) * distribution_[i] + alpha * z[i];
                        distribution_[i] = z[i];
                    }
                }
                DeleteConfiguration(active_set_[blocking]);
                InvertAfterRemoval(active_set_, blocking);
                active_set_.erase(active_set_.begin() + blocking);
                z.erase(z.begin() + blocking);
                distribution_.erase(distribution_.begin() + blocking);
                changed_active_set = true;
            }
        }
    }
    // Maximum number of iterations: marginals of the current z.
    ComputeMarginalsFromSparseDistribution(
      active_set_, z, variable_posteriors, additional_posteriors);
}

void GenericFactor::DistJacobianVec(const vector<double>& dp,
                                    vector<double>* out_u,
                                    vector<double>* out_v)
{
    int m = static_cast<int>(active_set_.size());
    int size_A = m + 1;
    vector<double> w(m, 0.0);
    for (int i = 0; i < m; ++i) {
        double acc = 0.0;
        for (int j = 0; j < m; ++j)
            acc += inverse_A_[(i + 1) * size_A + (j + 1)] * dp[j];
        w[i] = acc;
    }
    out_u->assign(out_u->size(), 0.0);
    out_v->assign(out_v->size(), 0.0);
    for (int i = 0; i < m; ++i)
        UpdateMarginalsFromConfiguration(active_set_[i], w[i], out_u, out_v);
}

} // namespace AD3
