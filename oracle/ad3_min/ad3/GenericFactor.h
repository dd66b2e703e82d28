// TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// Minimal restatement of the AD3 / LP-SparseMAP `GenericFactor` base class
// (third-party `lpsmap`, libad3 -- NOT present under /root/reference and not
// pinned by it: the reference links it as `-lad3`, setup.py:95-109, and calls
// it at lvmhelpers/sparsemap.py:15-20,27,29-30,42).
//
// PARITY UNPINNED: the real libad3 could not be opened in the build container.
// What follows restates the published AD3 2.x active-set algorithm
// (Martins et al., "AD3: Alternating Directions Dual Decomposition for MAP
// Inference in Graphical Models", JMLR 2015, Alg. 3 + `GenericFactor.cpp`)
// from knowledge of the public source. The class/virtual-method surface is
// exactly what the reference's in-tree factor headers need
// (lvmhelpers/bernoulli.h:8-137, budget.h:11-88, sequence.h:26-304,
// sequence_binary.h:7-80), so those headers compile unmodified against it
// (see oracle/ref/Makefile) and drive the same solver as our own restated
// factors (oracle/factors_restated.h).
#pragma once

#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <utility>
#include <vector>

using namespace std;  // the AD3 headers export std this way; the reference's
                      // factor headers rely on unqualified vector/pair/sort.

namespace AD3 {

typedef void* Configuration;

// Outcome flags of the restated SolveQP (not in AD3; test bookkeeping).
enum SolveStatus {
    SOLVE_CONVERGED = 0,      // value <= tau + 1e-9
    SOLVE_MAX_ITER = 1,       // loop ran out
    SOLVE_REPEATED_CONFIG = 2,// MAP oracle returned a vertex already active
    SOLVE_SINGULAR = 3        // Schur complement |s| < 1e-9 on insertion
};

class GenericFactor
{
  public:
    GenericFactor()
      : verbosity_(0)
      , num_max_iterations_QP_(10)
      , last_status_(SOLVE_CONVERGED)
      , last_iterations_(0)
    {}
    virtual ~GenericFactor() {}

    // ---- interface implemented by concrete factors -------------------------
    virtual void Maximize(const vector<double>& variable_log_potentials,
                          const vector<double>& additional_log_potentials,
                          Configuration& configuration,
                          double* value) = 0;
    virtual void Evaluate(const vector<double>& variable_log_potentials,
                          const vector<double>& additional_log_potentials,
                          const Configuration configuration,
                          double* value) = 0;
    virtual void UpdateMarginalsFromConfiguration(
      const Configuration& configuration,
      double weight,
      vector<double>* variable_posteriors,
      vector<double>* additional_posteriors) = 0;
    virtual int CountCommonValues(const Configuration& configuration1,
                                  const Configuration& configuration2) = 0;
    virtual bool SameConfiguration(const Configuration& configuration1,
                                   const Configuration& configuration2) = 0;
    virtual void DeleteConfiguration(Configuration configuration) = 0;
    virtual Configuration CreateConfiguration() = 0;
    virtual size_t GetNumAdditionals() { return 0; }

    // ---- solver state / API used through the Cython wrapper ---------------
    void SetVerbosity(int v) { verbosity_ = v; }
    void SetQPMaxIter(int it) { num_max_iterations_QP_ = it; }
    void SetNumVariables(int nv) { num_variables_ = nv; }

    void ClearActiveSet();

    // `init_active_set_from_scores` (sparsemap.py:27): the active set becomes
    // the single MAP vertex of the given scores, with weight 1.
    void InitActiveSetFromScores(const vector<double>& eta_u,
                                 const vector<double>& eta_v);

    // `solve_qp` (sparsemap.py:29).
    void SolveQP(const vector<double>& variable_log_potentials,
                 const vector<double>& additional_log_potentials,
                 vector<double>* variable_posteriors,
                 vector<double>* additional_posteriors);

    // `get_sparse_solution` (sparsemap.py:30).
    const vector<Configuration>& GetQPActiveSet() const { return active_set_; }
    const vector<double>& GetQPDistribution() const { return distribution_; }
    const vector<double>& GetQPInvA() const { return inverse_A_; }

    // `dist_jacobian_vec` (sparsemap.py:42): out_u = M S dp, out_v = N S dp,
    // S = inverse_A_[1:,1:].
    void DistJacobianVec(const vector<double>& dp,
                         vector<double>* out_u,
                         vector<double>* out_v);

    int last_status() const { return last_status_; }
    int last_iterations() const { return last_iterations_; }

  protected:
    void ComputeMarginalsFromSparseDistribution(
      const vector<Configuration>& active_set,
      const vector<double>& distribution,
      vector<double>* variable_posteriors,
      vector<double>* additional_posteriors);
    bool InvertAfterInsertion(const vector<Configuration>& active_set,
                              const Configuration& inserted_element);
    void InvertAfterRemoval(const vector<Configuration>& active_set,
                            int removed_index);
    void SeedActiveSet(Configuration configuration);

  protected:
    int verbosity_;
    int num_max_iterations_QP_;
    int num_variables_ = 0;
    vector<Configuration> active_set_;
    vector<double> distribution_;
    vector<double> inverse_A_;
    int last_status_;
    int last_iterations_;
};

} // namespace AD3
