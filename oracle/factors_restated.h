// TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// CPU restatement of the reference's factor semantics for n independent bits
// encoded as 2n variables [on_0..on_{n-1}, off_0..off_{n-1}]:
//   RBernoulli       <- lvmhelpers/bernoulli.h:14-130
//   RBudget          <- lvmhelpers/budget.h:18-85
//   RSequenceBinary  <- lvmhelpers/sequence.h:74-243 specialised by
//                       lvmhelpers/sequence_binary.h:10-73
//   RSequence        <- lvmhelpers/sequence.h:25-291, the general chain: num_states[i] states at
//                       position i, one variable per (position, state), one additional score per
//                       edge (start -> first, every transition, last -> stop)
// Written from the behaviour of those files (not copied): same tie rules
// (strict '>'), same accumulation order, same sort key for the budget.
// oracle/ref/Makefile builds the same driver against the reference's own
// headers; tests/test_oracle_ref.py checks both agree.
#pragma once

#include <algorithm>
#include <limits>

#include "ad3/GenericFactor.h"

namespace AD3 {

typedef vector<int> BitCfg;

class RBernoulli : public GenericFactor
{
  public:
    virtual ~RBernoulli() { ClearActiveSet(); }
    void Initialize(int n) { n_ = n; }

    // bernoulli.h:14-35 -- bit on iff on-score strictly beats off-score.
    void Maximize(const vector<double>& eta, const vector<double>&,
                  Configuration& cfg, double* value) override
    {
        BitCfg& y = *static_cast<BitCfg*>(cfg);
        double acc = 0;
        for (int i = 0; i < n_; ++i) {
            bool on = eta[i] > eta[n_ + i];
            y[i] = on ? 1 : 0;
            acc += on ? eta[i] : eta[n_ + i];
        }
        *value = acc;
    }

    // bernoulli.h:38-54
    void Evaluate(const vector<double>& eta, const vector<double>&,
                  const Configuration cfg, double* value) override
    {
        const BitCfg& y = *static_cast<const BitCfg*>(cfg);
        double acc = 0.0;
        for (int i = 0; i < n_; ++i)
            acc += (y[i] == 1) ? eta[i] : eta[n_ + i];
        *value = acc;
    }

    // bernoulli.h:58-73
    void UpdateMarginalsFromConfiguration(const Configuration& cfg, double w,
                                          vector<double>* post,
                                          vector<double>*) override
    {
        const BitCfg& y = *static_cast<const BitCfg*>(cfg);
        for (int i = 0; i < n_; ++i)
            (*post)[(y[i] == 1) ? i : n_ + i] += w;
    }

    // bernoulli.h:76-91 -- Gram entry = n - Hamming distance.
    int CountCommonValues(const Configuration& a,
                          const Configuration& b) override
    {
        const BitCfg& ya = *static_cast<const BitCfg*>(a);
        const BitCfg& yb = *static_cast<const BitCfg*>(b);
        int c = 0;
        for (size_t i = 0; i < ya.size(); ++i) c += (ya[i] == yb[i]);
        return c;
    }

    // bernoulli.h:94-109
    bool SameConfiguration(const Configuration& a,
                           const Configuration& b) override
    {
        return *static_cast<const BitCfg*>(a) == *static_cast<const BitCfg*>(b);
    }

    void DeleteConfiguration(Configuration cfg) override
    {
        delete static_cast<BitCfg*>(cfg);
    }
    Configuration CreateConfiguration() override
    {
        return static_cast<Configuration>(new BitCfg(n_, -1));
    }

  protected:
    int n_ = 0;
};

class RBudget : public RBernoulli
{
  public:
    virtual ~RBudget() { ClearActiveSet(); }
    void Initialize(int n, int budget) { n_ = n; budget_ = budget; }

    // budget.h:24-85 -- all positive-gain bits; if more than `budget`, keep
    // the `budget` largest gains, ties broken towards the lower index
    // (ascending sort of (-gain, index) pairs).
    void Maximize(const vector<double>& eta, const vector<double>&,
                  Configuration& cfg, double* value) override
    {
        BitCfg& y = *static_cast<BitCfg*>(cfg);
        vector<double> gain(n_);
        double base = 0.0;
        for (int i = 0; i < n_; ++i) {
            gain[i] = eta[i] - eta[n_ + i];
            base += eta[n_ + i];
            y[i] = 0;
        }
        size_t active = 0;
        double sum = 0.0;
        for (int i = 0; i < n_; ++i) {
            if (gain[i] > 0) { sum += gain[i]; y[i] = 1; ++active; }
        }
        if (active > static_cast<size_t>(budget_)) {
            vector<pair<double, int>> key(n_);
            for (int i = 0; i < n_; ++i) key[i] = make_pair(-gain[i], i);
            sort(key.begin(), key.end());
            active = 0;
            sum = 0.0;
            for (int k = 0; k < budget_; ++k) {
                double g = -key[k].first;
                if (g < 0) break;
                y[key[k].second] = 1;
                sum += g;
                ++active;
            }
            for (size_t k = active; k < static_cast<size_t>(n_); ++k)
                y[key[k].second] = 0;
        }
        *value = base + sum;
    }

  protected:
    int budget_ = 0;
};

// Two-state linear chain; only the 1->1 transition at position pos (pos>=1)
// carries a score, additional[pos-1] (sequence_binary.h:25-37).
class RSequenceBinary : public RBernoulli
{
  public:
    virtual ~RSequenceBinary() { ClearActiveSet(); }
    void Initialize(int n) { n_ = n; }
    size_t GetNumAdditionals() override { return n_ > 0 ? n_ - 1 : 0; }

    inline double node(const vector<double>& eta, int pos, int s) const
    {
        return s == 1 ? eta[pos] : eta[n_ + pos];
    }
    inline double edge(const vector<double>& add, int pos, int prev, int s) const
    {
        return (prev == 1 && s == 1) ? add[pos - 1] : 0.0;
    }

    // sequence.h:74-151 -- Viterbi; among equal predecessors the lower state
    // wins (`best < 0 || val > best_value`, sequence.h:112,136).
    void Maximize(const vector<double>& eta, const vector<double>& add,
                  Configuration& cfg, double* value) override
    {
        BitCfg& y = *static_cast<BitCfg*>(cfg);
        vector<double> v0(n_), v1(n_);
        vector<int> b0(n_, -1), b1(n_, -1);
        v0[0] = node(eta, 0, 0);  // start edge scores are 0 (prev state 0)
        v1[0] = node(eta, 0, 1);
        for (int i = 0; i + 1 < n_; ++i) {
            for (int k = 0; k < 2; ++k) {
                double best_value = -std::numeric_limits<double>::infinity();
                int best = -1;
                for (int l = 0; l < 2; ++l) {
                    double val = (l ? v1[i] : v0[i]) + edge(add, i + 1, l, k);
                    if (best < 0 || val > best_value) { best_value = val; best = l; }
                }
                double tot = best_value + node(eta, i + 1, k);
                if (k) { v1[i + 1] = tot; b1[i + 1] = best; }
                else   { v0[i + 1] = tot; b0[i + 1] = best; }
            }
        }
        // Termination: final edge (prev, 0) is always 0.
        double best_value = -std::numeric_limits<double>::infinity();
        int best = -1;
        for (int l = 0; l < 2; ++l) {
            double val = (l ? v1[n_ - 1] : v0[n_ - 1]) + 0.0;
            if (best < 0 || val > best_value) { best_value = val; best = l; }
        }
        y[n_ - 1] = best;
        for (int i = n_ - 1; i > 0; --i)
            y[i - 1] = y[i] ? b1[i] : b0[i];
        *value = best_value;
    }

    // sequence.h:154-179
    void Evaluate(const vector<double>& eta, const vector<double>& add,
                  const Configuration cfg, double* value) override
    {
        const BitCfg& y = *static_cast<const BitCfg*>(cfg);
        double acc = 0.0;
        int prev = 0;
        for (int i = 0; i < n_; ++i) {
            acc += node(eta, i, y[i]);
            acc += (i > 0) ? edge(add, i, prev, y[i]) : 0.0;
            prev = y[i];
        }
        *value = acc;
    }

    // sequence.h:183-209 with sequence_binary.h:39-65
    void UpdateMarginalsFromConfiguration(const Configuration& cfg, double w,
                                          vector<double>* post,
                                          vector<double>* add_post) override
    {
        const BitCfg& y = *static_cast<const BitCfg*>(cfg);
        int prev = 0;
        for (int i = 0; i < n_; ++i) {
            (*post)[(y[i] == 1) ? i : n_ + i] += w;
            if (i > 0 && prev == 1 && y[i] == 1) (*add_post)[i - 1] += w;
            prev = y[i];
        }
    }
};

// General linear chain (sequence.h).  Variables: offset[i] + state; additionals indexed as
// Initialize builds them (sequence.h:264-291): position 0 has 1 x S_0 start edges, position i
// S_{i-1} x S_i transitions (row-major in (previous, current)), position L has S_{L-1} x 1 stop edges.
class RSequence : public GenericFactor
{
  public:
    virtual ~RSequence() { ClearActiveSet(); }
    void Initialize(const vector<int>& num_states)
    {
        ns_ = num_states;
        const int L = static_cast<int>(ns_.size());
        off_.assign(L, 0);
        int o = 0;
        for (int i = 0; i < L; ++i) { off_[i] = o; o += ns_[i]; }
        ebase_.assign(L + 2, 0);
        int e = 0;
        for (int i = 0; i <= L; ++i) {
            ebase_[i] = e;
            e += ((i > 0) ? ns_[i - 1] : 1) * ((i < L) ? ns_[i] : 1);
        }
        ebase_[L + 1] = e;
    }
    size_t GetNumAdditionals() override { return ebase_.empty() ? 0 : ebase_.back(); }
    inline double node(const vector<double>& eta, int pos, int s) const { return eta[off_[pos] + s]; }
    inline int eidx(int pos, int prev, int s) const
    {
        const int cur = (pos < static_cast<int>(ns_.size())) ? ns_[pos] : 1;
        return ebase_[pos] + prev * cur + s;
    }

    // sequence.h:74-151 -- Viterbi; among equal predecessors the lower state wins.
    void Maximize(const vector<double>& eta, const vector<double>& add, Configuration& cfg, double* value) override
    {
        BitCfg& y = *static_cast<BitCfg*>(cfg);
        const int L = static_cast<int>(ns_.size());
        vector<vector<double>> val(L);
        vector<vector<int>> path(L);
        val[0].resize(ns_[0]);
        path[0].assign(ns_[0], -1);
        for (int l = 0; l < ns_[0]; ++l) val[0][l] = node(eta, 0, l) + add[eidx(0, 0, l)];
        for (int i = 0; i + 1 < L; ++i) {
            val[i + 1].resize(ns_[i + 1]);
            path[i + 1].resize(ns_[i + 1]);
            for (int k = 0; k < ns_[i + 1]; ++k) {
                double best_value = -std::numeric_limits<double>::infinity();
                int best = -1;
                for (int l = 0; l < ns_[i]; ++l) {
                    const double v = val[i][l] + add[eidx(i + 1, l, k)];
                    if (best < 0 || v > best_value) { best_value = v; best = l; }
                }
                val[i + 1][k] = best_value + node(eta, i + 1, k);
                path[i + 1][k] = best;
            }
        }
        double best_value = -std::numeric_limits<double>::infinity();
        int best = -1;
        for (int l = 0; l < ns_[L - 1]; ++l) {
            const double v = val[L - 1][l] + add[eidx(L, l, 0)];
            if (best < 0 || v > best_value) { best_value = v; best = l; }
        }
        y[L - 1] = best;
        for (int i = L - 1; i > 0; --i) y[i - 1] = path[i][y[i]];
        *value = best_value;
    }

    // sequence.h:154-179: node then edge at every position, the stop edge last
    void Evaluate(const vector<double>& eta, const vector<double>& add, const Configuration cfg, double* value) override
    {
        const BitCfg& y = *static_cast<const BitCfg*>(cfg);
        double acc = 0.0;
        int prev = 0;
        for (size_t i = 0; i < y.size(); ++i) {
            acc += node(eta, static_cast<int>(i), y[i]);
            acc += add[eidx(static_cast<int>(i), prev, y[i])];
            prev = y[i];
        }
        acc += add[eidx(static_cast<int>(y.size()), prev, 0)];
        *value = acc;
    }

    // sequence.h:183-209
    void UpdateMarginalsFromConfiguration(const Configuration& cfg, double w, vector<double>* post,
                                          vector<double>* add_post) override
    {
        const BitCfg& y = *static_cast<const BitCfg*>(cfg);
        int prev = 0;
        for (size_t i = 0; i < y.size(); ++i) {
            (*post)[off_[i] + y[i]] += w;
            (*add_post)[eidx(static_cast<int>(i), prev, y[i])] += w;
            prev = y[i];
        }
        (*add_post)[eidx(static_cast<int>(y.size()), prev, 0)] += w;
    }

    // sequence.h:212-243
    int CountCommonValues(const Configuration& a, const Configuration& b) override
    {
        const BitCfg& ya = *static_cast<const BitCfg*>(a);
        const BitCfg& yb = *static_cast<const BitCfg*>(b);
        int c = 0;
        for (size_t i = 0; i < ya.size(); ++i) c += (ya[i] == yb[i]);
        return c;
    }
    bool SameConfiguration(const Configuration& a, const Configuration& b) override
    {
        return *static_cast<const BitCfg*>(a) == *static_cast<const BitCfg*>(b);
    }
    void DeleteConfiguration(Configuration cfg) override { delete static_cast<BitCfg*>(cfg); }
    Configuration CreateConfiguration() override { return static_cast<Configuration>(new BitCfg(ns_.size(), -1)); }

  protected:
    vector<int> ns_, off_, ebase_;
};

} // namespace AD3
