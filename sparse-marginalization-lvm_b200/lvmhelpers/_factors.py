"""Factor handles of the SparseMAP solver.

The reference's ``PFactor*`` classes (lvmhelpers/pbernoulli.pyx:10-22, pbudget.pyx:10-22,
psequence.pyx:10-37) are Cython shims that own a libad3 ``GenericFactor`` (the MAP oracle of
bernoulli.h / budget.h / sequence.h / sequence_binary.h plus the active-set state).  Here the
oracles are block primitives INSIDE the solver kernel (csrc/sparsemap.cu), so a handle only
records which oracle to run and its parameters; ``lvmhelpers.sparsemap`` reads them when it
launches ``smlvm_sparsemap_fwd``.  No native object outlives a call.
"""
from ._native import FACTOR_BERNOULLI, FACTOR_BUDGET, FACTOR_SEQUENCE_BINARY, FACTOR_SEQUENCE


class PGenericFactor:
    """Common part of the handles: ``kind`` (SMLVM_FACTOR_*), ``length`` and ``budget``."""
    kind = None

    def __init__(self, allocate=True):
        self.allocate = allocate
        self.length = None
        self.budget = 0
        self.num_states = None

    def _require_initialized(self):
        if self.length is None:
            raise RuntimeError("%s: call initialize(...) first" % type(self).__name__)

    def __repr__(self):
        return "%s(length=%r, budget=%r)" % (type(self).__name__, self.length, self.budget)


class PFactorBernoulli(PGenericFactor):
    """n independent bits; MAP = [eta_on > eta_off] (bernoulli.h:14-35)."""
    kind = FACTOR_BERNOULLI

    def initialize(self, length):
        self.length = int(length)


class PFactorBudget(PGenericFactor):
    """n bits with at most ``budget`` of them on (budget.h:24-85)."""
    kind = FACTOR_BUDGET

    def initialize(self, length, budget):
        if int(budget) < 0:
            raise ValueError("budget must be >= 0")
        self.length, self.budget = int(length), int(budget)


class PFactorSequenceBinary(PGenericFactor):
    """2-state chain, edge score only on 1 -> 1 (sequence_binary.h:10-37)."""
    kind = FACTOR_SEQUENCE_BINARY

    def initialize(self, length):
        self.length = int(length)


class PFactorSequence(PGenericFactor):
    """Chain with ``num_states[i]`` states at position i and a score on every edge, start and stop
    included (sequence.h:74-151, 264-291)."""
    kind = FACTOR_SEQUENCE

    def initialize(self, num_states):
        num_states = [int(s) for s in num_states]
        if not num_states or min(num_states) < 1:
            raise ValueError("num_states must be a non-empty list of positive ints")
        self.num_states = num_states
        self.length = len(num_states)

    @property
    def num_variables(self):
        """Total number of (position, state) variables = sum(num_states)."""
        self._require_initialized()
        return sum(self.num_states)

    @property
    def num_additionals(self):
        """Number of edge scores in sequence.h's order: start -> s0, s_{i-1} -> s_i, s_last -> stop
        (sequence.h:274-289)."""
        self._require_initialized()
        ns = self.num_states
        return ns[0] + sum(a * b for a, b in zip(ns[:-1], ns[1:])) + ns[-1]
