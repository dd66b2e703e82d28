"""lvmhelpers.sparsemap counterpart: SparseMAP as ``torch.autograd.Function``s.

Reference: lvmhelpers/sparsemap.py.  Same class tree -- ``SparseMAP`` / ``SequenceSparseMAP`` with
``run_sparsemap`` / ``jv`` and a ``make_factor`` hook per subclass (sparsemap.py:10-87), the three
concrete functions (:90-148) and ``bernoulli_smap`` / ``budget_smap`` / ``sequence_smap`` (:151-160),
which return ``(p, aset)``: a distribution over the active vertices (fp32 ``[|A|]``) and the
vertices themselves (fp32 ``[|A|, n]``, non-differentiable) in the solver's order.

What changed underneath: ``make_factor`` returns a plain handle (``lvmhelpers._factors``) that
names the MAP oracle; the factor graph, the libad3 active-set QP and its Jacobian-vector product
run in one CTA of the ``smlvm_sparsemap_*`` kernels, and there is no per-sample host round trip
(reference: sparsemap.py:21,33-34,42-43).

``init=True`` draws the same ``torch.rand(2n, dtype=double)`` from the CPU generator as
sparsemap.py:26 and seeds the active set with the factor's MAP of those scores.
(``sequence_smap(init=True)`` raises in the reference, sparsemap.py:68; here it does what that
line intended: zero edge scores.)
"""
import torch

from . import ops
from .pbernoulli import PFactorBernoulli
from .pbudget import PFactorBudget
from .psequence import PFactorSequence, PFactorSequenceBinary


def draw_init_scores(rows, n):
    """The uniform draws of sparsemap.py:26, one row per sample, consumed from the
    CPU generator in the same order as the reference's per-sample loop."""
    return torch.rand(rows, 2 * n, dtype=torch.double)


def _solve_one(ctx, x, t):
    """One sample through the batched solver (rows = 1) with the factor ``ctx.f``."""
    if x.dim() != 1:
        raise ValueError("expected a 1-D score vector, got shape %s" % (tuple(x.shape),))
    f = ctx.f
    f._require_initialized()
    n = x.shape[0]
    if f.length != n:
        raise ValueError("factor initialised for %d variables, scores have %d" % (f.length, n))
    xb = x.detach().reshape(1, n)
    tb = None if t is None else t.detach().reshape(1, n - 1)
    init_scores = draw_init_scores(1, n) if ctx.init else None
    state = ops.sparsemap_solve(xb, f.kind, budget=f.budget, t=tb, init_scores=init_scores,
                                max_iter=ctx.max_iter, keep_S=True)
    offsets = ops.scan_counts(state["counts"])
    total = int(offsets[-1].item())
    p, aset, _ = ops.sparsemap_expand(state, False, offsets, total, want_idx=False)
    ctx.state, ctx.offsets = state, offsets
    ctx.mark_non_differentiable(aset)
    return p, aset


class SparseMAP(torch.autograd.Function):
    """reference sparsemap.py:10-44: factors over 2n binary variables (on / off halves)."""

    @classmethod
    def make_factor(cls, ctx):
        raise NotImplementedError

    @classmethod
    def run_sparsemap(cls, ctx, x):
        ctx.n = x.shape[0]
        ctx.f = cls.make_factor(ctx)
        return _solve_one(ctx, x, None)

    @classmethod
    def jv(cls, ctx, dp):
        dx, _ = ops.sparsemap_bwd(ctx.state, dp, ctx.offsets, False, want_dt=False)
        return dx.reshape(ctx.n)


class SequenceSparseMAP(torch.autograd.Function):
    """reference sparsemap.py:47-87: the same with edge potentials ``t`` as additional scores."""

    @classmethod
    def make_factor(cls, ctx):
        raise NotImplementedError

    @classmethod
    def run_sparsemap(cls, ctx, x, t):
        ctx.n = x.shape[0]
        ctx.f = cls.make_factor(ctx)
        return _solve_one(ctx, x, t)

    @classmethod
    def jv(cls, ctx, dp):
        dx, dt = ops.sparsemap_bwd(ctx.state, dp, ctx.offsets, False, want_dt=True)
        return dx.reshape(ctx.n), dt.reshape(ctx.n - 1)


class BernSparseMAP(SparseMAP):
    """reference sparsemap.py:90-106"""

    @classmethod
    def make_factor(cls, ctx):
        f = PFactorBernoulli()
        f.initialize(ctx.n)
        return f

    @classmethod
    def forward(cls, ctx, x, max_iter, init):
        ctx.max_iter = max_iter
        ctx.init = init
        return cls.run_sparsemap(ctx, x)

    @classmethod
    def backward(cls, ctx, dp, daset):
        return cls.jv(ctx, dp), None, None


class BudgetSparseMAP(SparseMAP):
    """reference sparsemap.py:109-127"""

    @classmethod
    def make_factor(cls, ctx):
        f = PFactorBudget()
        f.initialize(ctx.n, ctx.budget)
        return f

    @classmethod
    def forward(cls, ctx, x, budget, max_iter, init):
        ctx.n = x.shape[0]
        ctx.init = init
        ctx.max_iter = max_iter
        ctx.budget = budget
        return cls.run_sparsemap(ctx, x)

    @classmethod
    def backward(cls, ctx, dp, daset):
        return cls.jv(ctx, dp), None, None, None


class SequenceBinarySparseMAP(SequenceSparseMAP):
    """reference sparsemap.py:130-148"""

    @classmethod
    def make_factor(cls, ctx):
        f = PFactorSequenceBinary()
        f.initialize(ctx.n)
        return f

    @classmethod
    def forward(cls, ctx, x, t, max_iter, init):
        ctx.n = x.shape[0]
        ctx.init = init
        ctx.max_iter = max_iter
        return cls.run_sparsemap(ctx, x, t)

    @classmethod
    def backward(cls, ctx, dp, daset):
        d_eta_u, d_eta_v = cls.jv(ctx, dp)
        return d_eta_u, d_eta_v, None, None


class GeneralSequenceSparseMAP(SequenceSparseMAP):
    """SparseMAP over ``PFactorSequence`` (lvmhelpers/psequence.pyx:10-22, sequence.h:25-291): a chain whose position
    i has ``num_states[i]`` states, node scores ``x [sum num_states]`` and one additional score per edge, ``t [E]``
    (start, transitions, stop in the order of sequence.h:264-291).  The reference binds the factor but has no
    autograd function over it; this one follows ``SequenceSparseMAP`` (sparsemap.py:47-87): returns ``(p [|A|],
    aset [|A|, L])`` with the state sequences as the configurations ``get_sparse_solution`` returns, and the
    backward gives ``(d x, d t)``.  ``init=False`` only (the reference's init path raises for sequences, :68)."""

    @classmethod
    def make_factor(cls, ctx):
        f = PFactorSequence()
        f.initialize(ctx.num_states)
        return f

    @classmethod
    def forward(cls, ctx, x, t, num_states, max_iter):
        ctx.num_states = [int(s) for s in num_states]
        ctx.f = cls.make_factor(ctx)
        if x.dim() != 1 or x.shape[0] != ctx.f.num_variables:
            raise ValueError("expected %d node scores, got shape %s" % (ctx.f.num_variables, tuple(x.shape)))
        if t.dim() != 1 or t.shape[0] != ctx.f.num_additionals:
            raise ValueError("expected %d edge scores, got shape %s" % (ctx.f.num_additionals, tuple(t.shape)))
        state = ops.sparsemap_solve_sequence(x.detach().reshape(1, -1), t.detach().reshape(1, -1), ctx.num_states,
                                             max_iter=max_iter)
        offsets = ops.scan_counts(state["counts"])
        total = int(offsets[-1].item())
        p, onehot, _ = ops.sparsemap_expand(state, False, offsets, total, want_idx=False)
        aset = ops.onehot_to_states(onehot, ctx.num_states)
        ctx.state, ctx.offsets = state, offsets
        ctx.mark_non_differentiable(aset)
        return p, aset

    @classmethod
    def backward(cls, ctx, dp, daset):
        dx, dt = ops.sparsemap_bwd_sequence(ctx.state, dp, ctx.offsets, False)
        return dx.reshape(-1), dt.reshape(-1), None, None


def general_sequence_smap(x, t, num_states, max_iter=100):
    """SparseMAP over the general chain factor (``PFactorSequence``): see ``GeneralSequenceSparseMAP``."""
    return GeneralSequenceSparseMAP.apply(x, t, num_states, max_iter)


def bernoulli_smap(x, max_iter=100, init=True):
    return BernSparseMAP.apply(x, max_iter, init)


def budget_smap(x, budget=5, max_iter=100, init=True):
    return BudgetSparseMAP.apply(x, budget, max_iter, init)


def sequence_smap(x, t, max_iter=100, init=True):
    return SequenceBinarySparseMAP.apply(x, t, max_iter, init)
