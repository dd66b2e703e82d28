"""lvmhelpers.sparsemap counterpart: SparseMAP as ``torch.autograd.Function``s.

Reference: lvmhelpers/sparsemap.py.  ``bernoulli_smap`` / ``budget_smap`` /
``sequence_smap`` keep their signatures and return ``(p, aset)`` -- a distribution
over the active vertices (fp32 ``[|A|]``) and the vertices themselves (fp32
``[|A|, n]``, non-differentiable) in the solver's order.  The factor graph, the
libad3 active-set QP and its Jacobian-vector product run in one CTA of the
``smlvm_sparsemap_*`` kernels; there is no per-sample host round trip
(reference: sparsemap.py:21,33-34,42-43).

``init=True`` draws the same ``torch.rand(2n, dtype=double)`` from the CPU
generator as sparsemap.py:26 and seeds the active set with the factor's MAP of
those scores.  (``sequence_smap(init=True)`` raises in the reference,
sparsemap.py:68; here it does what that line intended: zero edge scores.)
"""
import torch

from . import ops
from ._native import FACTOR_BERNOULLI, FACTOR_BUDGET, FACTOR_SEQUENCE_BINARY


def draw_init_scores(rows, n):
    """The uniform draws of sparsemap.py:26, one row per sample, consumed from the
    CPU generator in the same order as the reference's per-sample loop."""
    return torch.rand(rows, 2 * n, dtype=torch.double)


class _SingleSparseMAP(torch.autograd.Function):
    kind = None

    @classmethod
    def _solve(cls, ctx, x, t, budget, max_iter, init):
        if x.dim() != 1:
            raise ValueError("expected a 1-D score vector, got shape %s" % (tuple(x.shape),))
        n = x.shape[0]
        xb = x.detach().reshape(1, n)
        tb = None if t is None else t.detach().reshape(1, n - 1)
        init_scores = draw_init_scores(1, n) if init else None
        state = ops.sparsemap_solve(xb, cls.kind, budget=budget, t=tb, init_scores=init_scores,
                                    max_iter=max_iter, keep_S=True)
        offsets = ops.scan_counts(state["counts"])
        total = int(offsets[-1].item())
        p, aset, _ = ops.sparsemap_expand(state, False, offsets, total, want_idx=False)
        ctx.state, ctx.offsets, ctx.n = state, offsets, n
        ctx.mark_non_differentiable(aset)
        return p, aset

    @classmethod
    def _jv(cls, ctx, dp, want_dt):
        dx, dt = ops.sparsemap_bwd(ctx.state, dp, ctx.offsets, False, want_dt=want_dt)
        return dx.reshape(ctx.n), (None if dt is None else dt.reshape(ctx.n - 1))


class BernSparseMAP(_SingleSparseMAP):
    """reference sparsemap.py:90-106"""
    kind = FACTOR_BERNOULLI

    @classmethod
    def forward(cls, ctx, x, max_iter, init):
        return cls._solve(ctx, x, None, 0, max_iter, init)

    @classmethod
    def backward(cls, ctx, dp, daset):
        return cls._jv(ctx, dp, False)[0], None, None


class BudgetSparseMAP(_SingleSparseMAP):
    """reference sparsemap.py:109-127"""
    kind = FACTOR_BUDGET

    @classmethod
    def forward(cls, ctx, x, budget, max_iter, init):
        return cls._solve(ctx, x, None, budget, max_iter, init)

    @classmethod
    def backward(cls, ctx, dp, daset):
        return cls._jv(ctx, dp, False)[0], None, None, None


class SequenceBinarySparseMAP(_SingleSparseMAP):
    """reference sparsemap.py:130-148"""
    kind = FACTOR_SEQUENCE_BINARY

    @classmethod
    def forward(cls, ctx, x, t, max_iter, init):
        return cls._solve(ctx, x, t, 0, max_iter, init)

    @classmethod
    def backward(cls, ctx, dp, daset):
        dx, dt = cls._jv(ctx, dp, True)
        return dx, dt, None, None


def bernoulli_smap(x, max_iter=100, init=True):
    return BernSparseMAP.apply(x, max_iter, init)


def budget_smap(x, budget=5, max_iter=100, init=True):
    return BudgetSparseMAP.apply(x, budget, max_iter, init)


def sequence_smap(x, t, max_iter=100, init=True):
    return SequenceBinarySparseMAP.apply(x, t, max_iter, init)
