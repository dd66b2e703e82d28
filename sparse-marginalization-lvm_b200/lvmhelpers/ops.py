"""Torch-facing operators over the C-ABI kernels (``include/smlvm.h``).

Raw launches (``*_raw`` / plain functions) return tensors and never synchronise;
``torch.autograd.Function`` subclasses own the forward/backward pairing.  All
tensors must be CUDA fp32 contiguous; inputs are made contiguous, never copied to
the CPU.  The only host round trip is the one read of the ragged total in
``compact`` / ``sparsemap_solve`` (the reference builds Python lists there,
lvmhelpers/structmarg.py:195-202).
"""
import torch

from . import _native as N
from ._native import lib, check, ptr, stream


SLOTS_MIN_ROWS = 4096
SLOTS_MIN_COLS = 16


def _slots_or_none(slots):
    return slots if (slots is not None and slots.numel() > 0) else None


def _f32c(t, name):
    N.require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# --------------------------------------------------------------------------------------
# (1) sparsemax
# --------------------------------------------------------------------------------------
def sparsemax_fwd(x, want_entropy=False, want_argmax=False, want_nnz=True, want_slots=False):
    """x [rows, K] -> dict(p, supp [int32], nnz [int32]|None, entropy|None, argmax [int64]|None,
    slots [rows, 8, 2] fp32 | None  (support (value, column) pairs for ``compact``)).

    entmax.sparsemax forward (lvmhelpers/marg.py:51) fused with entropy() (marg.py:6-28)
    and scores.argmax(-1) (marg.py:53).
    """
    x = _f32c(x, "scores")
    rows, K = x.shape
    p = torch.empty_like(x)
    supp = torch.empty(rows, dtype=torch.int32, device=x.device)
    nnz = torch.empty(rows, dtype=torch.int32, device=x.device) if want_nnz else None
    ent = torch.empty(rows, dtype=torch.float32, device=x.device) if want_entropy else None
    amax = torch.empty(rows, dtype=torch.int64, device=x.device) if want_argmax else None
    slots = torch.empty((rows, N.SUPPORT_SLOTS, 2), dtype=torch.float32, device=x.device) if want_slots else None
    check(lib.smlvm_sparsemax_fwd(ptr(x), ptr(p), ptr(supp), ptr(nnz), ptr(ent), ptr(amax), ptr(slots),
                                  rows, K, stream()), "smlvm_sparsemax_fwd")
    return dict(p=p, supp=supp, nnz=nnz, entropy=ent, argmax=amax, slots=slots)


def sparsemax_bwd(p, supp, dp=None, loss=None, loss_scale=None, loss_scale_host=1.0, dent=None,
                  want_dloss=False, out=None, out_dloss=None, slots=None, loss_slots=None):
    """Fused Jacobian-vector product; see smlvm_sparsemax_bwd in include/smlvm.h.
    With want_dloss also returns dloss = loss_scale * p (decoder side of the weighted loss).
    ``slots``: the forward's support pairs for this ``p`` (only the support entries of dp / loss are
    then read)."""
    rows, K = p.shape
    dx = torch.empty_like(p) if out is None else out
    dloss = None
    if want_dloss:
        dloss = torch.empty_like(p) if out_dloss is None else out_dloss
    dp = None if dp is None else _f32c(dp, "dp")
    loss = None if loss is None else _f32c(loss, "loss")
    dent = None if dent is None else _f32c(dent, "dent")
    check(lib.smlvm_sparsemax_bwd(ptr(p), ptr(supp), ptr(dp), ptr(loss), ptr(loss_scale),
                                  float(loss_scale_host), ptr(dent), ptr(_slots_or_none(slots)), ptr(_slots_or_none(loss_slots)),
                                  ptr(dx), ptr(dloss), rows, K,
                                  stream()), "smlvm_sparsemax_bwd")
    return (dx, dloss) if want_dloss else dx


def sparsemax_bwd_ragged(comp, supp, shape, dp=None, loss=None, loss_scale=None, loss_scale_host=1.0,
                         dent=None, want_dloss=False, out=None):
    """The Jacobian-vector product from gradients that live on the support only (``comp`` = the
    compaction of P: offsets / cols / vals); see smlvm_sparsemax_bwd_ragged in include/smlvm.h.
    Returns dx [rows, K] (and dloss [N] = loss_scale * vals with ``want_dloss``)."""
    rows, K = shape
    dev = comp["vals"].device
    dx = torch.empty((rows, K), dtype=torch.float32, device=dev) if out is None else out
    dloss = torch.empty_like(comp["vals"]) if want_dloss else None
    dp = None if dp is None else _f32c(dp, "dp")
    loss = None if loss is None else _f32c(loss, "loss")
    dent = None if dent is None else _f32c(dent, "dent")
    check(lib.smlvm_sparsemax_bwd_ragged(ptr(comp["offsets"]), ptr(comp["cols"]), ptr(comp["vals"]), ptr(supp),
                                         ptr(dp), ptr(loss), ptr(loss_scale), float(loss_scale_host), ptr(dent),
                                         ptr(dx), ptr(dloss), rows, K, comp["vals"].numel(), stream()),
          "smlvm_sparsemax_bwd_ragged")
    return (dx, dloss) if want_dloss else dx


class _Sparsemax(torch.autograd.Function):
    """P = sparsemax(X) with entmax's saved state (supp_size, P)."""

    @staticmethod
    def forward(ctx, x):
        out = sparsemax_fwd(x, want_nnz=False)
        ctx.save_for_backward(out["supp"], out["p"])
        return out["p"]

    @staticmethod
    def backward(ctx, dp):
        supp, p = ctx.saved_tensors
        return sparsemax_bwd(p, supp, dp=dp)


class _SlotPolicy:
    """Whether the support slots pay, per (device, width), WITHOUT a host synchronisation: after every large forward the
    number of rows whose support does not fit the 8 slots goes to pinned memory with an asynchronous copy; the NEXT
    call reads whatever copy has completed (an event query, never a wait).  Slots stay on while at most 10 % of the
    rows of the last observed batch overflowed them; flat score rows (the signal game at temperature 10: supports of
    20+) switch to the cooperative forward and the dense-read consumers, which do not depend on the support size."""
    _state = {}

    @classmethod
    def want(cls, dev, K):
        st = cls._state.get((dev, K))
        if st is None:
            return True
        if st["event"] is not None and st["event"].query():
            st["frac"] = float(st["host"][0]) / max(st["rows"], 1)
            st["event"] = None
        return st["frac"] <= 0.10

    @classmethod
    def observe(cls, dev, K, nnz):
        st = cls._state.setdefault((dev, K), {"host": torch.zeros(1, dtype=torch.int64).pin_memory(), "event": None,
                                              "frac": 0.0, "rows": 1})
        if st["event"] is not None:
            return                      # a copy is still in flight: keep it
        st["rows"] = nnz.numel()
        st["host"].copy_((nnz > N.SUPPORT_SLOTS).sum().reshape(1), non_blocking=True)
        st["event"] = torch.cuda.Event()
        st["event"].record()


class _SparsemaxStats(torch.autograd.Function):
    """(P, entropy, argmax, nnz) in one launch; backward fuses d entropy / d P."""

    @staticmethod
    def forward(ctx, x):
        # support slots only pay on large batches (rows that do not fit them are processed one by one
        # inside a warp); small batches are launch-bound and keep the plain kernels.  Rows of <= 16
        # columns are no longer than their 64 bytes of slots: the consumers read them directly
        big = x.shape[0] >= SLOTS_MIN_ROWS and x.shape[1] > SLOTS_MIN_COLS
        observe = big and not torch.cuda.is_current_stream_capturing()
        big = big and _SlotPolicy.want(x.device, x.shape[1])
        out = sparsemax_fwd(x, want_entropy=True, want_argmax=True, want_nnz=True, want_slots=big)
        if observe:
            _SlotPolicy.observe(x.device, x.shape[1], out["nnz"])
        ctx.save_for_backward(out["supp"], out["p"])
        slots = out["slots"] if big else torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(out["argmax"], out["nnz"], out["supp"], slots)
        ctx.set_materialize_grads(False)
        return out["p"], out["entropy"], out["argmax"], out["nnz"], out["supp"], slots

    @staticmethod
    def backward(ctx, dp, dent, _damax, _dnnz, _dsupp, _dslots):
        supp, p = ctx.saved_tensors
        if dp is None and dent is None:
            return torch.zeros_like(p)
        return sparsemax_bwd(p, supp, dp=dp, dent=dent)


def sparsemax(X, dim=-1):
    """Drop-in for ``entmax.sparsemax(X, dim=-1)`` on CUDA tensors."""
    if dim not in (-1, X.dim() - 1):
        X = X.transpose(dim, -1)
        return sparsemax(X, -1).transpose(dim, -1)
    shp = X.shape
    return _Sparsemax.apply(X.reshape(-1, shp[-1])).reshape(shp)


def sparsemax_stats(X):
    """X [rows, K] -> (P, entropy [rows], argmax [rows] int64, nnz [rows] int32, supp [rows] int32,
    support slots [rows, 8, 2] for ``compact``)."""
    return _SparsemaxStats.apply(X)


ENTMAX15_MAX_COLS = 1024


class _Entmax15(torch.autograd.Function):
    """P = entmax15(X) over the last dimension of a [rows, K] fp32 CUDA tensor (K <= 1024) -> (P, entropy, argmax)."""

    @staticmethod
    def forward(ctx, x, want_stats):
        rows, K = x.shape
        p = torch.empty_like(x)
        ent = torch.empty(rows, dtype=torch.float32, device=x.device) if want_stats else None
        amax = torch.empty(rows, dtype=torch.int64, device=x.device) if want_stats else None
        check(lib.smlvm_entmax15_fwd(ptr(x), ptr(p), ptr(ent), ptr(amax), rows, K, stream()), "smlvm_entmax15_fwd")
        ctx.save_for_backward(p)
        if not want_stats:
            return p
        ctx.mark_non_differentiable(amax)
        ctx.ent_needs_p = True
        return p, ent, amax

    @staticmethod
    def backward(ctx, dp, dent=None, _damax=None):
        (p,) = ctx.saved_tensors
        if dent is not None:
            # d entropy / d p of marg.py:6-28 on the support (the clamp passes gradient inside [eps, 1 - eps])
            eps = torch.finfo(torch.float32).eps
            inside = (p >= eps) & (p <= 1 - eps)
            de = torch.where(inside, -(torch.log(p.clamp(min=eps)) + 1.0), torch.zeros_like(p)) * dent.unsqueeze(-1)
            dp = de if dp is None else dp + de
        if dp is None:
            return torch.zeros_like(p), None
        dp = dp.contiguous().float()
        dx = torch.empty_like(p)
        check(lib.smlvm_entmax15_bwd(ptr(p), ptr(dp), ptr(dx), p.shape[0], p.shape[1], stream()), "smlvm_entmax15_bwd")
        return dx, None


def entmax15(X, want_stats=False):
    """Drop-in for ``entmax.entmax15(X, dim=-1)`` on fp32 CUDA tensors with at most 1024 columns
    (lvmhelpers/marg.py:46); ``want_stats``: also entropy() of marg.py:6-28 and X.argmax(-1) from the same launch."""
    shp = X.shape
    flat = _f32c(X.reshape(-1, shp[-1]), "scores")
    if not want_stats:
        return _Entmax15.apply(flat, False).reshape(shp)
    p, ent, amax = _Entmax15.apply(flat, True)
    return p.reshape(shp), ent.reshape(shp[:-1]), amax.reshape(shp[:-1])


class _FYLoss(torch.autograd.Function):
    """Fenchel-Young loss rows (entmax.SparsemaxLoss / Entmax15Loss) from scores and the normalizer's output."""

    @staticmethod
    def forward(ctx, x, p, target, alpha15):
        rows, K = x.shape
        loss = torch.empty(rows, dtype=torch.float32, device=x.device)
        pm = torch.empty_like(x)
        check(lib.smlvm_fy_loss_fwd(ptr(x), ptr(p), ptr(target), int(alpha15), ptr(loss), ptr(pm), rows, K, stream()),
              "smlvm_fy_loss_fwd")
        ctx.save_for_backward(pm)
        return loss

    @staticmethod
    def backward(ctx, g):
        (pm,) = ctx.saved_tensors
        g = g.contiguous().float()
        dx = torch.empty_like(pm)
        check(lib.smlvm_row_scale(ptr(pm), ptr(g), ptr(dx), pm.shape[0], pm.shape[1], stream()), "smlvm_row_scale")
        return dx, None, None, None


def fy_loss(x, target, alpha15):
    """[rows] Fenchel-Young losses of scores x [rows, K] against int64 targets; alpha15: entmax15 (else sparsemax)."""
    x = _f32c(x, "scores")
    target = target.to(device=x.device, dtype=torch.int64).contiguous()
    with torch.no_grad():
        p = (_Entmax15.apply(x.detach(), False) if alpha15 else sparsemax_fwd(x.detach(), want_nnz=False)["p"])
    return _FYLoss.apply(x, p, target, bool(alpha15))


# --------------------------------------------------------------------------------------
# (2) compaction
# --------------------------------------------------------------------------------------
def scan_counts(counts):
    """int32 [rows] -> int64 [rows+1] exclusive prefix sum (last entry = total)."""
    N.require_cuda(counts, "counts")
    counts = counts.contiguous()
    rows = counts.numel()
    offsets = torch.empty(rows + 1, dtype=torch.int64, device=counts.device)
    ws = N.scan_workspace(counts.device, rows)
    check(lib.smlvm_scan_counts(ptr(counts), ptr(offsets), rows, ptr(ws), ws.numel(), stream()),
          "smlvm_scan_counts")
    return offsets


def compact_count(p):
    p = _f32c(p, "p")
    rows, K = p.shape
    counts = torch.empty(rows, dtype=torch.int32, device=p.device)
    check(lib.smlvm_compact_count(ptr(p), ptr(counts), rows, K, stream()), "smlvm_compact_count")
    return counts


def compact_fill(p, offsets, total, want_vals=True, want_rows=True, want_cols=True, slots=None):
    p = _f32c(p, "p")
    rows, K = p.shape
    dev = p.device
    vals = torch.empty(total, dtype=torch.float32, device=dev) if want_vals else None
    ri = torch.empty(total, dtype=torch.int32, device=dev) if want_rows else None
    ci = torch.empty(total, dtype=torch.int32, device=dev) if want_cols else None
    check(lib.smlvm_compact_fill(ptr(p), ptr(offsets), ptr(slots), ptr(vals), ptr(ri), ptr(ci), rows, K, int(total),
                                 stream()), "smlvm_compact_fill")
    return vals, ri, ci


def compact(p, nnz=None, slots=None):
    """Row-major compaction of p > 0 (== torch.nonzero(p > 0) order).  ``slots``: the support
    pairs of ``sparsemax_fwd(want_slots=True)`` (the fill then does not re-read p).

    Returns dict(offsets [rows+1] int64, total int, vals, rows, cols).  One D2H read (total).
    """
    if nnz is None:
        nnz = compact_count(p)
    offsets = scan_counts(nnz)
    total = int(offsets[-1].item())
    vals, ri, ci = compact_fill(p, offsets, total, slots=_slots_or_none(slots))
    return dict(offsets=offsets, total=total, vals=vals, rows=ri, cols=ci)


# --------------------------------------------------------------------------------------
# (4) weighted loss
# --------------------------------------------------------------------------------------
def wloss_dense_fwd(p, loss, scale=1.0, want_rows=True, want_total=True, slots=None, want_loss_slots=False):
    """Row losses sum_c p[r,c] * loss[r,c] and scale * their sum (marg.py:125-126) -> (row_loss, total).
    ``slots``: the forward's support pairs -- only the loss entries on the support are then read;
    with ``want_loss_slots`` the gathered losses ([rows, 8]) are returned as a third value for
    ``sparsemax_bwd(loss_slots=...)``."""
    p = _f32c(p, "p")
    loss = _f32c(loss, "loss")
    rows, K = p.shape
    dev = p.device
    row_loss = torch.empty(rows, dtype=torch.float32, device=dev) if want_rows else None
    total = torch.empty((), dtype=torch.float32, device=dev) if want_total else None
    ws = N.reduce_workspace(dev)
    slots = _slots_or_none(slots)
    lslots = torch.empty((rows, N.SUPPORT_SLOTS), dtype=torch.float32, device=dev) \
        if (want_loss_slots and slots is not None) else None
    check(lib.smlvm_wloss_dense_fwd(ptr(p), ptr(loss), ptr(slots), ptr(lslots), ptr(row_loss), ptr(total), float(scale),
                                    rows, K, ptr(ws), ws.numel(), stream()),
          "smlvm_wloss_dense_fwd")
    if want_loss_slots:
        return row_loss, total, lslots
    return row_loss, total


def wloss_ragged_fwd(p, loss, offsets=None, scale=1.0, want_rows=False, want_total=True,
                     want_entropy=False):
    p = _f32c(p, "p")
    loss = None if loss is None else _f32c(loss, "loss")
    n_items = p.numel()
    dev = p.device
    rows = 0 if offsets is None else offsets.numel() - 1
    row_loss = torch.empty(rows, dtype=torch.float32, device=dev) if (want_rows and offsets is not None) else None
    total = torch.empty((), dtype=torch.float32, device=dev) if want_total else None
    ent = torch.empty((), dtype=torch.float32, device=dev) if want_entropy else None
    ws = N.reduce_workspace(dev)
    check(lib.smlvm_wloss_ragged_fwd(ptr(p), ptr(loss), ptr(offsets), ptr(row_loss), ptr(total),
                                     ptr(ent), float(scale), rows, n_items, ptr(ws), ws.numel(),
                                     stream()), "smlvm_wloss_ragged_fwd")
    return row_loss, total, ent


class _MarginalLossDense(torch.autograd.Function):
    """mean_b sum_z P[b,z] L[b,z] for the categorical Marginalizer (marg.py:125-126).

    ``scores`` is listed as an input only so that autograd routes the fused gradient
    (sparsemax Jacobian applied to g * L / B, dP never materialised) straight to it;
    ``p`` must be the detached output of the sparsemax forward on ``scores``.
    """

    @staticmethod
    def forward(ctx, scores, p, supp, losses):
        rows = p.shape[0]
        row_loss, total = wloss_dense_fwd(p, losses, scale=1.0 / rows)
        ctx.save_for_backward(p, supp, losses)
        ctx.mark_non_differentiable(row_loss)
        return total, row_loss

    @staticmethod
    def backward(ctx, g, _grow):
        p, supp, losses = ctx.saved_tensors
        rows = p.shape[0]
        g = g.contiguous().float()
        dscores = None
        dlosses = None
        if ctx.needs_input_grad[0]:
            dscores = sparsemax_bwd(p, supp, loss=losses, loss_scale=g, loss_scale_host=1.0 / rows)
        if ctx.needs_input_grad[3]:
            dlosses = p * (g / rows)
        return dscores, None, None, dlosses


class _RaggedLoss(torch.autograd.Function):
    """scale * (p . loss) over a ragged batch (structmarg.py:140,242)."""

    @staticmethod
    def forward(ctx, p, loss, scale):
        _, total, _ = wloss_ragged_fwd(p, loss, scale=scale, want_total=True)
        ctx.save_for_backward(p, loss)
        ctx.scale = scale
        return total

    @staticmethod
    def backward(ctx, g):
        p, loss = ctx.saved_tensors
        dp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        dl = torch.empty_like(loss) if ctx.needs_input_grad[1] else None
        g = g.contiguous().float()
        check(lib.smlvm_wloss_ragged_bwd(ptr(p), ptr(loss), ptr(g), None, float(ctx.scale),
                                         ptr(dp), ptr(dl), p.numel(), stream()),
              "smlvm_wloss_ragged_bwd")
        return dp, dl, None


class _RaggedEntropy(torch.autograd.Function):
    """-scale * (p . log p) over the (strictly positive) entries of a ragged batch
    (structmarg.py:51-58,200-202)."""

    @staticmethod
    def forward(ctx, p, scale):
        _, _, ent = wloss_ragged_fwd(p, None, scale=scale, want_total=False, want_entropy=True)
        ctx.save_for_backward(p)
        ctx.scale = scale
        return ent

    @staticmethod
    def backward(ctx, gent):
        (p,) = ctx.saved_tensors
        dp = torch.empty_like(p)
        gent = gent.contiguous().float()
        check(lib.smlvm_wloss_ragged_bwd(ptr(p), None, None, ptr(gent), float(ctx.scale),
                                         ptr(dp), None, p.numel(), stream()),
              "smlvm_wloss_ragged_bwd")
        return dp, None


def ragged_loss(p, loss, scale):
    return _RaggedLoss.apply(_f32c(p, "p"), _f32c(loss, "loss"), float(scale))


def ragged_entropy(p, scale):
    return _RaggedEntropy.apply(_f32c(p, "p"), float(scale))


def segment_logsumexp(values, offsets, sign=1.0, shift=0.0):
    """out[r] = logsumexp(sign * values[offsets[r]:offsets[r+1]] + shift)  ([rows] fp32, -inf for an
    empty segment).  The ragged counterpart of the per-sample loop of ``compute_importance_sampling``
    (experiments/bit_vector-vae/train.py:320-345): ``segment_logsumexp(loss_output, offsets, -1,
    logp_z)`` is its ``logp_x_deterministic_term``.  No gradient (evaluation metric)."""
    values = _f32c(values.detach(), "values")
    rows = offsets.numel() - 1
    out = torch.empty(rows, dtype=torch.float32, device=values.device)
    check(lib.smlvm_segment_logsumexp(ptr(values), ptr(offsets), float(sign), float(shift), ptr(out),
                                      rows, stream()), "smlvm_segment_logsumexp")
    return out


# --------------------------------------------------------------------------------------
# (3a) top-k bit-vectors
# --------------------------------------------------------------------------------------
def topk_bits(x, k, want_configs=True, want_bits=True, want_scores=False):
    """x [rows, n] -> (configs [rows,k,n] fp32 | None, bits [rows,k,W] int32 | None,
    dp_scores [rows,k] | None); lvmhelpers/pbinary_topk.pyx:23-34."""
    x = _f32c(x, "scores")
    rows, n = x.shape
    dev = x.device
    W = (n + 31) // 32
    cfg = torch.empty((rows, k, n), dtype=torch.float32, device=dev) if want_configs else None
    bits = torch.empty((rows, k, W), dtype=torch.int32, device=dev) if want_bits else None
    sc = torch.empty((rows, k), dtype=torch.float32, device=dev) if want_scores else None
    check(lib.smlvm_topk_bits(ptr(x), ptr(cfg), ptr(bits), ptr(sc), rows, n, k, stream()),
          "smlvm_topk_bits")
    return cfg, bits, sc


class _BitsScore(torch.autograd.Function):
    """scores_k = einsum('bkj,bj->bk', bits, scores) (structmarg.py:47) from packed bits."""

    @staticmethod
    def forward(ctx, x, bits):
        rows, n = x.shape
        k = bits.shape[1]
        out = torch.empty((rows, k), dtype=torch.float32, device=x.device)
        check(lib.smlvm_bits_score_fwd(ptr(bits), ptr(x), ptr(out), rows, n, k, stream()),
              "smlvm_bits_score_fwd")
        ctx.save_for_backward(bits)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, ds):
        (bits,) = ctx.saved_tensors
        rows, k = ds.shape
        ds = ds.contiguous().float()
        dx = torch.empty((rows, ctx.n), dtype=torch.float32, device=ds.device)
        check(lib.smlvm_bits_score_bwd(ptr(bits), ptr(ds), ptr(dx), rows, ctx.n, k, stream()),
              "smlvm_bits_score_bwd")
        return dx, None


def bits_score(x, bits):
    return _BitsScore.apply(_f32c(x, "scores"), bits)


KSPARSEMAX_MAX_K = 16


class _TopKSparsemax(torch.autograd.Function):
    """scores [B, n] -> (distr [B, k], structured entropy / B, packed candidates, ...) for k <= 16 and n % 8 == 0:
    k-best search, re-scoring, sparsemax over k with the entropy -- three launches, no host read, no support
    compaction (structmarg.py:38-58); backward ONE launch (``smlvm_topk_sparsemax_bwd``)."""

    @staticmethod
    def forward(ctx, scores, k, want_configs):
        rows, n = scores.shape
        dev = scores.device
        cfg, bits, _ = topk_bits(scores, k, want_configs=want_configs, want_bits=True)
        sk = torch.empty((rows, k), dtype=torch.float32, device=dev)
        check(lib.smlvm_bits_score_fwd(ptr(bits), ptr(scores), ptr(sk), rows, n, k, stream()), "smlvm_bits_score_fwd")
        distr = torch.empty_like(sk)
        supp = torch.empty(rows, dtype=torch.int32, device=dev)
        nnz = torch.empty(rows, dtype=torch.int32, device=dev)
        ent = torch.empty((), dtype=torch.float32, device=dev)
        ws = N.reduce_workspace(dev)
        scale = 1.0 / rows
        check(lib.smlvm_ksparsemax_fwd(ptr(sk), ptr(distr), ptr(supp), ptr(nnz), None, ptr(ent), scale, rows, k,
                                       ptr(ws), ws.numel(), stream()), "smlvm_ksparsemax_fwd")
        ctx.save_for_backward(bits, distr, supp)
        ctx.n, ctx.scale = n, scale
        if cfg is None:
            cfg = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(bits, cfg, nnz, supp)
        ctx.set_materialize_grads(False)
        return distr, ent, bits, cfg, nnz, supp

    @staticmethod
    def backward(ctx, ddistr, dent, _db, _dc, _dn, _ds):
        bits, distr, supp = ctx.saved_tensors
        rows, k = distr.shape
        dx = torch.empty((rows, ctx.n), dtype=torch.float32, device=distr.device)
        if ddistr is None and dent is None:
            return dx.zero_(), None, None
        ddistr = None if ddistr is None else ddistr.contiguous().float()
        dent = None if dent is None else dent.contiguous().float()
        check(lib.smlvm_topk_sparsemax_bwd(ptr(bits), ptr(distr), ptr(supp), ptr(ddistr), ptr(dent), ctx.scale, ptr(dx),
                                           rows, ctx.n, k, stream()), "smlvm_topk_sparsemax_bwd")
        return dx, None, None


def topk_sparsemax_supported(n, k):
    return 2 <= k <= KSPARSEMAX_MAX_K and n % 8 == 0


def topk_sparsemax(scores, k, want_configs=False):
    """(distr [B,k], entropy scalar, bits [B,k,W] int32, configs [B,k,n] | None, nnz [B] int32, supp [B] int32)."""
    scores = _f32c(scores, "scores")
    distr, ent, bits, cfg, nnz, supp = _TopKSparsemax.apply(scores, int(k), bool(want_configs))
    return distr, ent, bits, (cfg if want_configs else None), nnz, supp


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, idx):
        out = torch.empty((idx.numel(),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        row_bytes = (src.numel() // max(src.shape[0], 1)) * src.element_size()
        check(lib.smlvm_gather_rows(ptr(src), ptr(idx), ptr(out), idx.numel(), row_bytes, stream()),
              "smlvm_gather_rows")
        ctx.save_for_backward(idx)
        ctx.src_shape = src.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        d = torch.zeros(ctx.src_shape, dtype=dout.dtype, device=dout.device)
        d.index_add_(0, idx.long(), dout)
        return d, None


def gather_rows(t, idx):
    """``t[idx]`` along dim 0 for the compacted row ids (int32) of the support: the repeat-and-mask of
    structmarg.py:95-124 and the ``[idxs]`` gathers of :236-240.  Rows whose size is not a multiple of
    4 bytes, empty tensors and non-CUDA tensors go through ``index_select``."""
    if not torch.is_tensor(t):
        return t
    row_bytes = (t.numel() // t.shape[0]) * t.element_size() if t.dim() > 0 and t.shape[0] > 0 else 0
    if not t.is_cuda or row_bytes == 0 or row_bytes % 4 != 0 or idx.numel() == 0:
        return t.index_select(0, idx.to(torch.int64) if idx.dtype != torch.int64 else idx)
    if idx.dtype != torch.int32 or idx.device != t.device:
        idx = idx.to(device=t.device, dtype=torch.int32)
    idx, t = idx.contiguous(), t.contiguous()
    if t.requires_grad and torch.is_grad_enabled():
        return _GatherRows.apply(t, idx)
    # data (no gradient): skip the autograd node, one launch
    out = torch.empty((idx.numel(),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    check(lib.smlvm_gather_rows(ptr(t), ptr(idx), ptr(out), idx.numel(), row_bytes, stream()), "smlvm_gather_rows")
    return out


def bits_expand(bits, index, n):
    """packed bits [..., W] int32 + candidate ids [N] int64 (None: every row) -> fp32 {0,1} [N, n]:
    ``bit_vector_z.view(-1, n)[mask]`` (structmarg.py:116-126) without the dense [B,k,n] tensor."""
    W = (n + 31) // 32
    if not bits.is_cuda:
        raise RuntimeError("lvmhelpers (B200): packed bits must be a CUDA tensor; there is no CPU fallback")
    flat = bits.contiguous().view(-1, W)
    if index is not None:
        index = index.to(device=bits.device, dtype=torch.int64).contiguous()
    n_out = flat.shape[0] if index is None else index.numel()
    out = torch.empty((n_out, n), dtype=torch.float32, device=bits.device)
    check(lib.smlvm_bits_expand(ptr(flat), ptr(index), ptr(out), n_out, n, stream()), "smlvm_bits_expand")
    return out


# --------------------------------------------------------------------------------------
# (3b) SparseMAP
# --------------------------------------------------------------------------------------
def sparsemap_plan(n, kind=0):
    """[(capacity, shared-memory bytes per CTA, CTAs per SM)] for each capacity tier of the solver."""
    import ctypes
    caps, smem, ctas = ((ctypes.c_int32 * 4)() for _ in range(3))
    nt = lib.smlvm_sparsemap_plan(int(n), int(kind), caps, smem, ctas)
    return [(caps[i], smem[i], ctas[i]) for i in range(nt)]


# The inverses the backward needs (|A|^2 doubles per sample) are bump-allocated by the solver from ONE arena.  Up to
# this many bytes the arena is sized for the worst case, (n+1)^2 per sample, and cannot overflow; beyond it (the
# sweep's 65536+ samples at n = 128: 133 KB per sample padded, ~16-28 KB used) it is sized from the largest per-sample
# need seen so far for this (n, factor, budget), and a caller that reads the solver's cursor (one scalar, together
# with the ragged sizes it reads anyway) re-solves with the exact size on overflow.
S_WORST_CASE_MAX_BYTES = 2 << 30
SMAP_CHUNK_LANES = 1      # streams a large batch is split over (1 = one launch sequence; 2 measured slower)
SMAP_SPLIT_MIN_ROWS = 16384
_S_NEED = {}              # (n, kind, budget) -> doubles per sample (largest mean seen)


_SIDE_STREAMS = {}


def _side_stream(dev):
    """One side stream per device for the life of the process: the caching allocator keeps a pool per stream,
    so a fresh stream per call would send every call's temporaries to cudaMalloc."""
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def _handoff_pool_bytes(rows, n):
    """Room for the states of the samples that outgrow a capacity tier (resumed by the next tier instead of re-solved):
    3-7 % of the samples at n = 128, ~8 (|A|+1)^2 bytes each; sized for 10 % at the second tier's entry size, capped."""
    if n < 48:
        return 0        # a single tier: nothing is handed over
    per = 8 * (n * 3 // 4 + 1) ** 2
    return int(min(max(rows // 10, 64) * per, 4 << 30))


def sparsemap_arena_doubles(rows, n, kind, budget=0):
    """Size (in doubles) of the arena ``sparsemap_solve`` allocates for the inverses of ``rows`` samples."""
    worst = rows * (n + 1) ** 2
    if worst * 8 <= S_WORST_CASE_MAX_BYTES:
        return worst
    per_sample = _S_NEED.get((n, kind, budget), 0.45 * (n + 1) ** 2)
    return int(min(worst, max(S_WORST_CASE_MAX_BYTES // 8, rows * per_sample * 1.25 + 65536)))


def sparsemap_solve(x, kind, budget=0, t=None, init_scores=None, max_iter=100, keep_S=True, arena_doubles=None):
    """Batched active-set solve.  x [rows, n] -> padded solver state (dict).  Asynchronous: no host read, no
    allocation inside the library.  ``state["S"]`` is the arena of the inverses (or None), ``state["S_offsets"]``
    each sample's offset in it (-1: did not fit), ``state["S_cursor"]`` the device counter that ends at the number
    of doubles needed, ``state["S_capacity"]`` the arena size; ``sparsemap_arena_overflow(state)`` compares them."""
    x = _f32c(x, "scores")
    rows, n = x.shape
    dev = x.device
    cap = lib.smlvm_sparsemap_capacity(n)
    W = (n + 31) // 32
    t = None if t is None else _f32c(t, "t")
    if init_scores is not None:
        init_scores = init_scores.to(device=dev, dtype=torch.float64).contiguous()
        assert init_scores.shape == (rows, 2 * n)
    p_pad = torch.empty((rows, cap), dtype=torch.float32, device=dev)
    bits = torch.empty((rows, cap, W), dtype=torch.int32, device=dev)
    counts = torch.empty(rows, dtype=torch.int32, device=dev)
    kept = torch.empty(rows, dtype=torch.int32, device=dev)
    iters = torch.empty(rows, dtype=torch.int32, device=dev)
    status = torch.empty(rows, dtype=torch.int32, device=dev)
    S = S_off = cursor = None
    capacity = 0
    if keep_S:
        capacity = int(arena_doubles) if arena_doubles is not None else sparsemap_arena_doubles(rows, n, kind, budget)
        S = torch.empty(max(capacity, 1), dtype=torch.float64, device=dev)
        S_off = torch.empty(rows, dtype=torch.int64, device=dev)
        cursor = torch.zeros(1, dtype=torch.int64, device=dev)
    state = dict(p_pad=p_pad, bits=bits, counts=counts, kept=kept, iters=iters, status=status, S=S, S_offsets=S_off,
                 S_cursor=cursor, S_capacity=capacity, n=n, rows=rows, kind=kind, budget=budget)
    if rows == 0:
        return state

    def solve(r0, r1):
        sl = slice(r0, r1)
        ws = torch.empty(lib.smlvm_sparsemap_workspace_bytes(r1 - r0) + _handoff_pool_bytes(r1 - r0, n), dtype=torch.uint8,
                         device=dev)
        check(lib.smlvm_sparsemap_fwd(ptr(x[sl]), ptr(None if t is None else t[sl]),
                                      ptr(None if init_scores is None else init_scores[sl]), kind, budget, max_iter,
                                      r1 - r0, n, ptr(p_pad[sl]), ptr(bits[sl]), ptr(counts[sl]), ptr(kept[sl]),
                                      ptr(iters[sl]), ptr(status[sl]), ptr(S), capacity,
                                      ptr(None if S_off is None else S_off[sl]), ptr(cursor), ptr(ws), ws.numel(),
                                      stream()), "smlvm_sparsemap_fwd")
        return ws

    if rows < SMAP_SPLIT_MIN_ROWS or SMAP_CHUNK_LANES < 2:
        solve(0, rows)
        return state
    # Two halves on two streams over the same arena: a launch sequence ends with a latency-bound tail (the few
    # samples that outgrew the first capacity tier are re-solved from scratch by a small persistent grid; they are the
    # LARGE ones, ~3 ms each at n = 128) under which the other half's first tier keeps the SMs busy.
    cur = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    half = (rows + 1) // 2
    solve(0, half)
    side.wait_stream(cur)            # inputs, the zeroed cursor
    with torch.cuda.stream(side):
        ws = solve(half, rows)
    ws.record_stream(side)
    cur.wait_stream(side)
    return state


def sequence_layout(num_states, device):
    """int32 [2L+3] on ``device``: offsets of the states of position 0..L, then the first edge of position 0..L+1, in
    the order FactorSequence::Initialize numbers them (sequence.h:264-291)."""
    ns = [int(s) for s in num_states]
    L = len(ns)
    off, o = [], 0
    for s_ in ns:
        off.append(o)
        o += s_
    off.append(o)
    eb, e = [], 0
    for i in range(L + 1):
        eb.append(e)
        e += (ns[i - 1] if i > 0 else 1) * (ns[i] if i < L else 1)
    eb.append(e)
    return torch.tensor(off + eb, dtype=torch.int32, device=device), o, e


def sparsemap_solve_sequence(x, edge_scores, num_states, max_iter=100):
    """General linear chain (lvmhelpers/sequence.h; PFactorSequence): x [rows, sum num_states] node scores,
    edge_scores [rows, E] (None: zeros) -> the padded solver state of ``sparsemap_solve`` (vertices as one-hot bits
    over the (position, state) variables) + the layout.  The arena of the inverses is sized for the worst case."""
    x = _f32c(x, "scores")
    rows, n = x.shape
    dev = x.device
    lay, V, E = sequence_layout(num_states, dev)
    if V != n:
        raise ValueError("scores have %d columns, num_states sums to %d" % (n, V))
    if max(num_states) > 32 or n > N.SPARSEMAP_MAX_BITS:
        raise ValueError("the general chain holds at most 32 states per position and %d variables" % N.SPARSEMAP_MAX_BITS)
    if edge_scores is not None:
        edge_scores = _f32c(edge_scores, "edge scores")
        if tuple(edge_scores.shape) != (rows, E):
            raise ValueError("edge scores must be [%d, %d]" % (rows, E))
    cap = lib.smlvm_sparsemap_capacity(n)
    W = (n + 31) // 32
    p_pad = torch.empty((rows, cap), dtype=torch.float32, device=dev)
    bits = torch.empty((rows, cap, W), dtype=torch.int32, device=dev)
    counts, kept, iters, status = (torch.empty(rows, dtype=torch.int32, device=dev) for _ in range(4))
    capacity = max(rows * (n + 1) ** 2, 1)
    S = torch.empty(capacity, dtype=torch.float64, device=dev)
    S_off = torch.empty(rows, dtype=torch.int64, device=dev)
    cursor = torch.zeros(1, dtype=torch.int64, device=dev)
    state = dict(p_pad=p_pad, bits=bits, counts=counts, kept=kept, iters=iters, status=status, S=S, S_offsets=S_off,
                 S_cursor=cursor, S_capacity=capacity, n=n, rows=rows, kind=N.FACTOR_SEQUENCE, budget=0,
                 layout=lay, length=len(num_states), num_edges=E, num_states=[int(s) for s in num_states])
    if rows == 0:
        return state
    ws = torch.empty(lib.smlvm_sparsemap_workspace_bytes(rows) + _handoff_pool_bytes(rows, n), dtype=torch.uint8, device=dev)
    check(lib.smlvm_sparsemap_seq_fwd(ptr(x), ptr(edge_scores), ptr(lay), len(num_states), E, int(max_iter), rows, n,
                                      ptr(p_pad), ptr(bits), ptr(counts), ptr(kept), ptr(iters), ptr(status), ptr(S),
                                      capacity, ptr(S_off), ptr(cursor), ptr(ws), ws.numel(), stream()),
          "smlvm_sparsemap_seq_fwd")
    return state


def sparsemap_bwd_sequence(state, dp, offsets, only_positive):
    """(dx [rows, n], dedge [rows, E]) of the general chain."""
    rows, n, E = state["rows"], state["n"], state["num_edges"]
    dp = dp.contiguous().float()
    dx = torch.empty((rows, n), dtype=torch.float32, device=dp.device)
    de = torch.empty((rows, E), dtype=torch.float32, device=dp.device)
    check(lib.smlvm_sparsemap_seq_bwd(ptr(dp), ptr(offsets), int(only_positive), ptr(state["p_pad"]), ptr(state["bits"]),
                                      ptr(state["counts"]), ptr(state["S"]), ptr(state["S_offsets"]), ptr(state["layout"]),
                                      state["length"], E, ptr(dx), ptr(de), rows, n, stream()), "smlvm_sparsemap_seq_bwd")
    return dx, de


def onehot_to_states(aset, num_states):
    """[N, sum num_states] one-hot rows -> [N, L] fp32 state ids (the configurations get_sparse_solution returns)."""
    cols, o = [], 0
    for s_ in num_states:
        cols.append(aset[:, o:o + s_].argmax(-1))
        o += s_
    return torch.stack(cols, dim=-1).to(torch.float32)


def sparsemap_arena_overflow(state, needed=None):
    """Doubles the arena should have had if some inverse did not fit, else 0.  Reads the solver's cursor (a host
    synchronisation) unless the caller already has it (``needed``); remembers the per-sample need."""
    if state["S"] is None or state["rows"] == 0:
        return 0
    needed = int(state["S_cursor"].item()) if needed is None else int(needed)
    key = (state["n"], state["kind"], state["budget"])
    _S_NEED[key] = max(_S_NEED.get(key, 0.0), needed / state["rows"])
    return needed if needed > state["S_capacity"] else 0


def sparsemap_expand(state, only_positive, offsets, total, want_aset=True, want_idx=True):
    rows, n = state["rows"], state["n"]
    dev = state["p_pad"].device
    p = torch.empty(total, dtype=torch.float32, device=dev)
    aset = torch.empty((total, n), dtype=torch.float32, device=dev) if want_aset else None
    idx = torch.empty(total, dtype=torch.int32, device=dev) if want_idx else None
    check(lib.smlvm_sparsemap_expand(ptr(state["p_pad"]), ptr(state["bits"]), ptr(state["counts"]),
                                     ptr(offsets), int(only_positive), ptr(p), ptr(aset), ptr(idx),
                                     rows, n, stream()), "smlvm_sparsemap_expand")
    return p, aset, idx


def sparsemap_bwd(state, dp, offsets, only_positive, want_dt=False):
    rows, n = state["rows"], state["n"]
    dev = dp.device
    dp = dp.contiguous().float()
    dx = torch.empty((rows, n), dtype=torch.float32, device=dev)
    dt = torch.empty((rows, max(n - 1, 1)), dtype=torch.float32, device=dev) if want_dt else None
    if state["S"] is None:
        raise RuntimeError("sparsemap_bwd: the forward solve was run with keep_S=False")
    check(lib.smlvm_sparsemap_bwd(ptr(dp), ptr(offsets), int(only_positive), ptr(state["p_pad"]),
                                  ptr(state["bits"]), ptr(state["counts"]), ptr(state["S"]), ptr(state["S_offsets"]),
                                  ptr(dx), ptr(dt), rows, n, stream()), "smlvm_sparsemap_bwd")
    if want_dt:
        dt = dt[:, :n - 1]
    return dx, dt


class _SparseMAPBatch(torch.autograd.Function):
    """Batched SparseMAP: scores [rows, n] -> ragged (p [N], aset [N, n]) in sample order.

    One autograd node for the whole batch instead of the reference's one node (and one
    factor graph) per sample (lvmhelpers/structmarg.py:181-196, sparsemap.py:90-127).
    """

    @staticmethod
    def forward(ctx, x, t, kind, budget, max_iter, init_scores, only_positive, aux):
        if aux.get("static"):
            return _SparseMAPBatch._forward_static(ctx, x, t, kind, budget, max_iter, init_scores, only_positive, aux)
        state = sparsemap_solve(x, kind, budget=budget, t=t, init_scores=init_scores,
                                max_iter=max_iter, keep_S=True)
        sizes = state["kept"] if only_positive else state["counts"]
        offsets = scan_counts(sizes)
        host = torch.cat([offsets[-1:], state["S_cursor"], sizes.to(torch.int64)]).cpu()  # the one D2H per batch
        needed = sparsemap_arena_overflow(state, int(host[1]))
        if needed:   # first large batch of this shape: the inverses did not fit the estimated arena -- exact size now
            state = sparsemap_solve(x, kind, budget=budget, t=t, init_scores=init_scores, max_iter=max_iter,
                                    keep_S=True, arena_doubles=needed)
            sizes = state["kept"] if only_positive else state["counts"]
        host = torch.cat([host[:1], host[2:]])
        total = int(host[0])
        p, aset, idx = sparsemap_expand(state, only_positive, offsets, total)
        ctx.state = state
        ctx.offsets = offsets
        ctx.only_positive = only_positive
        ctx.has_t = t is not None
        ctx.mark_non_differentiable(aset)
        ctx.set_materialize_grads(False)  # no [N, n] zeros for the non-differentiable vertices in backward
        aux["sizes"] = host[1:]
        aux["idx"] = idx
        aux["offsets"] = offsets
        aux["state"] = state
        return p, aset

    @staticmethod
    def _forward_static(ctx, x, t, kind, budget, max_iter, init_scores, only_positive, aux):
        """No host read: the ragged outputs are allocated for the worst case, rows * (n + 1) entries -- the
        reference's (p, aset) in sample order, followed by zeros (p = 0: no weight in the loss, none in the entropy).
        Fixed shapes and no synchronisation: the whole forward + backward can be captured in a CUDA graph."""
        rows, n = x.shape
        if rows * (n + 1) ** 2 * 8 > S_WORST_CASE_MAX_BYTES:
            raise ValueError("static SparseMAP outputs need the worst-case arena (rows * (n+1)^2 doubles <= 2 GB)")
        state = sparsemap_solve(x, kind, budget=budget, t=t, init_scores=init_scores, max_iter=max_iter, keep_S=True,
                                arena_doubles=rows * (n + 1) ** 2)
        sizes = state["kept"] if only_positive else state["counts"]
        offsets = scan_counts(sizes)
        total = rows * (n + 1)
        dev = x.device
        p = torch.zeros(total, dtype=torch.float32, device=dev)
        aset = torch.zeros((total, n), dtype=torch.float32, device=dev)
        idx = torch.zeros(total, dtype=torch.int32, device=dev)
        check(lib.smlvm_sparsemap_expand(ptr(state["p_pad"]), ptr(state["bits"]), ptr(state["counts"]), ptr(offsets),
                                         int(only_positive), ptr(p), ptr(aset), ptr(idx), rows, n, stream()),
              "smlvm_sparsemap_expand")
        ctx.state, ctx.offsets, ctx.only_positive, ctx.has_t = state, offsets, only_positive, t is not None
        ctx.mark_non_differentiable(aset)
        ctx.set_materialize_grads(False)
        aux.update(sizes=sizes, idx=idx, offsets=offsets, state=state)
        return p, aset

    @staticmethod
    def backward(ctx, dp, _daset):
        if dp is None:
            return None, None, None, None, None, None, None, None
        dx, dt = sparsemap_bwd(ctx.state, dp, ctx.offsets, ctx.only_positive, want_dt=ctx.has_t)
        return dx, dt, None, None, None, None, None, None


def sparsemap_batch(x, kind, budget=0, t=None, max_iter=100, init_scores=None, only_positive=True, static=False):
    """Returns (p [N], aset [N, n], aux) with aux = dict(sizes [rows] int64 CPU, idx [N] int32,
    offsets, state).  ``static``: N = rows * (n + 1) whatever the supports (zero padding behind the ragged
    entries), ``aux["sizes"]`` stays on the device (int32) and nothing is read back -- graph-capturable."""
    aux = {"static": bool(static)}
    p, aset = _SparseMAPBatch.apply(_f32c(x, "scores"), t, kind, int(budget), int(max_iter),
                                    init_scores, bool(only_positive), aux)
    return p, aset, aux
