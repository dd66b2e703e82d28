"""CPU oracle for the sparse-marginalization hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``sparse-marginalization-lvm_b200/lvmhelpers``) never does: it calls the CUDA
C-ABI library and fails loudly when that is missing.

What is restated, and how each piece is pinned:

* ``sparsemax_*`` -- ``entmax.sparsemax`` (third-party PyPI ``entmax``, unpinned
  by the reference: requirements.txt:2; call sites lvmhelpers/marg.py:51,
  lvmhelpers/structmarg.py:48).  Restated with the same torch CPU ops
  (max-shift, ``torch.sort`` descending, ``cumsum`` which accumulates fp64 on
  CPU, strict ``>`` support test, ``tau = cumsum[k-1]/k``, ``clamp``).
  PARITY UNPINNED against the real package (absent here); pinned instead by
  restatement-independent properties (tests/test_oracle.py).
* ``topk`` -- lvmhelpers/binary_topk.h:57-129.  Pinned against the reference
  itself, compiled from its sources into ``oracle/_ref`` (tests/test_oracle_ref.py)
  and by brute-force enumeration.
* ``sparsemap_*`` -- libad3 ``GenericFactor::SolveQP`` (third-party, absent) +
  the in-tree factor headers.  PARITY UNPINNED for ``(p, aset)``; marginals are
  pinned by closed forms and KKT checks; factor MAP oracles are pinned against
  the reference headers compiled into ``oracle/_ref/libref_factors.so``.
* ``entropy`` / ``struct_entropy`` -- lvmhelpers/marg.py:6-28,
  lvmhelpers/structmarg.py:51-58,200-202 (in-tree, restated verbatim in torch).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
KIND_BERNOULLI, KIND_BUDGET, KIND_SEQUENCE_BINARY = 0, 1, 2
STATUS_NAMES = {0: "converged", 1: "max_iter", 2: "repeated_config", 3: "singular"}


# ----------------------------------------------------------------------------
# build / load
# ----------------------------------------------------------------------------
def build(ref=True):
    """Compile the C/C++ restatement (and oracle/_ref when /root/reference is here)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])
    if ref and os.path.isdir("/root/reference/lvmhelpers"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref")])


_f64p = ctypes.POINTER(ctypes.c_double)
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_intp = ctypes.POINTER(ctypes.c_int)


def _bind_solver(lib):
    lib.smlvm_oracle_sparsemap_fwd.restype = ctypes.c_int
    lib.smlvm_oracle_sparsemap_fwd.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p, _f64p, _f64p,
        ctypes.c_int, ctypes.c_int, _f64p, _i32p, _f64p, _intp, _intp, _intp, _f64p]
    lib.smlvm_oracle_sparsemap_fwd_bwd.restype = ctypes.c_int
    lib.smlvm_oracle_sparsemap_fwd_bwd.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p, _f64p, _f64p,
        ctypes.c_int, _f64p, ctypes.c_int, _f64p, _f64p]
    lib.smlvm_oracle_factor_map.restype = ctypes.c_int
    lib.smlvm_oracle_factor_map.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p, _f64p, _i32p, _f64p]
    lib.smlvm_oracle_seq_fwd.restype = ctypes.c_int
    lib.smlvm_oracle_seq_fwd.argtypes = [ctypes.c_int, _i32p, _f64p, _f64p, ctypes.c_int, ctypes.c_int, _f64p, _i32p, _f64p,
                                         _intp, _intp, _intp]
    lib.smlvm_oracle_seq_fwd_bwd.restype = ctypes.c_int
    lib.smlvm_oracle_seq_fwd_bwd.argtypes = [ctypes.c_int, _i32p, _f64p, _f64p, ctypes.c_int, _f64p, ctypes.c_int, _f64p, _f64p]
    lib.smlvm_oracle_seq_map.restype = ctypes.c_int
    lib.smlvm_oracle_seq_map.argtypes = [ctypes.c_int, _i32p, _f64p, _f64p, _i32p, _f64p]
    lib.smlvm_oracle_sparsemap_batch.restype = ctypes.c_long
    lib.smlvm_oracle_sparsemap_batch.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_long,
        ctypes.c_int, _intp, _intp]
    return lib


_LIBS = {}


def _lib(which="oracle"):
    """which: 'oracle' (our restatement) | 'ref_factors' | 'ref_topk' (oracle/_ref)."""
    if which in _LIBS:
        return _LIBS[which]
    path = {
        "oracle": os.path.join(_HERE, "_build", "liboracle.so"),
        "ref_factors": os.path.join(_HERE, "_ref", "libref_factors.so"),
        "ref_topk": os.path.join(_HERE, "_ref", "libref_topk.so"),
    }[which]
    if not os.path.exists(path) and which == "oracle":
        build(ref=False)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    lib = ctypes.CDLL(path)
    if which in ("oracle", "ref_factors"):
        _bind_solver(lib)
    if which == "oracle":
        lib.smlvm_oracle_topk.restype = ctypes.c_int
        lib.smlvm_oracle_topk.argtypes = [_f32p, _f32p, _f32p, ctypes.c_long,
                                          ctypes.c_int, ctypes.c_int]
    if which == "ref_topk":
        lib.smlvm_ref_topk.restype = ctypes.c_int
        lib.smlvm_ref_topk.argtypes = [_f32p, _f32p, _f32p, ctypes.c_long,
                                       ctypes.c_int, ctypes.c_int]
    _LIBS[which] = lib
    return lib


def have_ref(which):
    return os.path.exists(os.path.join(
        _HERE, "_ref", {"ref_factors": "libref_factors.so", "ref_topk": "libref_topk.so"}[which]))


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(ct)


# ----------------------------------------------------------------------------
# sparsemax (entmax.sparsemax restated)
# ----------------------------------------------------------------------------
def sparsemax_fwd(X):
    """X [rows, K] fp32 CPU -> (P [rows, K] fp32, supp_size [rows] int64).

    entmax ``SparsemaxFunction.forward`` / ``_sparsemax_threshold_and_support``.
    """
    X = X.detach().to("cpu", torch.float32)
    max_val, _ = X.max(dim=-1, keepdim=True)
    Xs = X - max_val
    topk, _ = torch.sort(Xs, dim=-1, descending=True)
    topk_cumsum = topk.cumsum(-1) - 1
    rhos = torch.arange(1, X.shape[-1] + 1, dtype=X.dtype).view(1, -1)
    support = rhos * topk > topk_cumsum
    support_size = support.sum(dim=-1).unsqueeze(-1)
    tau = topk_cumsum.gather(-1, support_size - 1)
    tau = tau / support_size.to(X.dtype)
    P = torch.clamp(Xs - tau, min=0)
    return P, support_size.squeeze(-1)


def sparsemax_bwd(P, supp_size, dP):
    """entmax ``SparsemaxFunction.backward``: masked mean-centred gradient."""
    P = P.detach().to("cpu", torch.float32)
    grad_input = dP.detach().to("cpu", torch.float32).clone()
    grad_input[P == 0] = 0
    v_hat = grad_input.sum(dim=-1) / supp_size.to(P.dtype)
    v_hat = v_hat.unsqueeze(-1)
    return torch.where(P != 0, grad_input - v_hat, grad_input)


class SparsemaxFunction(torch.autograd.Function):
    """Autograd pairing of the two restated halves (CPU)."""

    @staticmethod
    def forward(ctx, X):
        P, k = sparsemax_fwd(X)
        ctx.save_for_backward(k, P)
        return P

    @staticmethod
    def backward(ctx, dP):
        k, P = ctx.saved_tensors
        return sparsemax_bwd(P, k, dP)


def sparsemax(X, dim=-1):
    assert dim in (-1, X.dim() - 1)
    shp = X.shape
    return SparsemaxFunction.apply(X.reshape(-1, shp[-1])).reshape(shp)


def entropy(p):
    """lvmhelpers/marg.py:6-28 (duplicate at lvmhelpers/structmarg.py:9-20)."""
    nz = p > 0
    eps = torch.finfo(p.dtype).eps
    p_stable = p.clone().clamp(min=eps, max=1 - eps)
    out = torch.where(nz, p_stable * torch.log(p_stable),
                      torch.tensor(0., dtype=torch.float))
    return -(out).sum(-1)


def struct_entropy(distr, batch_size):
    """lvmhelpers/structmarg.py:51-58 / :200-202: -p+ . log p+ / B."""
    flat = distr.reshape(-1)
    flat = flat[flat > 0]
    return -(flat @ torch.log(flat)) / batch_size


# ----------------------------------------------------------------------------
# top-k bit-vectors (binary_topk.h restated in C)
# ----------------------------------------------------------------------------
def topk(scores, k, which="oracle", return_scores=False):
    """scores [rows, n] fp32 -> configs [rows, k, n] fp32 in {0,1} (+ DP scores)."""
    s = np.ascontiguousarray(scores, dtype=np.float32)
    rows, n = s.shape
    out = np.empty((rows, k, n), dtype=np.float32)
    osc = np.empty((rows, k), dtype=np.float32)
    lib = _lib(which)
    fn = lib.smlvm_oracle_topk if which == "oracle" else lib.smlvm_ref_topk
    rc = fn(_p(s, _f32p), _p(out, _f32p), _p(osc, _f32p), rows, n, k)
    if rc != 0:
        raise ValueError("top-k oracle: bad arguments (k=%d, n=%d)" % (k, n))
    return (out, osc) if return_scores else out


def topk_bruteforce(x, k):
    """Enumerate all 2^n bit-vectors (n <= 16).  Returns the sorted list of the
    k best *exact* (fp64) scores -- configurations are only unique without ties."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    codes = np.arange(1 << n, dtype=np.int64)
    bits = ((codes[:, None] >> np.arange(n)[None, :]) & 1).astype(np.float64)
    sc = bits @ x
    order = np.argsort(-sc, kind="stable")[:k]
    return sc[order], bits[order]


# ----------------------------------------------------------------------------
# SparseMAP (restated AD3 SolveQP + factors; SURVEY.md §8 a-8 .. a-12)
# ----------------------------------------------------------------------------
def sparsemap_fwd(x, kind=KIND_BERNOULLI, budget=0, t=None, init=None,
                  max_iter=100, which="oracle"):
    """One sample.  x [n] -> dict(p [|A|] f64, aset [|A|, n] int32, S [|A|,|A|] f64,
    iters, status, u [2n] f64).  ``init`` = the 2n uniform draws of sparsemap.py:26."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    cap = n + 2
    t_ = None if t is None else np.ascontiguousarray(t, dtype=np.float64)
    i_ = None if init is None else np.ascontiguousarray(init, dtype=np.float64)
    p = np.zeros(cap, dtype=np.float64)
    aset = np.zeros((cap, n), dtype=np.int32)
    S = np.zeros(cap * cap, dtype=np.float64)
    u = np.zeros(2 * n, dtype=np.float64)
    cnt, it, st = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    rc = _lib(which).smlvm_oracle_sparsemap_fwd(
        kind, n, budget, _p(x, _f64p), _p(t_, _f64p), _p(i_, _f64p), max_iter, cap,
        _p(p, _f64p), _p(aset, _i32p), _p(S, _f64p),
        ctypes.byref(cnt), ctypes.byref(it), ctypes.byref(st), _p(u, _f64p))
    if rc != 0:
        raise RuntimeError("sparsemap oracle failed rc=%d" % rc)
    m = cnt.value
    return dict(p=p[:m].copy(), aset=aset[:m].copy(), S=S[:m * m].reshape(m, m).copy(),
                iters=it.value, status=st.value, u=u)


def sparsemap_bwd(x, dp, kind=KIND_BERNOULLI, budget=0, t=None, init=None,
                  max_iter=100, which="oracle"):
    """One sample, re-solves then applies dist_jacobian_vec.  Returns (dx [n], dt [n-1]|None)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    dp = np.ascontiguousarray(dp, dtype=np.float64)
    t_ = None if t is None else np.ascontiguousarray(t, dtype=np.float64)
    i_ = None if init is None else np.ascontiguousarray(init, dtype=np.float64)
    dx = np.zeros(n, dtype=np.float64)
    dt = np.zeros(max(n - 1, 1), dtype=np.float64) if kind == KIND_SEQUENCE_BINARY else None
    rc = _lib(which).smlvm_oracle_sparsemap_fwd_bwd(
        kind, n, budget, _p(x, _f64p), _p(t_, _f64p), _p(i_, _f64p), max_iter,
        _p(dp, _f64p), dp.shape[0], _p(dx, _f64p), _p(dt, _f64p))
    if rc != 0:
        raise RuntimeError("sparsemap oracle bwd failed rc=%d" % rc)
    return dx, (None if dt is None else dt[:n - 1])


def factor_map(eta2n, kind, budget=0, t=None, which="oracle"):
    """Factor MAP oracle alone: eta [2n] (+t) -> (y [n] int32, value)."""
    eta = np.ascontiguousarray(eta2n, dtype=np.float64)
    n = eta.shape[0] // 2
    t_ = None if t is None else np.ascontiguousarray(t, dtype=np.float64)
    y = np.zeros(n, dtype=np.int32)
    val = ctypes.c_double(0)
    rc = _lib(which).smlvm_oracle_factor_map(kind, n, budget, _p(eta, _f64p),
                                             _p(t_, _f64p), _p(y, _i32p), ctypes.byref(val))
    if rc != 0:
        raise RuntimeError("factor_map failed")
    return y, val.value


def seq_num_edges(num_states):
    """Additional (edge) scores of the general chain: start, transitions, stop (sequence.h:264-291)."""
    ns = list(num_states)
    return ns[0] + sum(a * b for a, b in zip(ns[:-1], ns[1:])) + ns[-1]


def seq_fwd(x, add, num_states, max_iter=100, which="oracle"):
    """General linear chain, one sample.  x [sum num_states] node scores, add [seq_num_edges] edge scores ->
    dict(p [|A|], aset [|A|, L] int32 state sequences, S, iters, status)."""
    ns = np.ascontiguousarray(num_states, dtype=np.int32)
    L, V = ns.shape[0], int(ns.sum())
    x = np.ascontiguousarray(x, dtype=np.float64)
    add = np.ascontiguousarray(add, dtype=np.float64)
    assert x.shape[0] == V and add.shape[0] == seq_num_edges(ns)
    cap = V + 2
    p = np.zeros(cap, dtype=np.float64)
    aset = np.zeros((cap, L), dtype=np.int32)
    S = np.zeros(cap * cap, dtype=np.float64)
    cnt, it, st = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    rc = _lib(which).smlvm_oracle_seq_fwd(L, _p(ns, _i32p), _p(x, _f64p), _p(add, _f64p), max_iter, cap, _p(p, _f64p),
                                          _p(aset, _i32p), _p(S, _f64p), ctypes.byref(cnt), ctypes.byref(it), ctypes.byref(st))
    if rc != 0:
        raise RuntimeError("sequence oracle failed rc=%d" % rc)
    m = cnt.value
    return dict(p=p[:m].copy(), aset=aset[:m].copy(), S=S[:m * m].reshape(m, m).copy(), iters=it.value, status=st.value)


def seq_bwd(x, add, num_states, dp, max_iter=100, which="oracle"):
    """Re-solves, then dist_jacobian_vec: (dx [V], dadd [E])."""
    ns = np.ascontiguousarray(num_states, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    add = np.ascontiguousarray(add, dtype=np.float64)
    dp = np.ascontiguousarray(dp, dtype=np.float64)
    dx, dt = np.zeros(x.shape[0]), np.zeros(add.shape[0])
    rc = _lib(which).smlvm_oracle_seq_fwd_bwd(ns.shape[0], _p(ns, _i32p), _p(x, _f64p), _p(add, _f64p), max_iter,
                                              _p(dp, _f64p), dp.shape[0], _p(dx, _f64p), _p(dt, _f64p))
    if rc != 0:
        raise RuntimeError("sequence oracle bwd failed rc=%d" % rc)
    return dx, dt


def seq_map(x, add, num_states, which="oracle"):
    """Viterbi of the general chain alone: (y [L] int32, value)."""
    ns = np.ascontiguousarray(num_states, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    add = np.ascontiguousarray(add, dtype=np.float64)
    y = np.zeros(ns.shape[0], dtype=np.int32)
    val = ctypes.c_double(0)
    _lib(which).smlvm_oracle_seq_map(ns.shape[0], _p(ns, _i32p), _p(x, _f64p), _p(add, _f64p), _p(y, _i32p), ctypes.byref(val))
    return y, val.value


def sparsemap_batch_counts(x, kind=KIND_BERNOULLI, budget=0, max_iter=300, which="oracle"):
    """Tight C++ loop over rows (CPU-baseline timing).  Returns (counts, iters)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, n = x.shape
    counts = np.zeros(rows, dtype=np.int32)
    iters = np.zeros(rows, dtype=np.int32)
    _lib(which).smlvm_oracle_sparsemap_batch(
        kind, n, budget, _p(x, _f32p), rows, max_iter,
        counts.ctypes.data_as(_intp), iters.ctypes.data_as(_intp))
    return counts, iters


# closed-form marginals (restatement-independent known answers)
def bernoulli_marginals(x):
    return np.clip((np.asarray(x, dtype=np.float64) + 1.0) / 2.0, 0.0, 1.0)


def budget_marginals(x, budget):
    """Projection of (x+1)/2 onto {0<=u<=1, sum u <= budget} (bisection on the multiplier)."""
    a = (np.asarray(x, dtype=np.float64) + 1.0) / 2.0
    u = np.clip(a, 0, 1)
    if u.sum() <= budget:
        return u
    lo, hi = 0.0, float(a.max())
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if np.clip(a - mid, 0, 1).sum() > budget:
            lo = mid
        else:
            hi = mid
    return np.clip(a - hi, 0, 1)
