"""GPU parity: k-best bit-vectors (bit-exact vs the CPU oracle / the reference build)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,n,k", [(64, 32, 10), (64, 128, 10), (16, 1, 2), (16, 2, 4), (16, 3, 8),
                                      (50, 5, 4), (40, 12, 16), (30, 33, 17), (20, 64, 32), (10, 200, 7),
                                      (6, 512, 10), (6, 512, 32), (1000, 40, 3), (5, 31, 2)])
def test_topk_matches_oracle(oracle, rows, n, k):
    from lvmhelpers import ops
    rng = np.random.default_rng(n * 100 + k)
    s = rng.standard_normal((rows, n)).astype(np.float32)
    ref, ref_sc = oracle.topk(s, k, return_scores=True)
    cfg, bits, sc = ops.topk_bits(torch.from_numpy(s).cuda(), k, want_scores=True)
    assert np.array_equal(cfg.cpu().numpy(), ref)
    assert np.array_equal(sc.cpu().numpy(), ref_sc)
    # packed bits agree with the fp32 configs
    b = bits.cpu().numpy().astype(np.uint32)
    idx = np.arange(n)
    unpacked = ((b[:, :, idx // 32] >> (idx % 32)) & 1).astype(np.float32)
    assert np.array_equal(unpacked, ref)
    if oracle.have_ref("ref_topk"):
        assert np.array_equal(cfg.cpu().numpy(), oracle.topk(s, k, which="ref_topk"))


@pytest.mark.parametrize("rows,n,k", [(16384 + 5, 128, 10), (20000, 32, 10), (16400, 100, 16), (16390, 7, 5),
                                      (16384, 1, 2), (16384, 64, 2), (16384, 33, 16)])
def test_topk_tile_kernel(oracle, rows, n, k):
    """rows >= 16384, k <= 16, n <= 128: the thread-per-row kernel (topk_tile.cu), with ties
    (quantised rows), a partial last tile and all three word widths."""
    from lvmhelpers import ops
    rng = np.random.default_rng(n * 31 + k)
    s = rng.standard_normal((rows, n)).astype(np.float32)
    s[::7] = np.round(s[::7] * 2) / 2
    s[5] = 0.0
    ref, ref_sc = oracle.topk(s, k, return_scores=True)
    cfg, bits, sc = ops.topk_bits(torch.from_numpy(s).cuda(), k, want_scores=True)
    assert np.array_equal(cfg.cpu().numpy(), ref)
    assert np.array_equal(sc.cpu().numpy(), ref_sc)
    b = bits.cpu().numpy().astype(np.uint32)
    idx = np.arange(n)
    assert np.array_equal(((b[:, :, idx // 32] >> (idx % 32)) & 1).astype(np.float32), ref)
    # bits-only / configs-only calls give the same answers
    _, bits2, _ = ops.topk_bits(torch.from_numpy(s).cuda(), k, want_configs=False)
    assert torch.equal(bits2, bits)


def test_topk_ties_and_zeros(oracle):
    """Quantised scores: many exact ties and exact zeros exercise the strict-> rule."""
    from lvmhelpers import ops
    rng = np.random.default_rng(5)
    for n, k in [(20, 10), (8, 16), (40, 32)]:
        s = rng.integers(-3, 4, (300, n)).astype(np.float32)
        ref = oracle.topk(s, k)
        cfg, _, _ = ops.topk_bits(torch.from_numpy(s).cuda(), k)
        assert np.array_equal(cfg.cpu().numpy(), ref)
    s = np.zeros((4, 10), dtype=np.float32)
    cfg, _, _ = ops.topk_bits(torch.from_numpy(s).cuda(), 5)
    assert np.array_equal(cfg.cpu().numpy(), oracle.topk(s, 5))


def test_topk_bruteforce_small(oracle):
    from lvmhelpers import ops
    rng = np.random.default_rng(9)
    s = rng.standard_normal((20, 10)).astype(np.float32)
    cfg, _, sc = ops.topk_bits(torch.from_numpy(s).cuda(), 12, want_scores=True)
    cfg = cfg.cpu().numpy()
    for r in range(20):
        best, best_bits = oracle.topk_bruteforce(s[r], 12)
        assert np.allclose(sc[r].cpu().numpy(), best, atol=1e-5)
        assert np.array_equal(cfg[r], best_bits.astype(np.float32))
        assert np.array_equal(cfg[r, 0], (s[r] > 0).astype(np.float32))


def test_pbinary_topk_api(oracle):
    """Drop-in module: in-place fill of a host buffer, returns None (pbinary_topk.pyx:23-34)."""
    from lvmhelpers.pbinary_topk import batched_topk, binary_topk
    s = np.random.default_rng(1).standard_normal((8, 32)).astype(np.float32)
    out = torch.empty((8, 10, 32), dtype=torch.float32)
    assert batched_topk(s, out.numpy(), 10) is None
    assert np.array_equal(out.numpy(), oracle.topk(s, 10))
    res = binary_topk(list(s[0]), 3)
    assert len(res) == 3 and res[0][1] == [int(v > 0) for v in s[0]]
    with pytest.raises(ValueError):
        batched_topk(s, out.numpy(), 1)
    with pytest.raises(ValueError):
        batched_topk(s.astype(np.float64), out.numpy(), 10)


def test_bits_score_fwd_bwd():
    from lvmhelpers import ops
    x = torch.randn(64, 32, device="cuda", requires_grad=True)
    cfg, bits, _ = ops.topk_bits(x.detach(), 10)
    sk = ops.bits_score(x, bits)
    w = torch.randn(64, 10, device="cuda")
    (sk * w).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    ref = torch.einsum("bkj,bj->bk", cfg.double(), xr.double())  # structmarg.py:47
    (ref * w.double()).sum().backward()
    # kernel accumulates in fp64 and rounds once; einsum in fp32 would differ by ~1e-6*|score|
    assert (sk.double() - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()
    assert (x.grad.double() - xr.grad.double()).abs().max().item() <= 1e-5


@pytest.mark.parametrize("rows,n,k", [(2048, 128, 10), (1030, 32, 16), (1500, 100, 2), (4096, 4, 7), (1024, 64, 13),
                                      (1100, 36, 3), (2000, 128, 17), (1200, 130, 4)])
def test_bits_score_large_shapes(rows, n, k):
    """Re-scoring einsum from packed bits at larger shapes, top-k patterns and arbitrary ones: fp64 sums
    rounded once, so equal to the fp64 einsum to the last fp32 bit (1 ulp allowed for double rounding)."""
    from lvmhelpers import ops
    g = torch.Generator().manual_seed(n * 17 + k)
    x = (torch.randn(rows, n, generator=g) * 3).cuda()
    k_eff = min(k, 2 ** min(n, 20))
    cfg, bits, _ = ops.topk_bits(x, k_eff)
    sk = ops.bits_score(x, bits)
    ref = torch.einsum("bkj,bj->bk", cfg.double(), x.double())
    def close(a, b):
        return bool(((a.double() - b).abs() <= 1.2e-7 * b.abs().clamp(min=1e-30)).all())

    assert close(sk, ref)
    # arbitrary (not top-k) bit patterns too
    rb = torch.randint(-2 ** 31, 2 ** 31 - 1, bits.shape, generator=g, dtype=torch.int64).to(torch.int32).cuda()
    W = (n + 31) // 32
    if n % 32:
        rb[..., W - 1] &= (1 << (n % 32)) - 1
    idx = torch.arange(n, device="cuda")
    dense = ((rb[..., idx // 32] >> (idx % 32)) & 1).double()
    assert close(ops.bits_score(x, rb), torch.einsum("bkj,bj->bk", dense, x.double()))


@pytest.mark.parametrize("rows,n,k", [(1000, 128, 10), (333, 64, 5), (200, 32, 10)])
def test_bits_score_non_finite_scores(rows, n, k):
    """Masked (-inf) and +inf scores: the re-scoring that starts from configuration 0's sum must fall back to
    the plain sum when that base is not finite, and -inf entries outside every configuration must not leak."""
    from lvmhelpers import ops
    g = torch.Generator().manual_seed(n + k)
    x = torch.randn(rows, n, generator=g) * 2
    x[torch.rand(rows, n, generator=g) < 0.1] = float("-inf")
    x[::7, 3] = float("inf")
    x = x.cuda()
    _, bits, _ = ops.topk_bits(x, k, want_configs=False)
    rb = torch.randint(-2 ** 31, 2 ** 31 - 1, bits.shape, generator=g, dtype=torch.int64).to(torch.int32).cuda()
    idx = torch.arange(n, device="cuda")
    for b in (bits, rb):
        dense = ((b[..., idx // 32] >> (idx % 32)) & 1).bool()
        ref = torch.where(dense, x.double()[:, None, :], torch.zeros((), device="cuda", dtype=torch.float64)).sum(-1)
        got = ops.bits_score(x, b).double()
        fin = torch.isfinite(ref)
        assert torch.equal(torch.isnan(got), torch.isnan(ref))
        assert torch.equal(got[~fin & ~torch.isnan(ref)], ref[~fin & ~torch.isnan(ref)])
        assert bool(((got[fin] - ref[fin]).abs() <= 1.2e-7 * ref[fin].abs().clamp(min=1e-30)).all())


@pytest.mark.parametrize("rows,n,k", [(300, 128, 10), (64, 32, 10), (50, 33, 4), (20, 7, 5), (10, 200, 3)])
def test_bits_expand_matches_dense_configs(rows, n, k):
    """smlvm_bits_expand: packed candidates -> fp32 rows, all of them and an arbitrary selection
    (`bit_vector_z.view(-1, n)[mask]`, structmarg.py:116-126)."""
    from lvmhelpers import ops
    g = torch.Generator().manual_seed(n + k)
    x = torch.randn(rows, n, generator=g).cuda()
    cfg, bits, _ = ops.topk_bits(x, k)
    assert torch.equal(ops.bits_expand(bits, None, n).view(rows, k, n), cfg)
    sel = torch.randperm(rows * k, generator=g)[: rows * k // 3].cuda()
    assert torch.equal(ops.bits_expand(bits, sel, n), cfg.view(-1, n).index_select(0, sel))
    assert ops.bits_expand(bits, sel[:0], n).shape == (0, n)


def test_topk_marg_packed_candidates_equal_dense_path():
    """TopKSparsemaxMarg hands the k candidates over packed and expands only the support; the result
    (loss, logs, gradients) is the one of the dense [B,k,n] path; called on its own the wrapper still
    returns the reference's dense tensor."""
    from lvmhelpers.structmarg import TopKSparsemaxWrapper, TopKSparsemaxMarg, PackedBitVectors

    class Dec(torch.nn.Module):
        def __init__(self, n):
            super().__init__()
            self.lin = torch.nn.Linear(n, 16)
            self.seen = None

        def forward(self, z, decoder_input):
            self.seen = z
            return self.lin(z)

    def loss_fun(enc_in, z, dec_in, dec_out, labels):
        return ((dec_out - labels) ** 2).sum(-1), {"acc": dec_out.mean(-1)}

    torch.manual_seed(3)
    n, B = 40, 96
    enc = torch.nn.Linear(24, n).cuda()
    dec = Dec(n).cuda()
    x = torch.randn(B, 24).cuda()
    labels = torch.randn(B, 16).cuda()
    wrapper = TopKSparsemaxWrapper(enc, k=7)
    marg = TopKSparsemaxMarg(wrapper, dec, loss_fun, encoder_entropy_coeff=0.5)
    res = marg(x, x, labels)
    res["loss"].backward()
    g_packed = [p.grad.clone() for p in list(enc.parameters()) + list(dec.parameters())]
    z_packed = dec.seen.clone()

    # the wrapper on its own: dense tensor, as in the reference
    z_dense, distr, _ = wrapper(x)
    assert isinstance(z_dense, torch.Tensor) and z_dense.shape == (B, 7, n) and not wrapper.packed_output
    mask = distr.reshape(-1) > 0
    assert torch.equal(z_packed, z_dense.view(-1, n)[mask])

    # dense path through the same Marg: a wrapper subclass the Marg does not recognise
    class Opaque(torch.nn.Module):
        def __init__(self, w):
            super().__init__()
            self.w = w

        def forward(self, *a):
            return self.w(*a)

    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = None
    marg2 = TopKSparsemaxMarg(Opaque(wrapper), dec, loss_fun, encoder_entropy_coeff=0.5)
    res2 = marg2(x, x, labels)
    res2["loss"].backward()
    assert not isinstance(dec.seen, PackedBitVectors) and torch.equal(dec.seen, z_packed)
    assert torch.equal(res["loss"], res2["loss"])
    for a, b in zip(g_packed, [p.grad for p in list(enc.parameters()) + list(dec.parameters())]):
        assert torch.equal(a, b)
    assert torch.equal(res["log"]["loss_output"], res2["log"]["loss_output"])


@pytest.mark.parametrize("rows,k,scale", [(64, 10, 1.0), (5000, 10, 0.1), (5000, 16, 3.0), (333, 2, 1.0), (40000, 7, 0.3),
                                          (1, 3, 1.0)])
def test_ksparsemax_fwd_matches_oracle(oracle, rows, k, scale):
    """smlvm_ksparsemax_fwd: sparsemax over the k <= 16 re-scored candidates (structmarg.py:48), bit-equal to the
    oracle's restatement of entmax.sparsemax, with supp / nnz / entropy over the support (:51-58)."""
    from lvmhelpers import ops, _native as N
    g = torch.Generator().manual_seed(rows + k)
    s = torch.randn(rows, k, generator=g) * scale
    s[::5] = torch.round(s[::5] * 4) / 4          # exact ties
    if rows > 10:
        s[3] = 0.0                                # constant row
        s[4, 0] = s[4, 1:].max() + 1.0            # gap of exactly 1
        s[7, 1] = float("-inf")
    P_ref, k_ref = oracle.sparsemax_fwd(s)
    sk = s.cuda()
    distr = torch.empty_like(sk)
    supp = torch.empty(rows, dtype=torch.int32, device="cuda")
    nnz = torch.empty(rows, dtype=torch.int32, device="cuda")
    rent = torch.empty(rows, dtype=torch.float32, device="cuda")
    tot = torch.empty((), dtype=torch.float32, device="cuda")
    ws = N.reduce_workspace(sk.device)
    N.check(N.lib.smlvm_ksparsemax_fwd(N.ptr(sk), N.ptr(distr), N.ptr(supp), N.ptr(nnz), N.ptr(rent), N.ptr(tot), 1.0 / rows,
                                       rows, k, N.ptr(ws), ws.numel(), N.stream()), "smlvm_ksparsemax_fwd")
    assert torch.equal(distr.cpu(), P_ref)
    assert torch.equal(supp.cpu().long(), k_ref)
    assert torch.equal(nnz.cpu().long(), (P_ref > 0).sum(-1))
    pd = P_ref.double()
    h_ref = -torch.where(pd > 0, pd * pd.clamp(min=1e-300).log(), torch.zeros_like(pd)).sum(-1)
    assert (rent.cpu().double() - h_ref).abs().max().item() < 2e-6
    assert abs(tot.item() - h_ref.sum().item() / rows) < 1e-5 * max(1.0, abs(h_ref.sum().item() / rows))


@pytest.mark.parametrize("rows,n,k", [(64, 32, 10), (3000, 128, 10), (500, 40, 16), (20000, 64, 4), (100, 8, 2), (70, 96, 5)])
def test_topk_sparsemax_fused_equals_composed(rows, n, k):
    """ops.topk_sparsemax (k-best -> re-score -> k-sparsemax + entropy; backward in one launch) against the separate
    kernels composed with autograd: same candidates, same distribution bits, gradients within 1e-6."""
    from lvmhelpers import ops
    g = torch.Generator().manual_seed(n * 7 + k)
    x0 = (torch.randn(rows, n, generator=g) * 0.7).cuda()
    w = torch.randn(rows, k, generator=g).cuda()
    x1 = x0.clone().requires_grad_(True)
    distr, ent, bits, cfg, nnz, supp = ops.topk_sparsemax(x1, k, want_configs=True)
    (distr.mul(w).sum() + 0.37 * ent).backward()

    x2 = x0.clone().requires_grad_(True)
    cfg2, bits2, _ = ops.topk_bits(x2.detach(), k)
    assert torch.equal(bits, bits2) and torch.equal(cfg, cfg2)
    sk = ops.bits_score(x2, bits2)
    out = ops.sparsemax_fwd(sk.detach())
    assert torch.equal(distr.detach(), out["p"]) and torch.equal(supp, out["supp"]) and torch.equal(nnz, out["nnz"])
    # reference composition in torch autograd (fp64): einsum -> sparsemax Jacobian -> masked entropy
    P = out["p"].double()
    on = P > 0
    ent_ref = -(P[on] * P[on].log()).sum() / rows
    assert abs(ent.item() - ent_ref.item()) < 1e-5 * max(1.0, abs(ent_ref.item()))
    gd = w.double() - 0.37 / rows * torch.where(on, P.clamp(min=1e-300).log() + 1.0, torch.zeros_like(P))
    gd = torch.where(on, gd, torch.zeros_like(gd))
    ds = torch.where(on, gd - gd.sum(-1, keepdim=True) / on.sum(-1, keepdim=True), torch.zeros_like(gd))
    dx_ref = torch.einsum("bk,bkj->bj", ds, cfg2.double())
    tol = 1e-6 * max(1.0, dx_ref.abs().max().item())
    assert (x1.grad.double() - dx_ref).abs().max().item() <= tol
    # only one of the two gradient sources
    x3 = x0.clone().requires_grad_(True)
    d3, e3, *_ = ops.topk_sparsemax(x3, k)
    e3.backward()
    gd = torch.where(on, -(1.0 / rows) * (P.clamp(min=1e-300).log() + 1.0), torch.zeros_like(P))
    ds = torch.where(on, gd - gd.sum(-1, keepdim=True) / on.sum(-1, keepdim=True), torch.zeros_like(gd))
    assert (x3.grad.double() - torch.einsum("bk,bkj->bj", ds, cfg2.double())).abs().max().item() <= 1e-6


_TOPK_PATH_SCRIPT = r"""
import sys, numpy as np, torch
sys.path[:0] = [%r, %r]
import oracle
from lvmhelpers import ops
bad = []
for rows, n, k in [(16384 + 37, 128, 10), (16400, 128, 16), (16384, 100, 8), (16390, 96, 9), (16384, 64, 8), (16400, 33, 16),
                   (16384, 29, 10), (16384, 57, 12), (16390, 7, 5), (300, 128, 10), (16384, 128, 2)]:
    rng = np.random.default_rng(n * 131 + k)
    s = rng.standard_normal((rows, n)).astype(np.float32)
    s[::7] = np.round(s[::7] * 2) / 2
    s[5] = 0.0
    s[6, ::2] = 0.0
    chk = min(rows, 3000)
    ref, ref_sc = oracle.topk(s[:chk], k, return_scores=True)
    cfg, bits, sc = ops.topk_bits(torch.from_numpy(s).cuda(), k, want_scores=True)
    b = bits[:chk].cpu().numpy().astype(np.uint32)
    idx = np.arange(n)
    ok = np.array_equal(cfg[:chk].cpu().numpy(), ref) and np.array_equal(sc[:chk].cpu().numpy(), ref_sc) \
        and np.array_equal(((b[:, :, idx // 32] >> (idx %% 32)) & 1).astype(np.float32), ref)
    _, bits2, _ = ops.topk_bits(torch.from_numpy(s).cuda(), k, want_configs=False)
    ok = ok and torch.equal(bits2, bits)
    if not ok:
        bad.append((rows, n, k))
print("BAD", bad) if bad else print("ALL-OK")
"""


@pytest.mark.parametrize("path", ["reg", "tile", "blk"])
def test_every_topk_path_in_a_subprocess(path):
    """SMLVM_TOPK pins the k-best search to ONE of its three kernels (group-per-row, thread-per-row with the
    configurations riding along, thread-per-row with block-local configurations materialised every 28 bits): each must
    be bit-exact against the oracle, block boundaries (n = 29, 57, 96, 100, 128) and partial tiles included."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SMLVM_TOPK=path)
    script = _TOPK_PATH_SCRIPT % (root, os.path.join(root, "sparse-marginalization-lvm_b200"))
    res = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "ALL-OK" in res.stdout, res.stdout[-2000:]
