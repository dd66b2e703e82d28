"""GPU parity: batched sparsemax fwd/bwd (through the C ABI) vs the CPU oracle.

Bar (BASELINE.json north_star): supports bit-exact, probabilities within 1e-6.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(64, 10), (64, 256), (64, 128), (1, 1), (3, 2), (7, 5), (33, 17), (5, 33), (130, 100),
          (9, 255), (9, 257), (6, 500), (5, 1000), (4, 1024), (3, 1500), (2, 4096), (2, 8192),
          (4097, 128), (1000, 12)]


def _inputs(rows, K, scale, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(rows, K, generator=g) * scale


@pytest.mark.parametrize("rows,K", SHAPES)
@pytest.mark.parametrize("scale", [0.1, 1.0, 3.0])
def test_forward_parity(oracle, rows, K, scale):
    from lvmhelpers import ops
    x = _inputs(rows, K, scale, 42 + K)
    P_ref, k_ref = oracle.sparsemax_fwd(x)
    out = ops.sparsemax_fwd(x.cuda(), want_entropy=True, want_argmax=True)
    P = out["p"].cpu()
    # supports bit-exact
    assert torch.equal(P > 0, P_ref > 0)
    assert torch.equal(out["supp"].cpu().long(), k_ref)
    assert torch.equal(out["nnz"].cpu().long(), (P_ref > 0).sum(-1))
    # probabilities within 1e-6 (in practice bit-identical: same fp32 operations)
    assert (P - P_ref).abs().max().item() <= 1e-6
    assert torch.equal(P, P_ref)
    # fused extras
    assert torch.equal(out["argmax"].cpu(), x.argmax(-1))
    H_ref = oracle.entropy(P_ref)
    assert torch.allclose(out["entropy"].cpu(), H_ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("rows,K", SHAPES)
def test_backward_parity(oracle, rows, K):
    from lvmhelpers import ops
    x = _inputs(rows, K, 1.0, 7 + K)
    g = _inputs(rows, K, 1.0, 8 + K)
    P_ref, k_ref = oracle.sparsemax_fwd(x)
    dx_ref = oracle.sparsemax_bwd(P_ref, k_ref, g)
    out = ops.sparsemax_fwd(x.cuda())
    dx = ops.sparsemax_bwd(out["p"], out["supp"], dp=g.cuda()).cpu()
    assert torch.equal(dx != 0, dx_ref != 0) or (dx - dx_ref).abs().max() <= 1e-6
    # fp32 sum over the support in a different order: tolerance 1e-6 * max|g| * sqrt(k)
    tol = 1e-6 * max(1.0, g.abs().max().item()) * max(1.0, float(k_ref.max()) ** 0.5)
    assert (dx - dx_ref).abs().max().item() <= tol


def test_edge_rows(oracle):
    from lvmhelpers import ops
    K = 16
    rows = [
        torch.zeros(K),                                   # constant row -> uniform
        torch.full((K,), 5.0),                            # constant, shifted
        torch.tensor([10.0] + [0.0] * (K - 1)),           # one-hot (gap >= 1)
        torch.tensor([1.0, 0.0] + [-5.0] * (K - 2)),      # gap exactly 1: second entry at tau
        torch.tensor([0.0, -1.0] * (K // 2)),             # many entries exactly at the threshold
        torch.tensor([0.5, 0.5] + [-1.0] * (K - 2)),      # tied maxima
        torch.tensor([float("-inf")] * (K - 1) + [0.0]),  # masked logits
        torch.linspace(-1e-3, 1e-3, K),                   # nearly constant
        torch.tensor([1e4, 1e4 - 0.5] + [0.0] * (K - 2)), # large magnitude
    ]
    x = torch.stack(rows)
    P_ref, k_ref = oracle.sparsemax_fwd(x)
    out = ops.sparsemax_fwd(x.cuda())
    assert torch.equal(out["p"].cpu(), P_ref)
    assert torch.equal(out["supp"].cpu().long(), k_ref)
    assert torch.allclose(out["p"].sum(-1).cpu(), torch.ones(len(rows)), atol=1e-6)


TILE_SHAPES = [(4096, 32), (4113, 64), (4096 + 31, 128), (8192, 128), (4097, 256), (5000, 512)]


@pytest.mark.parametrize("rows,K", TILE_SHAPES + [(4100, 100), (4096, 36), (4099, 300)])
@pytest.mark.parametrize("scale", [0.01, 0.1, 0.3, 1.0, 3.0])
@pytest.mark.parametrize("want_slots", [False, True])
def test_tile_kernel_parity(oracle, rows, K, scale, want_slots):
    """rows >= 4096: the two throughput kernels.  Without support slots the cooperative register kernel
    (sparsemax_coop.cu: G lanes per row, Michelot + checked reference tau, 32 < K <= 512); with them (K in
    {32, 64, 128}) the TMA-staged tile kernel (sparsemax_tile.cu).  Small scales give supports wider than the tile
    kernel's 32-entry candidate list (whole-row / whole-warp paths), large scales the per-thread path; a partial last
    tile, unaligned widths and structured rows are included.  Bit-exact against the oracle."""
    from lvmhelpers import ops
    x = _inputs(rows, K, scale, 1000 + K)
    # structured rows inside the batch: ties, one-hot with gap exactly 1, constants, masked
    x[5] = 0.0
    x[6] = torch.tensor([1.0] + [0.0] * (K - 1))
    x[7] = torch.round(x[7] * 2) / 2
    x[8, : K // 2] = float("-inf")
    x[9] = 1e4 * torch.sign(x[9])
    x[rows - 1] = torch.linspace(-1e-3, 1e-3, K)
    P_ref, k_ref = oracle.sparsemax_fwd(x)
    out = ops.sparsemax_fwd(x.cuda(), want_entropy=True, want_argmax=True, want_slots=want_slots)
    P = out["p"].cpu()
    assert torch.equal(P > 0, P_ref > 0)
    assert torch.equal(out["supp"].cpu().long(), k_ref)
    assert torch.equal(out["nnz"].cpu().long(), (P_ref > 0).sum(-1))
    assert (P - P_ref).abs().max().item() <= 1e-6
    assert torch.equal(P, P_ref)
    assert torch.equal(out["argmax"].cpu(), x.argmax(-1))
    assert torch.allclose(out["entropy"].cpu(), oracle.entropy(P_ref), rtol=1e-5, atol=1e-6)
    if want_slots:   # rows the slots describe: the support pairs are those of P
        sl = out["slots"].cpu()
        cols = sl[..., 1].contiguous().view(torch.int32)
        fits = cols[:, 0] != -2
        assert bool((((P_ref > 0).sum(-1) <= 8) | ~fits).all())
        for r in torch.nonzero(fits)[:64, 0].tolist():
            on = cols[r] >= 0
            got = sorted(zip(cols[r][on].tolist(), sl[r, :, 0][on].tolist()))
            ref = [(c, P_ref[r, c].item()) for c in torch.nonzero(P_ref[r] > 0)[:, 0].tolist()]
            assert got == ref


_PATH_SCRIPT = r"""
import sys, torch
sys.path[:0] = [%r, %r]
import oracle
from lvmhelpers import ops
bad = []
for rows, K in [(64, 10), (64, 256), (300, 128), (4099, 128), (4100, 64), (4097, 256), (130, 100), (4096, 32), (5000, 500)]:
    for scale in (0.03, 0.3, 3.0):
        g = torch.Generator().manual_seed(rows + K)
        x = torch.randn(rows, K, generator=g) * scale
        x[1] = 0.0
        x[2, 1:] = x[2, 0] - 1.0
        x[3] = torch.round(x[3] * 4) / 4
        x[4, ::2] = float("-inf")
        P_ref, k_ref = oracle.sparsemax_fwd(x)
        out = ops.sparsemax_fwd(x.cuda(), want_entropy=True, want_argmax=True)
        ok = torch.equal(out["p"].cpu(), P_ref) and torch.equal(out["supp"].cpu().long(), k_ref) \
            and torch.equal(out["nnz"].cpu().long(), (P_ref > 0).sum(-1)) and torch.equal(out["argmax"].cpu(), x.argmax(-1)) \
            and torch.allclose(out["entropy"].cpu(), oracle.entropy(P_ref), rtol=1e-5, atol=1e-6)
        if not ok:
            bad.append((rows, K, scale))
print("BAD", bad) if bad else print("ALL-OK")
"""


@pytest.mark.parametrize("path", ["reg", "tile", "coop"])
def test_every_forward_path_in_a_subprocess(path):
    """SMLVM_SPARSEMAX_FWD pins the forward to ONE of its three kernels (read once per process, hence a subprocess):
    each must be bit-exact against the oracle on every shape it accepts (shapes a pinned kernel does not cover fall
    through to the register kernel)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SMLVM_SPARSEMAX_FWD=path)
    script = _PATH_SCRIPT % (root, os.path.join(root, "sparse-marginalization-lvm_b200"))
    res = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "ALL-OK" in res.stdout, res.stdout[-2000:]


def test_tile_kernel_optional_outputs():
    from lvmhelpers import ops
    x = torch.randn(4096, 128, device="cuda")
    a = ops.sparsemax_fwd(x, want_entropy=True, want_argmax=True)
    b = ops.sparsemax_fwd(x, want_nnz=False)
    assert torch.equal(a["p"], b["p"]) and torch.equal(a["supp"], b["supp"]) and b["nnz"] is None


def test_properties_full_size():
    """BASELINE config-5 size: size-independent properties instead of an oracle run."""
    from lvmhelpers import ops
    rows, K = 1 << 18, 128
    x = torch.randn(rows, K, device="cuda")
    out = ops.sparsemax_fwd(x)
    P = out["p"]
    assert (P >= 0).all()
    assert (P.sum(-1) - 1).abs().max().item() < 1e-5
    assert torch.equal(out["nnz"].long(), (P > 0).sum(-1))
    # shift invariance (exact: the max is subtracted first; shifts by powers of two are exact)
    out2 = ops.sparsemax_fwd(x + 4.0)
    assert (out2["p"] - P).abs().max().item() < 1e-5
    # p_i > 0  <=>  x_i above the threshold: support = the top-k entries of the row
    kth = torch.topk(x, int(out["nnz"].max()), dim=-1).values
    thr = kth.gather(1, (out["nnz"].long() - 1).unsqueeze(1))
    assert torch.equal(P > 0, x >= thr)
    # backward: gradient sums to zero over the support, is zero elsewhere
    g = torch.randn_like(x)
    dx = ops.sparsemax_bwd(P, out["supp"], dp=g)
    assert (dx[P == 0] == 0).all()
    assert dx.sum(-1).abs().max().item() < 1e-3


def test_autograd_and_dims(oracle):
    from lvmhelpers import ops
    x = torch.randn(4, 6, 20)
    xg = x.cuda().requires_grad_(True)
    P = ops.sparsemax(xg, dim=-1)
    w = torch.randn(4, 6, 20)
    (P * w.cuda()).sum().backward()
    xr = x.clone().requires_grad_(True)
    Pr = oracle.sparsemax(xr)
    (Pr * w).sum().backward()
    assert torch.equal(P.detach().cpu(), Pr.detach())
    assert (xg.grad.cpu() - xr.grad).abs().max().item() <= 2e-6
    # other dim
    P1 = ops.sparsemax(x.cuda(), dim=1)
    Pr1 = oracle.sparsemax(x.transpose(1, 2)).transpose(1, 2)
    assert torch.equal(P1.cpu(), Pr1)


def test_fused_entropy_and_loss_backward(oracle):
    """smlvm_sparsemax_bwd with loss / entropy terms == autograd of the reference expressions."""
    from lvmhelpers import ops
    rows, K = 64, 10
    x = _inputs(rows, K, 1.0, 3)
    L = torch.rand(rows, K) * 200 + 50
    coeff = 0.7
    xr = x.clone().requires_grad_(True)
    Pr = oracle.sparsemax(xr)
    obj = (Pr * L).sum(-1).mean() - coeff * oracle.entropy(Pr).mean()
    obj.backward()
    out = ops.sparsemax_fwd(x.cuda())
    dent = torch.full((rows,), -coeff / rows, device="cuda")
    dx = ops.sparsemax_bwd(out["p"], out["supp"], loss=L.cuda(), loss_scale_host=1.0 / rows, dent=dent)
    assert (dx.cpu() - xr.grad).abs().max().item() <= 1e-5 * max(1.0, xr.grad.abs().max().item())


def test_rejects_bad_arguments():
    from lvmhelpers import ops
    from lvmhelpers._native import SmlvmError
    with pytest.raises(RuntimeError):
        ops.sparsemax_fwd(torch.randn(2, 3))  # CPU tensor: no CPU path
    with pytest.raises(SmlvmError):
        ops.sparsemax_fwd(torch.randn(1, 8193, device="cuda"))


@pytest.mark.parametrize("rows,K", [(100, 16), (64, 10), (4096, 128), (5001, 64), (4100, 36), (300, 512), (50, 1000)])
def test_ragged_backward_equals_dense_backward(rows, K):
    """smlvm_sparsemax_bwd_ragged (gradients on the support only: all three code paths -- TMA tile,
    staged scatter, thread-per-row scatter) == the dense Jacobian-vector product fed with the same
    gradient scattered into a dense matrix."""
    from lvmhelpers import ops
    g = torch.Generator().manual_seed(rows + K)
    x = torch.randn(rows, K, generator=g).cuda()
    out = ops.sparsemax_fwd(x, want_slots=True)
    comp = ops.compact(out["p"], out["nnz"], slots=out["slots"])
    N = comp["total"]
    dp_r = torch.randn(N, generator=g).cuda()
    dent = torch.randn(rows, generator=g).cuda()
    dense_dp = torch.zeros(rows, K, device="cuda")
    dense_dp[comp["rows"].long(), comp["cols"].long()] = dp_r
    ref = ops.sparsemax_bwd(out["p"], out["supp"], dp=dense_dp, dent=dent)
    got, dl = ops.sparsemax_bwd_ragged(comp, out["supp"], (rows, K), dp=dp_r, dent=dent, loss_scale_host=0.5,
                                       want_dloss=True)
    assert torch.equal(got != 0, ref != 0)
    assert (got - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())
    assert torch.equal(dl, 0.5 * comp["vals"])


@pytest.mark.parametrize("rows,K", [(4096, 128), (5001, 64), (4100, 36), (4096, 256), (8192, 32)])
@pytest.mark.parametrize("scale", [0.05, 0.5, 2.0])
def test_slot_driven_dense_kernels_equal_dense_reads(rows, K, scale):
    """smlvm_wloss_dense_fwd / smlvm_sparsemax_bwd with the forward's support slots (reading L and dP
    only on the support; rows with more than 8 non-zeros fall back to dense processing inside the
    kernel) == the same calls without slots."""
    from lvmhelpers import ops
    g = torch.Generator().manual_seed(rows + K)
    x = (torch.randn(rows, K, generator=g) * scale).cuda()
    L = (torch.rand(rows, K, generator=g) * 3).cuda()
    dP = torch.randn(rows, K, generator=g).cuda()
    dent = torch.randn(rows, generator=g).cuda()
    out = ops.sparsemax_fwd(x, want_slots=True)
    P, supp, slots = out["p"], out["supp"], out["slots"]
    r0, t0 = ops.wloss_dense_fwd(P, L, scale=1.0 / rows)
    r1, t1 = ops.wloss_dense_fwd(P, L, scale=1.0 / rows, slots=slots)
    assert (r0 - r1).abs().max().item() <= 2e-6 * max(1.0, r0.abs().max().item())
    assert abs(t0.item() - t1.item()) <= 1e-5 * max(1.0, abs(t0.item()))
    for kw in (dict(dp=dP), dict(loss=L, loss_scale_host=0.25, dent=dent), dict(dp=dP, loss=L, dent=dent)):
        a, al = ops.sparsemax_bwd(P, supp, want_dloss=True, **kw)
        b, bl = ops.sparsemax_bwd(P, supp, want_dloss=True, slots=slots, **kw)
        assert torch.equal(a != 0, b != 0)
        assert (a - b).abs().max().item() <= 4e-6 * max(1.0, a.abs().max().item())
        assert torch.equal(al, bl)
    # losses gathered once by the weighted-loss forward, streamed by the backward
    _, _, lsl = ops.wloss_dense_fwd(P, L, scale=1.0 / rows, slots=slots, want_loss_slots=True)
    d0 = ops.sparsemax_bwd(P, supp, loss=L, loss_scale_host=0.25, dent=dent)
    d1 = ops.sparsemax_bwd(P, supp, loss=L, loss_scale_host=0.25, dent=dent, slots=slots, loss_slots=lsl)
    assert (d0 - d1).abs().max().item() <= 4e-6 * max(1.0, d0.abs().max().item())
    c = ops.sparsemax_bwd(P, supp, dp=dP, slots=slots)          # without dloss
    assert (c - ops.sparsemax_bwd(P, supp, dp=dP)).abs().max().item() <= 4e-6 * max(1.0, dP.abs().max().item())
