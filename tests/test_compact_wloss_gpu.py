"""GPU parity: support compaction (bit-exact indices) and the weighted-loss reduce."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,K", [(64, 10), (64, 256), (1, 1), (5, 3), (100, 33), (4097, 128),
                                    (20000, 12), (3, 1000), (70000, 8)])
def test_compaction_matches_nonzero(rows, K):
    from lvmhelpers import ops
    x = torch.randn(rows, K, generator=torch.Generator().manual_seed(K))
    out = ops.sparsemax_fwd(x.cuda())
    P = out["p"]
    comp = ops.compact(P, out["nnz"])
    ref = torch.nonzero(P > 0)  # row-major, the order of boolean-mask indexing
    assert comp["total"] == ref.shape[0]
    assert torch.equal(comp["rows"].long(), ref[:, 0])
    assert torch.equal(comp["cols"].long(), ref[:, 1])
    assert torch.equal(comp["vals"], P[P > 0])
    # offsets = exclusive scan of the row counts
    cnt = (P > 0).sum(-1)
    assert torch.equal(comp["offsets"][1:], cnt.cumsum(0))
    assert comp["offsets"][0].item() == 0
    # count kernel alone
    assert torch.equal(ops.compact_count(P).long(), cnt)


@pytest.mark.parametrize("rows,K", [(4096, 32), (4100, 64), (4096 + 31, 128), (9000, 128), (4097, 256)])
@pytest.mark.parametrize("density", [0.0, 0.03, 0.3, 1.0])
def test_compaction_tile_kernel(rows, K, density):
    """rows >= 4096, K in {32..256}: the TMA-staged tile kernel (compact_tile.cu), arbitrary
    non-negative matrices including empty and fully dense rows and a partial last tile."""
    from lvmhelpers import ops
    g = torch.Generator().manual_seed(rows + K)
    P = torch.rand(rows, K, generator=g)
    P = torch.where(torch.rand(rows, K, generator=g) < density, P, torch.zeros(()))
    P[3] = 0.0
    P[4] = 1.0 / K
    P[rows - 1, K - 1] = 0.5
    P[rows - 1, 0] = -1.0     # negative entries are not part of the support
    P = P.cuda()
    comp = ops.compact(P)
    ref = torch.nonzero(P > 0)
    assert comp["total"] == ref.shape[0]
    assert torch.equal(comp["rows"].long(), ref[:, 0])
    assert torch.equal(comp["cols"].long(), ref[:, 1])
    assert torch.equal(comp["vals"], P[P > 0])


@pytest.mark.parametrize("rows,K", [(64, 10), (100, 33), (4096 + 5, 128), (5000, 64), (4097, 256), (3, 1000)])
@pytest.mark.parametrize("scale", [0.03, 0.3, 1.0])
def test_compaction_from_forward_slots(rows, K, scale):
    """The forward kernel's support slots (tile path, and register path + slot kernel) feed the
    fill: same (value, row, col) triples as the scan of P, bit-exact -- including rows with more
    than 8 non-zeros (small score scales), which the fill scans from P."""
    from lvmhelpers import ops
    x = torch.randn(rows, K, generator=torch.Generator().manual_seed(K + rows)) * scale
    out = ops.sparsemax_fwd(x.cuda(), want_slots=True)
    P = out["p"]
    ref = torch.nonzero(P > 0)
    comp = ops.compact(P, out["nnz"], slots=out["slots"])
    assert comp["total"] == ref.shape[0]
    assert torch.equal(comp["rows"].long(), ref[:, 0]) and torch.equal(comp["cols"].long(), ref[:, 1])
    assert torch.equal(comp["vals"], P[P > 0])


def test_compaction_edge_cases():
    from lvmhelpers import ops
    # empty rows, full rows, negative and zero entries
    P = torch.tensor([[0.0, 0.0, 0.0, 0.0], [0.25, 0.25, 0.25, 0.25], [0.0, 1.0, 0.0, 0.0],
                      [-1.0, 0.0, 0.5, 0.5]], device="cuda")
    comp = ops.compact(P)
    ref = torch.nonzero(P > 0)
    assert torch.equal(comp["rows"].long(), ref[:, 0]) and torch.equal(comp["cols"].long(), ref[:, 1])
    assert comp["offsets"].tolist() == [0, 0, 4, 5, 7]
    # all-zero matrix -> empty ragged batch
    comp = ops.compact(torch.zeros(10, 7, device="cuda"))
    assert comp["total"] == 0 and comp["vals"].numel() == 0


@pytest.mark.parametrize("rows", [1, 7, 2048, 2049, 100000, 1 << 20])
def test_scan(rows):
    from lvmhelpers import ops
    c = torch.randint(0, 130, (rows,), dtype=torch.int32, device="cuda")
    off = ops.scan_counts(c)
    ref = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), c.long().cumsum(0)])
    assert torch.equal(off, ref)


@pytest.mark.parametrize("rows,K", [(64, 10), (64, 256), (4097, 128), (3, 7), (50, 1000)])
def test_dense_weighted_loss(rows, K):
    from lvmhelpers import ops
    g = torch.Generator().manual_seed(rows)
    P = torch.rand(rows, K, generator=g)
    L = torch.rand(rows, K, generator=g) * 3
    row_loss, total = ops.wloss_dense_fwd(P.cuda(), L.cuda(), scale=1.0 / rows)
    ref_rows = P.unsqueeze(1).bmm(L.unsqueeze(-1)).squeeze()  # marg.py:125
    assert torch.allclose(row_loss.cpu(), ref_rows.reshape(-1), rtol=1e-5, atol=1e-5)
    assert abs(total.item() - ref_rows.double().mean().item()) <= 1e-5 * abs(ref_rows.mean().item()) + 1e-6


def test_ragged_loss_and_entropy_autograd():
    from lvmhelpers import ops
    N, B = 1000, 64
    g = torch.Generator().manual_seed(1)
    p = (torch.rand(N, generator=g) * 0.1 + 1e-4)
    l = torch.rand(N, generator=g) * 500 + 1500
    pr, lr = p.clone().requires_grad_(True), l.clone().requires_grad_(True)
    ref = (pr @ lr) / B + 0.3 * (-(pr @ torch.log(pr)) / B)     # structmarg.py:140, :54-58
    ref.backward()
    pc, lc = p.cuda().requires_grad_(True), l.cuda().requires_grad_(True)
    out = ops.ragged_loss(pc, lc, 1.0 / B) + 0.3 * ops.ragged_entropy(pc, 1.0 / B)
    out.backward()
    assert abs(out.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert torch.allclose(pc.grad.cpu(), pr.grad, rtol=1e-5, atol=1e-6)
    assert torch.allclose(lc.grad.cpu(), lr.grad, rtol=1e-5, atol=1e-8)
    # segmented row sums
    offsets = torch.tensor([0, 10, 10, 500, 1000], dtype=torch.int64, device="cuda")
    rows, _, _ = ops.wloss_ragged_fwd(p.cuda(), l.cuda(), offsets=offsets, want_rows=True, want_total=False)
    pl = (p * l).double()
    ref_rows = torch.stack([pl[0:10].sum(), pl[10:10].sum(), pl[10:500].sum(), pl[500:].sum()])
    assert torch.allclose(rows.cpu().double(), ref_rows, rtol=1e-5)


def test_segment_logsumexp_matches_reference_loop():
    """compute_importance_sampling's per-sample loop (experiments/bit_vector-vae/train.py:320-330)
    restated with torch ops vs the segmented kernel; empty segments give -inf."""
    import math
    from lvmhelpers import ops
    g = torch.Generator().manual_seed(3)
    sizes = torch.randint(0, 40, (500,), generator=g)
    sizes[7] = 0
    idxs = torch.repeat_interleave(torch.arange(500), sizes)
    loss_output = torch.rand(int(sizes.sum()), generator=g) * 500 + 1500
    logp_z = 32 * math.log(0.5)
    ref = torch.stack([torch.logsumexp(-loss_output[idxs == k] + logp_z, dim=0) for k in range(500)])
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), sizes.cumsum(0)]).cuda()
    out = ops.segment_logsumexp(loss_output.cuda(), offsets, sign=-1.0, shift=logp_z).cpu()
    assert torch.isinf(out[7]) and out[7] < 0
    fin = torch.isfinite(ref)
    assert torch.equal(fin, torch.isfinite(out))
    assert (out[fin] - ref[fin]).abs().max().item() <= 1e-3 * 1e-2 + 2e-4 * 1  # |v| ~ 2000: fp32 ulp 1.2e-4


@pytest.mark.parametrize("shape,dtype", [((300, 128), torch.float32), ((257, 784), torch.float32), ((100,), torch.int64),
                                         ((64, 3), torch.float32), ((50, 2, 6), torch.float32), ((40, 5), torch.int32),
                                         ((33, 7), torch.float16), ((10, 1), torch.uint8)])
def test_gather_rows_matches_index_select(shape, dtype):
    """smlvm_gather_rows (`encoder_input[idxs]`, structmarg.py:95-124, 236-240): every chunk width (16 / 8 / 4
    bytes), odd row sizes through the index_select route, repeated and unsorted ids, empty selection, and the
    gradient of a differentiable source."""
    from lvmhelpers import ops
    g = torch.Generator().manual_seed(shape[0])
    if dtype.is_floating_point:
        t = torch.randn(shape, generator=g).to(dtype).cuda()
    else:
        t = torch.randint(0, 100, shape, generator=g).to(dtype).cuda()
    idx = torch.randint(0, shape[0], (4 * shape[0] + 3,), generator=g).to(torch.int32).cuda()
    assert torch.equal(ops.gather_rows(t, idx), t.index_select(0, idx.long()))
    srt = torch.sort(idx).values
    assert torch.equal(ops.gather_rows(t, srt), t.index_select(0, srt.long()))
    assert ops.gather_rows(t, idx[:0]).shape == (0,) + tuple(shape[1:])
    assert torch.equal(ops.gather_rows(t, idx.long()), t.index_select(0, idx.long()))       # int64 ids accepted
    if dtype == torch.float32:
        a = t.clone().requires_grad_(True)
        b = t.clone().requires_grad_(True)
        w = torch.randn((idx.numel(),) + tuple(shape[1:]), generator=g).cuda()
        (ops.gather_rows(a, idx) * w).sum().backward()
        (b.index_select(0, idx.long()) * w).sum().backward()
        assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-5)
    assert ops.gather_rows(None, idx) is None and ops.gather_rows(3.5, idx) == 3.5
