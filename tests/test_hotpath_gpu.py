"""GPU: the pre-allocated pipelines of lvmhelpers.hotpath (what bench.py times) against the
oracle and against each other (resident step vs host-pipelined micro-batches)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(rows, K, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(rows, K, generator=g), torch.rand(rows, K, generator=g) * 3


@pytest.mark.parametrize("rows,K", [(64, 10), (4096, 128), (8192, 256)])
def test_resident_step_matches_oracle(oracle, rows, K):
    from lvmhelpers.hotpath import SparsemaxMargStep
    x, L = _inputs(rows, K, 5)
    xd, Ld = x.cuda(), L.cuda()
    nnz = SparsemaxMargStep.size_capacity(xd)
    step = SparsemaxMargStep(rows, K, xd.device, nnz + 16, entropy_coeff=0.5)
    total = step.run(xd, Ld)
    torch.cuda.synchronize()
    xr = x.clone().requires_grad_(True)
    Lr = L.clone().requires_grad_(True)
    P = oracle.sparsemax(xr)
    obj = (P * Lr).sum(-1).mean() - 0.5 * oracle.entropy(P).mean()
    obj.backward()
    assert torch.equal(step.p.cpu(), P.detach())
    assert abs(total.item() - (P * L).sum(-1).mean().item()) <= 1e-5
    nz = torch.nonzero(P.detach() > 0)
    n = nz.shape[0]
    assert int(step.offsets[-1]) == n
    assert torch.equal(step.ri[:n].cpu().long(), nz[:, 0]) and torch.equal(step.ci[:n].cpu().long(), nz[:, 1])
    assert (step.dx.cpu() - xr.grad).abs().max().item() <= 1e-6 * max(1.0, xr.grad.abs().max().item()) * 4
    assert (step.dloss.cpu() - Lr.grad).abs().max().item() <= 1e-7


@pytest.mark.parametrize("rows,K,scale", [(64, 256, 0.1), (4099, 128, 0.1), (4096, 128, 0.03), (4100, 64, 0.3), (5000, 256, 0.1),
                                          (4097, 512, 0.05), (300, 36, 1.0), (4096, 100, 0.1)])
@pytest.mark.parametrize("fused", [True, False])
def test_wide_support_step_matches_oracle(oracle, rows, K, scale, fused):
    """Flat score rows (supports of 10-60 entries: no support slots): forward by the cooperative kernel, then loss,
    lists and gradients from ONE dense read of P and L (smlvm_marg_step_fused with slots = NULL) -- against the oracle
    and against the three separate dense-read kernels (fused=False)."""
    from lvmhelpers.hotpath import SparsemaxMargStep
    g = torch.Generator().manual_seed(rows + K)
    x, L = torch.randn(rows, K, generator=g) * scale, torch.rand(rows, K, generator=g) * 3
    x[1] = 0.0                                  # constant row: full support
    x[2, 1:] = x[2, 0] - 2.0                    # one-hot row: p = 1, upper clamp of the entropy
    xd, Ld = x.cuda(), L.cuda()
    nnz = SparsemaxMargStep.size_capacity(xd)
    step = SparsemaxMargStep(rows, K, xd.device, nnz + 16, entropy_coeff=0.5, use_slots=False, fused=fused)
    assert step.fused == fused
    total = step.run(xd, Ld)
    torch.cuda.synchronize()
    xr, Lr = x.clone().requires_grad_(True), L.clone().requires_grad_(True)
    P = oracle.sparsemax(xr)
    ((P * Lr).sum(-1).mean() - 0.5 * oracle.entropy(P).mean()).backward()
    assert torch.equal(step.p.cpu(), P.detach())
    assert abs(total.item() - (P * L).sum(-1).mean().item()) <= 1e-5
    assert (step.row_loss.cpu() - (P.detach() * L).sum(-1)).abs().max().item() <= 1e-5
    nz = torch.nonzero(P.detach() > 0)
    n = nz.shape[0]
    assert int(step.offsets[-1]) == n
    assert torch.equal(step.ri[:n].cpu().long(), nz[:, 0]) and torch.equal(step.ci[:n].cpu().long(), nz[:, 1])
    assert torch.equal(step.vals[:n].cpu(), P.detach()[P.detach() > 0])
    assert (step.dx.cpu() - xr.grad).abs().max().item() <= 1e-6 * max(1.0, xr.grad.abs().max().item()) * 4
    assert (step.dloss.cpu() - Lr.grad).abs().max().item() <= 1e-7
    # capacity: the lists are clamped, the overflow is reported, loss and gradients do not depend on them
    small = SparsemaxMargStep(rows, K, xd.device, n // 2, entropy_coeff=0.5, use_slots=False, fused=fused)
    guard = torch.full((64,), -7.0, device="cuda")
    small.vals = torch.cat([small.vals, guard])[: n // 2]   # (view of a larger buffer: writes past n // 2 would land in guard)
    t2 = small.run(xd, Ld)
    torch.cuda.synchronize()
    assert small.overflowed() and torch.equal(small.dx, step.dx) and t2.item() == total.item()
    assert torch.equal(small.ri.cpu().long(), nz[: n // 2, 0])


@pytest.mark.parametrize("zero_copy", [False, True])
@pytest.mark.parametrize("chunks", [1, 4, 8])
def test_host_pipelined_step_equals_resident(chunks, zero_copy):
    """zero_copy: the loss matrix is not uploaded, the fused kernel gathers it from the pinned host buffer."""
    from lvmhelpers.hotpath import SparsemaxMargStep, HostPipelinedMargStep
    rows, K = 8192 * 4, 128
    x, L = _inputs(rows, K, 9)
    xp, Lp = x.pin_memory(), L.pin_memory()
    dev = torch.device("cuda", 0)
    xd, Ld = xp.to(dev), Lp.to(dev)
    nnz = SparsemaxMargStep.size_capacity(xd)
    step = SparsemaxMargStep(rows, K, dev, nnz + 16, entropy_coeff=1.0)
    total = step.run(xd, Ld)
    pipe = HostPipelinedMargStep(rows, K, dev, capacity_per_row=nnz / rows * 1.2 + 1, entropy_coeff=1.0,
                                 chunks=chunks, zero_copy_losses=zero_copy)
    dx_host = torch.empty(rows, K).pin_memory()
    for _ in range(3):   # buffer rotation across calls
        pipe.run(xp, Lp, dx_host)
    torch.cuda.synchronize()
    assert torch.equal(dx_host, step.dx.cpu())
    assert abs(pipe.loss() - total.item()) <= 1e-5
    assert pipe.h2d_bytes == (1 if zero_copy else 2) * rows * K * 4


def test_losses_read_in_place_from_pinned_host_memory():
    """smlvm_marg_step_fused with `loss` in pinned host memory (include/smlvm.h): same loss, lists and gradients as with
    the matrix on the device, rows beyond the 8 support slots included (they read their loss row densely); pageable host
    memory and the paths that read the loss matrix densely refuse."""
    from lvmhelpers.hotpath import SparsemaxMargStep, HostPipelinedMargStep
    rows, K = 8192, 128
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(rows, K, generator=g)
    x[::97] *= 0.05                       # flat rows: supports far beyond the slots
    L = torch.rand(rows, K, generator=g) * 3
    xd, Ld, Lp = x.to(dev), L.to(dev), L.pin_memory()
    nnz = SparsemaxMargStep.size_capacity(xd)
    a = SparsemaxMargStep(rows, K, dev, nnz + 16, entropy_coeff=1.0, use_slots=True)
    b = SparsemaxMargStep(rows, K, dev, nnz + 16, entropy_coeff=1.0, use_slots=True)
    ta, tb = a.run(xd, Ld), b.run(xd, Lp)
    torch.cuda.synchronize()
    assert int((a.nnz > 8).sum()) > 0
    assert ta.item() == tb.item() and torch.equal(a.dx, b.dx) and torch.equal(a.dloss, b.dloss)
    assert torch.equal(a.vals[:nnz], b.vals[:nnz]) and torch.equal(a.ri[:nnz], b.ri[:nnz]) and torch.equal(a.ci[:nnz], b.ci[:nnz])
    with pytest.raises(RuntimeError):
        b.run(xd, L)                      # pageable host memory is not device-readable
    dense = SparsemaxMargStep(rows, K, dev, nnz + 16, entropy_coeff=1.0, use_slots=False)
    with pytest.raises(RuntimeError):
        dense.run(xd, Lp)                 # the dense-read step would stream the whole matrix over PCIe
    with pytest.raises(RuntimeError):
        HostPipelinedMargStep(rows, K, dev, capacity_per_row=40.0, zero_copy_losses=True)   # wide supports: no slots


@pytest.mark.parametrize("rows,K", [(64, 10), (64, 256), (8192, 128)])
def test_cuda_graph_replay_equals_eager(rows, K):
    from lvmhelpers.hotpath import SparsemaxMargStep
    x, L = _inputs(rows, K, 11)
    xd, Ld = x.cuda(), L.cuda()
    nnz = SparsemaxMargStep.size_capacity(xd)
    step = SparsemaxMargStep(rows, K, xd.device, nnz + 4096, entropy_coeff=1.0)
    step.run(xd, Ld)
    torch.cuda.synchronize()
    ref = (step.dx.clone(), step.dloss.clone(), step.total.clone(), step.ri[:nnz].clone())
    g = step.capture(xd, Ld)
    # new data in the SAME buffers, then replay: the graph must recompute from the buffers
    x2, L2 = _inputs(rows, K, 12)
    xd.copy_(x2.cuda()); Ld.copy_(L2.cuda())
    g.replay()
    torch.cuda.synchronize()
    assert not torch.equal(step.dx, ref[0])
    xd.copy_(x.cuda()); Ld.copy_(L.cuda())
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(step.dx, ref[0]) and torch.equal(step.dloss, ref[1])
    assert torch.equal(step.total, ref[2]) and torch.equal(step.ri[:nnz], ref[3])


@pytest.mark.parametrize("rows,K", [(64, 10), (4096, 128), (5000, 33)])
def test_compact_step_matches_oracle(oracle, rows, K):
    """Losses on the support only: same loss and dscores as the dense objective whose loss matrix
    carries the ragged values at the support positions (entries off the support never matter)."""
    from lvmhelpers.hotpath import SparsemaxMargStep, SparsemaxMargCompactStep
    x, L = _inputs(rows, K, 21)
    xd = x.cuda()
    nnz = SparsemaxMargStep.size_capacity(xd)
    step = SparsemaxMargCompactStep(rows, K, xd.device, nnz + 16, entropy_coeff=0.3)
    xr = x.clone().requires_grad_(True)
    P = oracle.sparsemax(xr)
    mask = P.detach() > 0
    Lr = L[mask].clone().requires_grad_(True)           # ragged losses, row-major support order
    total = step.run(xd, Lr.detach().cuda(), int(mask.sum()))
    torch.cuda.synchronize()
    obj = (P[mask] * Lr).sum() / rows - 0.3 * oracle.entropy(P).mean()
    obj.backward()
    assert abs(total.item() - (P[mask] * Lr).sum().item() / rows) <= 1e-5
    assert (step.base.dx.cpu() - xr.grad).abs().max().item() <= 4e-6 * max(1.0, xr.grad.abs().max().item())
    assert (step.dloss_ragged[:int(mask.sum())].cpu() - Lr.grad).abs().max().item() <= 1e-7
    assert (step.base.dx.cpu()[~mask] == 0).all()


def test_compaction_branch_on_side_stream_matches_serial():
    """rows >= 65536: scan + fill run on the step's side stream between two events; same results as the
    serial (instrumented) launch order, eager and as a CUDA-graph replay, repeated back to back."""
    from lvmhelpers import hotpath
    rows, K = hotpath.OVERLAP_MIN_ROWS, 32
    g = torch.Generator().manual_seed(9)
    x = torch.randn(rows, K, generator=g).cuda()
    L = torch.rand(rows, K, generator=g).cuda()
    nnz = hotpath.SparsemaxMargStep.size_capacity(x)
    serial = hotpath.SparsemaxMargStep(rows, K, x.device, nnz + 64, entropy_coeff=0.7, overlap_compaction=False,
                                       fused=False)
    forked = hotpath.SparsemaxMargStep(rows, K, x.device, nnz + 64, entropy_coeff=0.7, fused=False)
    assert serial.side is None and forked.side is not None and not forked.fused
    serial.run(x, L)
    names = ("p", "offsets", "dx", "dloss", "row_loss", "total")

    def same():
        torch.cuda.synchronize()
        for nme in names:
            assert torch.equal(getattr(serial, nme), getattr(forked, nme)), nme
        for nme in ("vals", "ri", "ci"):
            assert torch.equal(getattr(serial, nme)[:nnz], getattr(forked, nme)[:nnz]), nme

    for _ in range(3):
        forked.run(x, L)
    same()
    marks = []
    forked.run(x, L, rec=marks.append)          # instrumented: serial on the current stream
    assert marks == list(hotpath.SparsemaxMargStep.KERNELS) + [None]
    same()
    graph = forked.capture(x, L)
    for nme in names + ("vals", "ri", "ci"):
        getattr(forked, nme).zero_()
    graph.replay()
    graph.replay()
    same()


@pytest.mark.parametrize("rows,K,scale", [(8192, 128, 1.0), (8192, 64, 0.5), (4096 + 17, 32, 1.0), (65536, 128, 1.0)])
def test_fused_step_equals_separate_kernels(rows, K, scale):
    """smlvm_marg_step_fused (loss + lists + gradients in one launch behind the scan) against the three separate
    kernels: row losses, lists, dX and dL bit-equal (same arithmetic in the same order), the fp64-accumulated total
    to one fp32 ulp (another partition of the same sum); rows whose support does not fit the slots included."""
    from lvmhelpers import hotpath
    g = torch.Generator().manual_seed(rows + K)
    x = (torch.randn(rows, K, generator=g) * scale).cuda()
    x[5] = 0.0                      # a constant row: K non-zeros, does not fit the slots
    x[rows - 1, :20] = 3.0          # 20 tied maxima
    L = (torch.rand(rows, K, generator=g) * 3).cuda()
    nnz = hotpath.SparsemaxMargStep.size_capacity(x)
    a = hotpath.SparsemaxMargStep(rows, K, x.device, nnz + 64, entropy_coeff=0.7, use_slots=True, fused=False,
                                  overlap_compaction=False)
    b = hotpath.SparsemaxMargStep(rows, K, x.device, nnz + 64, entropy_coeff=0.7, use_slots=True)
    assert b.fused and not a.fused and b.kernels == ("sparsemax_fwd", "scan_counts", "marg_step_fused")
    a.run(x, L); b.run(x, L)
    torch.cuda.synchronize()
    for nme in ("p", "offsets", "dx", "dloss", "row_loss"):
        assert torch.equal(getattr(a, nme), getattr(b, nme)), nme
    for nme in ("vals", "ri", "ci"):
        assert torch.equal(getattr(a, nme)[:nnz], getattr(b, nme)[:nnz]), nme
    assert abs(a.total.item() - b.total.item()) <= 2e-7 * max(1.0, abs(a.total.item()))
    assert not b.overflowed()
    # as a CUDA-graph replay
    graph = b.capture(x, L)
    for nme in ("dx", "dloss", "row_loss", "vals", "ri", "ci"):
        getattr(b, nme).zero_()
    graph.replay(); graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(a.dx, b.dx) and torch.equal(a.ri[:nnz], b.ri[:nnz])


@pytest.mark.parametrize("fused", [False, True])
def test_capacity_overflow_is_clamped_and_reported(fused):
    """A batch with more non-zeros than the pre-allocated capacity: nothing is written past the lists (guard words
    behind them stay intact), loss and gradients are unaffected, overflowed() says so."""
    from lvmhelpers import hotpath
    rows, K = 8192, 128
    g = torch.Generator().manual_seed(4)
    x = torch.randn(rows, K, generator=g).cuda()
    L = (torch.rand(rows, K, generator=g) * 3).cuda()
    nnz = hotpath.SparsemaxMargStep.size_capacity(x)
    ref = hotpath.SparsemaxMargStep(rows, K, x.device, nnz + 64, entropy_coeff=1.0, use_slots=True, fused=fused)
    ref.run(x, L)
    cap = nnz // 2
    small = hotpath.SparsemaxMargStep(rows, K, x.device, cap + 256, entropy_coeff=1.0, use_slots=True, fused=fused)
    small.capacity = cap                      # the last 256 entries of each list are guard words
    for t in (small.vals, small.ri, small.ci):
        t[cap:] = 12345
    small.run(x, L)
    torch.cuda.synchronize()
    assert small.overflowed() and not ref.overflowed()
    for t in (small.vals, small.ri, small.ci):
        assert (t[cap:] == 12345).all()
    assert torch.equal(small.ri[:cap], ref.ri[:cap]) and torch.equal(small.ci[:cap], ref.ci[:cap])
    assert torch.equal(small.dx, ref.dx) and torch.equal(small.total, ref.total)
    # the ragged consumers never read past the lists either
    cs = hotpath.SparsemaxMargCompactStep(rows, K, x.device, cap + 256, entropy_coeff=1.0)
    cs.base.capacity = cap
    cs.run(x, torch.rand(cap + 256, device=x.device), nnz)      # caller claims nnz items: clamped to the capacity
    torch.cuda.synchronize()
    assert torch.isfinite(cs.base.dx).all()
