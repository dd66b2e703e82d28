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


@pytest.mark.parametrize("chunks", [1, 4, 8])
def test_host_pipelined_step_equals_resident(chunks):
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
                                 chunks=chunks)
    dx_host = torch.empty(rows, K).pin_memory()
    for _ in range(3):   # buffer rotation across calls
        pipe.run(xp, Lp, dx_host)
    torch.cuda.synchronize()
    assert torch.equal(dx_host, step.dx.cpu())
    assert abs(pipe.loss() - total.item()) <= 1e-5


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
    serial = hotpath.SparsemaxMargStep(rows, K, x.device, nnz + 64, entropy_coeff=0.7, overlap_compaction=False)
    forked = hotpath.SparsemaxMargStep(rows, K, x.device, nnz + 64, entropy_coeff=0.7)
    assert serial.side is None and forked.side is not None
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
