"""CPU suite: the N>1 host logic (row sharding, global loss scale, bucketed gradient
all-reduce) on world_size-2 gloo.  The kernels are row-independent, so the only thing that
must hold across ranks is: per-shard losses scaled by 1/B_global SUM to the single-process
loss, and all-reduced parameter gradients EQUAL the single-process gradients.  The CPU
oracle's sparsemax stands in for the device kernel here (test infrastructure)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    from lvmhelpers.dist import shard_bounds, to_global_rows
    for rows in (0, 1, 7, 64, 1000, 1 << 20):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard_bounds(rows, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert torch.equal(to_global_rows(torch.tensor([0, 2]), 1, 2, 7), torch.tensor([4, 6]))
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 10))


def _loss(model, x, L, scale, oracle):
    P = oracle.sparsemax(model(x))
    return (P * L).sum() * scale


def _worker(rank, ws, port, rows, out_path):
    for p in (ROOT, os.path.join(ROOT, "sparse-marginalization-lvm_b200")):
        sys.path.insert(0, p)
    import oracle
    from lvmhelpers.dist import shard_bounds, loss_scale, GradAllReducer, allreduce_stats
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    td.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        g = torch.Generator().manual_seed(0)
        x, L = torch.randn(rows, 12, generator=g), torch.rand(rows, 10, generator=g) * 3
        a, b = shard_bounds(rows, rank, ws)
        model = _model(1)
        loss = _loss(model, x[a:b], L[a:b], loss_scale(rows), oracle)
        loss.backward()
        GradAllReducer(model.parameters(), bucket_bytes=512)()   # tiny buckets: several all-reduces
        stats = allreduce_stats({"loss": loss.detach(), "rows": b - a})
        if rank == 0:
            torch.save({"loss": stats["loss"], "rows": stats["rows"],
                        "grads": [p.grad.clone() for p in model.parameters()]}, out_path)
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("rows", [64, 37])
def test_sharded_loss_and_gradients_match_single_process(oracle, tmp_path, rows):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), rows, out), nprocs=2, join=True)
    got = torch.load(out)
    from lvmhelpers.dist import loss_scale
    g = torch.Generator().manual_seed(0)
    x, L = torch.randn(rows, 12, generator=g), torch.rand(rows, 10, generator=g) * 3
    model = _model(1)
    loss = _loss(model, x, L, loss_scale(rows), oracle)
    loss.backward()
    assert got["rows"] == rows
    assert abs(got["loss"] - loss.item()) < 1e-5
    for ga, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(ga, p.grad, atol=1e-6)


def test_grad_allreducer_single_process_is_identity():
    from lvmhelpers.dist import GradAllReducer, allreduce_stats
    m = _model(2)
    m(torch.randn(3, 12)).sum().backward()
    before = [p.grad.clone() for p in m.parameters()]
    GradAllReducer(m.parameters())()
    assert all(torch.equal(a, p.grad) for a, p in zip(before, m.parameters()))
    assert allreduce_stats({"a": 2.0}) == {"a": 2.0}


def _worker_mean_hooks(rank, ws, port, rows, out_path):
    """Local 1/B on every rank (what the reference-API wrappers hard-code), mean all-reduce launched from backward."""
    for p in (ROOT, os.path.join(ROOT, "sparse-marginalization-lvm_b200")):
        sys.path.insert(0, p)
    import oracle
    from lvmhelpers.dist import shard_bounds, GradAllReducer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    td.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        g = torch.Generator().manual_seed(0)
        x, L = torch.randn(rows, 12, generator=g), torch.rand(rows, 10, generator=g) * 3
        a, b = shard_bounds(rows, rank, ws)
        model = _model(1)
        red = GradAllReducer(model.parameters(), bucket_bytes=512, average=True).attach()
        launched = []
        for step in range(2):                       # hooks survive across steps
            model.zero_grad(set_to_none=True)
            loss = _loss(model, x[a:b], L[a:b], 1.0 / (b - a), oracle)
            loss.backward()
            launched.append(len(red._pending))      # every bucket was launched from inside backward()
            red.finish()
        if rank == 0:
            torch.save({"launched": launched, "buckets": len(red.buckets),
                        "grads": [p.grad.clone() for p in model.parameters()]}, out_path)
    finally:
        td.destroy_process_group()


def test_mean_allreduce_from_backward_hooks_matches_single_process(oracle, tmp_path):
    rows = 64                                        # equal shards: the mean of local-1/B gradients is the global one
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker_mean_hooks, args=(2, _free_port(), rows, out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["buckets"] >= 2 and got["launched"] == [got["buckets"]] * 2
    g = torch.Generator().manual_seed(0)
    x, L = torch.randn(rows, 12, generator=g), torch.rand(rows, 10, generator=g) * 3
    model = _model(1)
    _loss(model, x, L, 1.0 / rows, oracle).backward()
    for ga, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(ga, p.grad, atol=1e-6)
