"""Wrapper-level SparseMAP step (ops.sparsemap_batch + ragged loss + entropy, forward + backward), CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
for kind, budget in ((0, 0), (1, 64)):
    xs = torch.randn(rows, n, generator=torch.Generator().manual_seed(1)).cuda().requires_grad_(True)

    def step():
        xs.grad = None
        p, aset, aux = ops.sparsemap_batch(xs, kind, budget=budget, max_iter=300, only_positive=True)
        loss = ops.ragged_loss(p, torch.ones_like(p), 1.0 / rows) - ops.ragged_entropy(p, 1.0 / rows)
        loss.backward()
        return aux

    for _ in range(2):
        aux = step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    st = aux["state"]
    print("wrapper step rows %d n %d kind %d budget %d: min %.2f ms = %.3e samples/s; arena %.2f GB used %.2f GB; status %s" % (
        rows, n, kind, budget, min(ts), rows / min(ts) * 1e3, st["S_capacity"] * 8 / 1e9, int(st["S_cursor"]) * 8 / 1e9,
        torch.bincount(st["status"], minlength=4).tolist()))
