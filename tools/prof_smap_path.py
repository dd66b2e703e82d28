"""Launch-list target: the SparseMAP wrapper path (Bernoulli factor) at the bench shape (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 0
budget = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if len(sys.argv) > 4:
    ops.SMAP_CHUNK_LANES = int(sys.argv[4])
g = torch.Generator().manual_seed(0)
xs = torch.randn(rows, 128, generator=g).cuda().requires_grad_(True)
for it in range(5):
    xs.grad = None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    p, aset, aux = ops.sparsemap_batch(xs, kind, budget=budget, max_iter=300, only_positive=True)
    w = torch.ones_like(p)
    loss = ops.ragged_loss(p, w, 1.0 / rows) - ops.ragged_entropy(p, 1.0 / rows)
    loss.backward()
    e1.record()
    torch.cuda.synchronize()
    print("step ms", e0.elapsed_time(e1))
print("ok", p.numel())
