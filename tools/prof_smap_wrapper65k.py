"""ncu target: one SparseMAP wrapper step (solve with the arena of inverses, expand, backward) at 65536 x 128
(development aid).  argv: kind budget"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops
kind, budget = int(sys.argv[1]), int(sys.argv[2])
rows = 65536
g = torch.Generator().manual_seed(0)
xs = torch.randn(rows, 128, generator=g).cuda().requires_grad_(True)
for _ in range(2):
    xs.grad = None
    p, aset, aux = ops.sparsemap_batch(xs, kind, budget=budget, max_iter=300, only_positive=True)
    (ops.ragged_loss(p, torch.ones_like(p), 1.0 / rows) - ops.ragged_entropy(p, 1.0 / rows)).backward()
torch.cuda.synchronize()
print("ok", p.numel() / rows)
