"""ncu target: two passes of the contract step at 1M x 128, N(0,1) (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers.hotpath import SparsemaxMargStep
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
rows, K = 1 << 20, 128
x = torch.randn(rows, K, generator=g).to(dev)
L = (torch.rand(rows, K, generator=g) * 3).to(dev)
nnz = SparsemaxMargStep.size_capacity(x)
step = SparsemaxMargStep(rows, K, dev, nnz + 1024, entropy_coeff=1.0)
for _ in range(2):
    step.run(x, L)
torch.cuda.synchronize()
print("ok", step.fused, step.use_slots)
