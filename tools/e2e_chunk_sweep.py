"""Host-buffer step (zero-copy losses, dense gradient read back) at 1M x 128 over micro-batch count and buffer depth."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers.hotpath import SparsemaxMargStep, HostPipelinedMargStep
from tools.quick_bench import timeit
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
rows, K = 1 << 20, 128
x_pin = torch.randn(rows, K, generator=g).pin_memory()
l_pin = (torch.rand(rows, K, generator=g) * 3).pin_memory()
dx_pin = torch.empty(rows, K).pin_memory()
nnz = SparsemaxMargStep.size_capacity(x_pin.to(dev))
for chunks in (8, 16, 32, 64, 128):
    for nbuf in (2, 3, 4, 8):
        pipe = HostPipelinedMargStep(rows, K, dev, capacity_per_row=nnz / rows * 1.05 + 0.5, entropy_coeff=1.0, chunks=chunks, nbuf=nbuf, zero_copy_losses=True)
        ms = timeit(lambda: pipe.run(x_pin, l_pin, dx_pin), warm=2, reps=8)
        print("chunks %3d nbuf %d  %.2f ms/step  %.1f M samples/s" % (chunks, nbuf, ms, rows / ms / 1e3), flush=True)
        del pipe
