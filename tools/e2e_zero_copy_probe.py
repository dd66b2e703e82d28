"""Host-buffer step at 1M x 128 with the loss matrix uploaded vs read in place over PCIe on the support only
(HostPipelinedMargStep(zero_copy_losses=True)): step time, and the gradient / loss of both against the resident step."""
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
x, L = x_pin.to(dev), l_pin.to(dev)
nnz = SparsemaxMargStep.size_capacity(x)
step = SparsemaxMargStep(rows, K, dev, nnz + 1024, entropy_coeff=1.0)
step.run(x, L); torch.cuda.synchronize()
ref_loss = float(step.total)
for chunks in (8, 16, 32):
    for zc in (False, True):
        for want_dx in (True, False):
            pipe = HostPipelinedMargStep(rows, K, dev, capacity_per_row=nnz / rows * 1.05 + 0.5, entropy_coeff=1.0, chunks=chunks, zero_copy_losses=zc)
            dx_pin = torch.empty(rows, K).pin_memory() if want_dx else None
            ms = timeit(lambda: pipe.run(x_pin, l_pin, dx_pin), warm=2, reps=8)
            torch.cuda.synchronize()
            err = float((dx_pin.to(dev) - step.dx).abs().max()) if want_dx else float("nan")
            print("chunks %2d zero_copy_losses=%-5s dX back=%-5s  %.2f ms/step  %.1f M samples/s  max|dX - resident| %.2e  loss %.7f (resident %.7f)"
                  % (chunks, zc, want_dx, ms, rows / ms / 1e3, err, pipe.loss(), ref_loss), flush=True)
            del pipe, dx_pin
