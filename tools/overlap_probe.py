"""Resident categorical step with and without the compaction branch on a side stream (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers.hotpath import SparsemaxMargStep
rows, K = 1 << 20, 128
g = torch.Generator().manual_seed(42)
x = torch.randn(rows, K, generator=g).cuda()
L = (torch.rand(rows, K, generator=g) * 3).cuda()
nnz = SparsemaxMargStep.size_capacity(x)
for rep in range(2):
    for ov in (False, True):
        step = SparsemaxMargStep(rows, K, x.device, nnz + 1024, entropy_coeff=1.0, overlap_compaction=ov)
        for mode in ("eager", "graph"):
            fn = (lambda: step.run(x, L)) if mode == "eager" else step.capture(x, L).replay
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print("overlap=%s %s: %.4f ms/step" % (ov, mode, e0.elapsed_time(e1) / 50))
        ref = step.dx.clone(); vals = step.vals[:nnz].clone()
        del step
