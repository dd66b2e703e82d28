"""ncu / timing target: the SparseMAP solver at the sweep shape (development aid)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
budget = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
scale = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
torch.manual_seed(0)
x = torch.randn(rows, n, device="cuda") * scale
st = ops.sparsemap_solve(x, kind, budget=budget, max_iter=300, keep_S=False)
torch.cuda.synchronize()
times = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = ops.sparsemap_solve(x, kind, budget=budget, max_iter=300, keep_S=False)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
times.sort()
ms = times[0]
print(os.environ.get("TAG", ""), "scale %g" % scale, "rows %d n %d kind %d budget %d: min %.3f ms (median %.3f, max %.3f of %d), %.0f samples/s, mean iters %.1f, mean |A| %.1f, status %s" % (
    rows, n, kind, budget, ms, times[len(times) // 2], times[-1], reps, rows / ms * 1e3, st["iters"].float().mean().item(),
    st["counts"].float().mean().item(), torch.bincount(st["status"], minlength=4).tolist()))
