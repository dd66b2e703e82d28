"""Contract step at 1M x 128 (and scale 3 / a mix with rows beyond the slots): per-kernel CUDA-event times of forward, scan, fused
under the current environment (development aid for A/B runs of kernel variants)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers.hotpath import SparsemaxMargStep
from tools.quick_bench import timeit
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
rows, K = 1 << 20, 128
tag = os.environ.get("TAG", "")
for scale in (1.0, 3.0, 0.6):
    x = (torch.randn(rows, K, generator=g) * scale).to(dev)
    L = (torch.rand(rows, K, generator=g) * 3).to(dev)
    nnz = SparsemaxMargStep.size_capacity(x)
    step = SparsemaxMargStep(rows, K, dev, nnz + 1024, entropy_coeff=1.0, use_slots=True)
    ms = timeit(lambda: step.run(x, L), warm=3, reps=20)
    per = {}
    for _ in range(10):
        evs = []
        def rec(name):
            e = torch.cuda.Event(enable_timing=True); e.record(); evs.append((name, e))
        step.run(x, L, rec=rec)
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            per.setdefault(n0, []).append(e0.elapsed_time(e1))
    per = {k: sorted(v)[len(v) // 2] for k, v in per.items()}
    chk = (float(step.dx.double().abs().sum()), float(step.dloss.double().sum()), float(step.total))
    print("%s scale %.1f nnz/row %.2f step %.4f ms  %s  checks dx|.|=%.6f dL=%.6f loss=%.6f" % (tag, scale, nnz / rows, ms, " ".join("%s %.4f" % kv for kv in per.items()), *chk), flush=True)
