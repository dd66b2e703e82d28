"""Wide-support categorical step (no slots) at 1M x 128, scales 0.3 / 0.1 / 0.03: per-kernel times under the current environment."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers.hotpath import SparsemaxMargStep
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
rows, K = 1 << 20, 128
for scale in (0.3, 0.1, 0.03):
    x = (torch.randn(rows, K, generator=g) * scale).to(dev)
    L = (torch.rand(rows, K, generator=g) * 3).to(dev)
    nnz = SparsemaxMargStep.size_capacity(x)
    step = SparsemaxMargStep(rows, K, dev, nnz + 1024, entropy_coeff=1.0)
    per = {}
    for it in range(13):
        evs = []
        def rec(name):
            e = torch.cuda.Event(enable_timing=True); e.record(); evs.append((name, e))
        step.run(x, L, rec=rec)
        torch.cuda.synchronize()
        if it >= 3:
            for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
                per.setdefault(n0, []).append(e0.elapsed_time(e1))
    per = {k: sorted(v)[len(v) // 2] for k, v in per.items() if k}
    print(os.environ.get("TAG", ""), "scale %.2f nnz/row %.1f slots=%s" % (scale, nnz / rows, step.use_slots), " ".join("%s %.4f" % kv for kv in per.items()), "total %.4f" % sum(per.values()), flush=True)
