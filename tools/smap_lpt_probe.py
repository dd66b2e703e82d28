"""How much of a SparseMAP solve is the ramp-down tail of its first tier?  The same batch solved in sample order, longest-first
(by the number of fractional coordinates, |x| < 1) and shortest-first (development probe for a permutation argument)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops
from tools.quick_bench import timeit
torch.manual_seed(0)
for rows in (16384, 65536, 262144):
    for kind, budget in ((0, 0), (1, 64)):
        x = torch.randn(rows, 128, device="cuda")
        key = (x.abs() < 1).sum(1)
        st = ops.sparsemap_solve(x, kind, budget=budget, max_iter=300, keep_S=False)
        it = st["iters"].float()
        res = []
        for name, perm in (("sample order", None), ("longest first (#|x|<1)", torch.argsort(key, descending=True)),
                           ("shortest first", torch.argsort(key)), ("longest first (true iterations)", torch.argsort(it, descending=True))):
            xx = x if perm is None else x[perm].contiguous()
            ms = timeit(lambda: ops.sparsemap_solve(xx, kind, budget=budget, max_iter=300, keep_S=False), warm=1, reps=3)
            res.append("%s %.2f ms" % (name, ms))
        print("rows %d kind %d budget %d: corr(key, iters) %.2f | %s" % (rows, kind, budget, torch.corrcoef(torch.stack([key.float(), it]))[0, 1].item(), " | ".join(res)), flush=True)
