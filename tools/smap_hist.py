import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/sparse-marginalization-lvm_b200")
import torch
from lvmhelpers import ops
torch.manual_seed(0)
for n, kind, budget, scale in [(128,0,0,1.0),(128,1,64,1.0),(128,0,0,0.3),(128,0,0,3.0),(128,1,64,0.3),(128,1,64,3.0),(100,0,0,1.0),(64,0,0,1.0),(64,1,32,1.0)]:
    x = torch.randn(8192, n, device="cuda") * scale
    st = ops.sparsemap_solve(x, kind, budget=budget, max_iter=300)
    c = st["counts"].float()
    q = torch.quantile(c, torch.tensor([0.5, 0.9, 0.99, 0.999], device="cuda")).tolist()
    print(n, kind, budget, scale, "mean %.1f" % c.mean().item(), "q50/90/99/99.9", [round(v) for v in q], "max", int(c.max()), "iters mean %.0f max %d" % (st["iters"].float().mean().item(), int(st["iters"].max())))
