import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import numpy as np, torch, oracle
from lvmhelpers import ops
n, kind, budget, scale = 32, 1, 3, 0.3
rows = 64
rng = np.random.default_rng(n * 7 + kind + budget)
x = (rng.standard_normal((rows, n)) * scale).astype(np.float32)
st = ops.sparsemap_solve(torch.from_numpy(x).cuda(), kind, budget=budget, max_iter=300)
idx = np.arange(n)
bad = 0
for r in range(rows):
    ref = oracle.sparsemap_fwd(x[r], kind, budget, max_iter=300)
    m = int(st["counts"][r]); p = st["p_pad"][r,:m].cpu().numpy()
    b = st["bits"][r,:m].cpu().numpy().astype(np.uint32)
    aset = ((b[:, idx // 32] >> (idx % 32)) & 1).astype(np.int32)
    if aset.shape != ref["aset"].shape or not np.array_equal(aset, ref["aset"]):
        bad += 1
        if bad > 3: continue
        print("row", r, "gpu |A|", m, "it", int(st["iters"][r]), "st", int(st["status"][r]), "| ref |A|", len(ref["p"]), "it", ref["iters"], "st", ref["status"])
        print(" gpu verts:", [tuple(np.nonzero(a)[0]) for a in aset])
        print(" ref verts:", [tuple(np.nonzero(a)[0]) for a in ref["aset"]])
        print(" gpu p:", np.round(p, 6)); print(" ref p:", np.round(ref["p"], 6))
        g = np.sort(x[r])[::-1][:8]; print(" top gains:", g)
print("bad rows:", bad)
