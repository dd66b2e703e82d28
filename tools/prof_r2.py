"""ncu target for the round-2 kernels: two passes of (a) the contract step at 1M x 128, N(0,1) scores (tile forward with
support slots, scan, slot-driven fused step), (b) the wide-support step at score scale 0.1 (cooperative forward, scan,
dense fused step), (c) the K = 256 signal-game regime, (d) the top-k wrapper path at 262144 x 128, k = 10, (e) the
stand-alone compaction and entmax15 at 1M x 128 (development aid; `ncu -k regex:smlvm`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops
from lvmhelpers.hotpath import SparsemaxMargStep
from lvmhelpers.structmarg import TopKSparsemaxWrapper

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for rows, K, scale in ((1 << 20, 128, 1.0), (1 << 20, 128, 0.1), (1 << 18, 256, 0.1)):
    x = (torch.randn(rows, K, generator=g) * scale).to(dev)
    L = (torch.rand(rows, K, generator=g) * 3).to(dev)
    nnz = SparsemaxMargStep.size_capacity(x)
    step = SparsemaxMargStep(rows, K, dev, nnz + 1024, entropy_coeff=1.0)
    for _ in range(passes):
        step.run(x, L)
    torch.cuda.synchronize()
    print("step", rows, K, scale, "slots", step.use_slots, "fused", step.fused, "support %.1f" % (nnz / rows))
    if K == 128 and scale == 0.1:
        out = ops.sparsemax_fwd(x)
        off = ops.scan_counts(out["nnz"])
        for _ in range(passes):
            ops.compact_fill(out["p"], off, nnz)
            ops.entmax15(x, want_stats=True)
    del step, x, L
x = torch.randn(1 << 18, 128, generator=g).to(dev).requires_grad_(True)
wrap = TopKSparsemaxWrapper(torch.nn.Identity(), k=10)
cl = (torch.rand((1 << 18) * 10, generator=g) * 500 + 1500).to(dev)
for _ in range(passes):
    x.grad = None
    _, distr, ent = wrap(x)
    (torch.dot(distr.view(-1), cl) * (1.0 / (1 << 18)) - ent).backward()
torch.cuda.synchronize()
print("ok")
