"""ncu target: a few launches of the sparsemax forward at the bench shape (development aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 128
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
torch.manual_seed(0)
x = torch.randn(rows, K, device="cuda") * scale
for _ in range(3):
    out = ops.sparsemax_fwd(x, want_entropy=True, want_argmax=True)
torch.cuda.synchronize()
print("ok", float(out["nnz"].float().mean()))
