"""Forward-kernel sweep over (K, score scale) for the current SMLVM_SPARSEMAX_FWD path (development aid)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops
from tools.quick_bench import timeit
torch.manual_seed(0)
for K in (32, 64, 128, 256):
    rows = (1 << 27) // K
    for scale in (0.03, 0.1, 0.3, 1.0, 3.0):
        x = torch.randn(rows, K, device="cuda") * scale
        out = ops.sparsemax_fwd(x, want_entropy=True, want_argmax=True)
        ms = timeit(lambda: ops.sparsemax_fwd(x, want_entropy=True, want_argmax=True))
        print("K=%d scale=%g nnz=%.1f  %.3f ms  %.0f GB/s" % (K, scale, out["nnz"].float().mean().item(), ms, rows * K * 8 / ms / 1e6), flush=True)
