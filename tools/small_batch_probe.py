"""Per-call GPU time (CUDA events, eager launches back to back) of the pieces of the B = 64 experiment shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops


def t(name, fn, reps=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-40s %8.1f us" % (name, e0.elapsed_time(e1) / reps * 1e3))


B, n, k = 64, 32, 10
x = torch.randn(B, n, device="cuda")
t("topk_bits 64x32 k10", lambda: ops.topk_bits(x, k))
t("topk_bits 64x32 k10 bits only", lambda: ops.topk_bits(x, k, want_configs=False))
cfg, bits, _ = ops.topk_bits(x, k)
t("bits_score fwd", lambda: ops.bits_score(x, bits))
sk = ops.bits_score(x, bits)
t("sparsemax_stats [64,10]", lambda: ops.sparsemax_stats(sk))
out = ops.sparsemax_fwd(sk)
t("sparsemax_fwd [64,10]", lambda: ops.sparsemax_fwd(sk))
t("compact [64,10]", lambda: ops.compact(out["p"], out["nnz"]))
t("sparsemap_solve bern 64x32", lambda: ops.sparsemap_solve(x, 0, max_iter=300))
t("sparsemap_solve budget16 64x32", lambda: ops.sparsemap_solve(x, 1, budget=16, max_iter=300))
x128 = torch.randn(B, 128, device="cuda")
t("topk_bits 64x128 k10", lambda: ops.topk_bits(x128, k))
