"""Per-kernel times of the top-k wrapper path at 262144 x 128, k = 10 (CUDA events; development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops, _native as N
from tools.quick_bench import timeit
rows, n, k = 1 << 18, 128, 10
x = torch.randn(rows, n, device="cuda")
_, bits, _ = ops.topk_bits(x, k, want_configs=False)
sk = ops.bits_score(x, bits)
out = ops.sparsemax_fwd(sk)
distr, supp = out["p"], out["supp"]
dd = torch.randn(rows, k, device="cuda")
de = torch.ones((), device="cuda")
dx = torch.empty(rows, n, device="cuda")
print("bits_score_fwd %.4f ms" % timeit(lambda: ops.bits_score(x, bits)))
print("topk_sparsemax_bwd %.4f ms" % timeit(lambda: N.check(N.lib.smlvm_topk_sparsemax_bwd(N.ptr(bits), N.ptr(distr), N.ptr(supp), N.ptr(dd), N.ptr(de), 1.0 / rows, N.ptr(dx), rows, n, k, N.stream()), "b")))
print("bits_score_bwd %.4f ms" % timeit(lambda: N.check(N.lib.smlvm_bits_score_bwd(N.ptr(bits), N.ptr(dd), N.ptr(dx), rows, n, k, N.stream()), "b")))
print("mean nnz %.2f" % out["nnz"].float().mean().item())
