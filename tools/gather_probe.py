"""gather_rows / bits_expand throughput at the TopKSparsemaxMarg shape (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops
from tools.quick_bench import timeit
rows, n, k = 1 << 18, 128, 10
g = torch.Generator().manual_seed(0)
x = torch.randn(rows, n, generator=g).cuda()
cnt = torch.randint(1, 7, (rows,), generator=g)
idx = torch.repeat_interleave(torch.arange(rows), cnt).to(torch.int32).cuda()
N = idx.numel()
ms = timeit(lambda: ops.gather_rows(x, idx))
print("gather_rows %d x 512 B: %.3f ms, %.0f GB/s written" % (N, ms, N * 512 / ms / 1e6))
ms = timeit(lambda: x.index_select(0, idx.long()))
print("torch index_select: %.3f ms" % ms)
_, bits, _ = ops.topk_bits(x, k, want_configs=False)
flat = torch.randint(0, rows * k, (N,), generator=g).sort().values.cuda()
ms = timeit(lambda: ops.bits_expand(bits, flat, n))
print("bits_expand %d x 512 B: %.3f ms, %.0f GB/s written" % (N, ms, N * 512 / ms / 1e6))
out = torch.empty(N, n, device="cuda")
ms = timeit(lambda: out.fill_(1.0))
print("fill of the same size: %.3f ms, %.0f GB/s" % (ms, N * 512 / ms / 1e6))
