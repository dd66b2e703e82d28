import sys
sys.path.insert(0, "sparse-marginalization-lvm_b200")
import torch
from lvmhelpers import ops
x = torch.randn(262144, 128, device="cuda")
for k in (16, 10, 8, 5, 4, 2):
    for _ in range(3): ops.topk_bits(x, k, want_configs=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.topk_bits(x, k, want_configs=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    per_warp = (2 * k * 4 * 32 + 32 * 33 + (2 * k + 1) * 32) * 4
    warps = min(16, (200 * 1024) // per_warp)
    print("k=%2d: %.3f ms, %.2f ns per (row, bit, slot), smem/warp %.1f KB, warps/SM %d" % (k, ms, ms * 1e6 / (262144 * 128 * k), per_warp / 1024, warps))
