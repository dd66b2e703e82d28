"""ncu target: re-scoring of the k-best configurations from packed bits at the sweep shape (development aid)."""
import sys
sys.path.insert(0, "sparse-marginalization-lvm_b200")
import torch
from lvmhelpers import ops
x = torch.randn(1 << 20, 128, device="cuda")
_, bits, _ = ops.topk_bits(x, 10, want_configs=False)
for _ in range(3): ops.bits_score(x, bits)
torch.cuda.synchronize()
