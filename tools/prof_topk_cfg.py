import sys
sys.path.insert(0, "sparse-marginalization-lvm_b200")
import torch
from lvmhelpers import ops
x = torch.randn(262144, 128, device="cuda")
for _ in range(2): ops.topk_bits(x, 10)
torch.cuda.synchronize()
