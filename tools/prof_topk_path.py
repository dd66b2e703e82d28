"""Launch list target: the top-k sparsemax wrapper path at the sweep shape (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops
from lvmhelpers.structmarg import TopKSparsemaxWrapper
rows = 1 << 18
g = torch.Generator().manual_seed(0)
x = torch.randn(rows, 128, generator=g).cuda().requires_grad_(True)
wrap = TopKSparsemaxWrapper(torch.nn.Identity(), k=10)
cand_loss = (torch.rand(rows * 10, generator=g) * 500 + 1500).cuda()
for it in range(3):
    x.grad = None
    bits, distr, ent = wrap(x)
    loss = torch.dot(distr.view(-1), cand_loss) * (1.0 / rows) - ent
    loss.backward()
torch.cuda.synchronize()
print("ok")
