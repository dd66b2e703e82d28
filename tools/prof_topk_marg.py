import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers.structmarg import TopKSparsemaxWrapper, TopKSparsemaxMarg
rows = 1 << 18
dev = "cuda"
g = torch.Generator().manual_seed(0)
class AddBias(torch.nn.Module):
    def __init__(self):
        super().__init__(); self.bias = torch.nn.Parameter(torch.zeros(128, device=dev))
    def forward(self, inp): return inp + self.bias
x = torch.randn(rows, 128, generator=g).to(dev)
w = TopKSparsemaxWrapper(AddBias(), k=10)
marg = TopKSparsemaxMarg(w, lambda z, d: z, lambda e, z, d, o, l: (o.sum(-1) * 3.0 + l, {}), encoder_entropy_coeff=1.0)
dec_in = torch.zeros(rows, 1, device=dev); lab = torch.rand(rows, generator=g).to(dev)
for _ in range(3):
    w.agent.bias.grad = None
    marg(x, dec_in, lab)["loss"].backward()
torch.cuda.synchronize(); print("ok")
