import sys, time
sys.path.insert(0, "sparse-marginalization-lvm_b200")
import torch
from lvmhelpers import ops
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
x = torch.randn(rows, 128, device="cuda", requires_grad=True)
for it in range(3):
    x.grad = None
    torch.cuda.synchronize(); t0 = time.time()
    st = ops.sparsemap_solve(x.detach(), 0, max_iter=300)
    torch.cuda.synchronize(); t1 = time.time()
    p, aset, aux = ops.sparsemap_batch(x, 0, max_iter=300)
    torch.cuda.synchronize(); t2 = time.time()
    p.sum().backward()
    torch.cuda.synchronize(); t3 = time.time()
    print("rows %d: solve %.1f ms, batch fwd %.1f ms, bwd %.1f ms; N=%d; peak %.2f GB" % (rows, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, p.numel(), torch.cuda.max_memory_allocated()/1e9))
    del st, p, aset, aux
