import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
for mb in (512, 2048):
    x=torch.empty(mb*1024*1024//4, device='cuda'); y=torch.empty_like(x)
    ms=t(lambda: x.zero_()); print(mb,'MB zero_: %.3f ms %.0f GB/s'%(ms, mb*1.048576/ms))
    ms=t(lambda: x.fill_(1.5)); print(mb,'MB fill_: %.3f ms %.0f GB/s'%(ms, mb*1.048576/ms))
    ms=t(lambda: y.copy_(x)); print(mb,'MB copy: %.3f ms %.0f GB/s (r+w)'%(ms, 2*mb*1.048576/ms))
    ms=t(lambda: x.sum()); print(mb,'MB sum: %.3f ms %.0f GB/s'%(ms, mb*1.048576/ms))
