"""Probe: H2D throughput from default pinned vs write-combined pinned host memory."""
import ctypes, torch, numpy as np
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostAlloc.restype = ctypes.c_int
def wc_tensor(n_floats, flags=4):  # cudaHostAllocWriteCombined = 4
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), n_floats * 4, flags)
    assert rc == 0, rc
    buf = (ctypes.c_float * n_floats).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.float32), p
torch.cuda.init(); torch.zeros(1, device="cuda")
n = 128 * 1024 * 1024  # 512 MB
dev = torch.empty(n, device="cuda")
def t(src, reps=10):
    for _ in range(2): dev.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): dev.copy_(src, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    return n * 4 / (a.elapsed_time(b) / reps) / 1e6
pin = torch.empty(n).pin_memory()
print("default pinned H2D: %.1f GB/s" % t(pin))
wc, _ = wc_tensor(n)
print("is_pinned(wc) =", wc.is_pinned())
print("write-combined H2D: %.1f GB/s" % t(wc))
host = torch.empty(n).pin_memory()
def t2(reps=10):
    for _ in range(2): host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): host.copy_(dev, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    return n * 4 / (a.elapsed_time(b) / reps) / 1e6
print("D2H default pinned: %.1f GB/s" % t2())
# full duplex: the same upload and download at once on two streams (what the host-buffer step does)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
dev2 = torch.empty(n, device="cuda")
def duplex(src, reps=10):
    def go():
        with torch.cuda.stream(s1): dev.copy_(src, non_blocking=True)
        with torch.cuda.stream(s2): host.copy_(dev2, non_blocking=True)
    for _ in range(2): go()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s1.wait_event(a); s2.wait_event(a)
    for _ in range(reps): go()
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1); e2.record(s2)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return n * 4 / ms / 1e6, ms
g, ms = duplex(pin)
print("duplex default pinned: %.1f GB/s each way (%.2f ms for 512 MiB up + 512 MiB down)" % (g, ms))
g, ms = duplex(wc)
print("duplex write-combined source: %.1f GB/s each way (%.2f ms)" % (g, ms))
