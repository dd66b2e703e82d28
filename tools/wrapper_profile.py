"""cProfile of one experiment-shaped wrapper step (B = 64) to see where the host time goes (development aid)."""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import importlib.util
spec = importlib.util.spec_from_file_location("wl", os.path.join(ROOT, "tools", "wrapper_latency.py"))
src = open(os.path.join(ROOT, "tools", "wrapper_latency.py")).read().split('bench("C1 marg')[0]
ns = {"__file__": os.path.join(ROOT, "tools", "wrapper_latency.py")}
exec(compile(src, "wrapper_latency.py", "exec"), ns)
import torch
which = sys.argv[1] if len(sys.argv) > 1 else "topk"
make = {"topk": ns["topk"], "smap": ns["smap"](0, False), "cat": ns["cat"](10, True)}[which]
m, x, labels, params = make()
def step():
    for p in params: p.grad = None
    res = m(x, x, labels)
    res["loss"].backward()
for _ in range(20): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
