"""Per-phase cycle breakdown of the SparseMAP solver (development aid; needs `make -C .../csrc timing`).
usage: SMLVM_LIB_PATH=tools/_dev/libsmlvm_timing.so python tools/prof_smap_phases.py rows n kind budget [scale]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
os.environ.setdefault("SMLVM_LIB_PATH", os.path.join(ROOT, "tools", "_dev", "libsmlvm_timing.so"))
import torch
from lvmhelpers import ops, _native

rows, n, kind, budget = (int(a) for a in sys.argv[1:5])
scale = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
NAMES = {0: "z-solve", 1: "line search", 2: "p=z (+z-solve)", 3: "u=Mz", 4: "oracle", 5: "common/r", 6: "eval sync + fill round 0 + sync",
         7: "wait: folder + evaluator done", 8: "update + z-solve", 9: "load+seed", 10: "line search + removal", 11: "exit", 12: "  rounds: fill next + d",
         13: "  rounds: wait for the folder"}
torch.manual_seed(0)
x = torch.randn(rows, n, device="cuda") * scale
ops.sparsemap_solve(x, kind, budget=budget, max_iter=300, keep_S=False)
buf = (ctypes.c_ulonglong * 16)()
_native.lib._handle  # noqa
fn = ctypes.CDLL(os.environ["SMLVM_LIB_PATH"]).smlvm_debug_phase_cycles
fn(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
st = ops.sparsemap_solve(x, kind, budget=budget, max_iter=300, keep_S=False)
e1.record()
torch.cuda.synchronize()
fn(buf)
tot = sum(buf)
print("rows %d n %d kind %d budget %d scale %g: %.3f ms, mean iters %.1f, mean |A| %.1f; cycles per sample %.0f" % (
    rows, n, kind, budget, scale, e0.elapsed_time(e1), st["iters"].float().mean().item(), st["counts"].float().mean().item(),
    tot / rows))
for k in sorted(NAMES):
    print("  %-16s %6.1f %%  %9.0f cycles/sample" % (NAMES[k], 100.0 * buf[k] / max(tot, 1), buf[k] / rows))
