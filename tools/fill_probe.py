"""compact_fill (no support slots) under the current SMLVM_COMPACT override: correctness vs torch.nonzero and timing
at 1M x 128 / 256K x 256 over score scales (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops
from tools.quick_bench import timeit
torch.manual_seed(0)
for K in (64, 128, 256):
    rows = (1 << 27) // K
    for scale in (0.03, 0.1, 0.3, 1.0):
        x = torch.randn(rows, K, device="cuda") * scale
        out = ops.sparsemax_fwd(x)
        off = ops.scan_counts(out["nnz"])
        total = int(off[-1])
        v, r, c = ops.compact_fill(out["p"], off, total)
        nz = torch.nonzero(out["p"][:65536] > 0)
        n = nz.shape[0]
        ok = torch.equal(r[:n].long(), nz[:, 0]) and torch.equal(c[:n].long(), nz[:, 1]) and torch.equal(v[:n], out["p"][:65536][out["p"][:65536] > 0])
        ms = timeit(lambda: ops.compact_fill(out["p"], off, total))
        print("K=%d scale=%g nnz=%.1f  %.3f ms  %.0f GB/s ok=%s" % (K, scale, total / rows, ms, (rows * K * 4 + 12 * total) / ms / 1e6, ok), flush=True)
