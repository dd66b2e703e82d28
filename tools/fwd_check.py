"""Forward kernel under the current SMLVM_SPARSEMAX_FWD override: bit-exactness against the oracle on 16384 rows per
(K, scale), then timing on 2^27 elements (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
import oracle
from lvmhelpers import ops
from tools.quick_bench import timeit
torch.manual_seed(0)
Ks = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 128, 256, 512]
for K in Ks:
    for scale in (0.03, 0.1, 0.3, 1.0, 3.0):
        xs = torch.randn(16384, K) * scale
        xs[::7] = torch.round(xs[::7] * 8) / 8
        xs[5] = 0.25
        xs[6, 1:] = xs[6, 0] - 1.0
        xs[9, ::3] = float("-inf")
        P_ref, k_ref = oracle.sparsemax_fwd(xs)
        out = ops.sparsemax_fwd(xs.cuda(), want_entropy=True, want_argmax=True)
        ok = torch.equal(out["p"].cpu(), P_ref) and torch.equal(out["supp"].cpu().long(), k_ref)
        ok_nnz = torch.equal(out["nnz"].cpu().long(), (P_ref > 0).sum(-1))
        ok_amax = torch.equal(out["argmax"].cpu(), xs.argmax(-1))
        ent_ref = oracle.entropy(P_ref)
        ent_err = (out["entropy"].cpu() - ent_ref).abs().max().item()
        rows = (1 << 27) // K
        x = torch.randn(rows, K, device="cuda") * scale
        o2 = ops.sparsemax_fwd(x, want_entropy=True, want_argmax=True)
        ms = timeit(lambda: ops.sparsemax_fwd(x, want_entropy=True, want_argmax=True))
        print("K=%d scale=%g nnz=%.1f  %.3f ms  %.0f GB/s  exact=%s nnz=%s amax=%s ent_err=%.2e" % (
            K, scale, o2["nnz"].float().mean().item(), ms, rows * K * 8 / ms / 1e6, ok, ok_nnz, ok_amax, ent_err), flush=True)
