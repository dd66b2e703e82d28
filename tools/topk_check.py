"""k-best kernel under the current SMLVM_TOPK override: bit-exactness against the oracle on 4096 rows per shape (with
ties and zeros), then timing with and without the fp32 configurations (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import numpy as np
import torch
import oracle
from lvmhelpers import ops
from tools.quick_bench import timeit
for rows, n, k in ((262144, 128, 10), (1 << 20, 32, 10), (262144, 128, 16), (262144, 64, 8), (262144, 100, 5), (262144, 128, 2), (262144, 7, 5)):
    rng = np.random.default_rng(n * 31 + k)
    s = rng.standard_normal((4096, n)).astype(np.float32)
    s[::7] = np.round(s[::7] * 2) / 2
    s[5] = 0.0
    s[6, ::2] = 0.0
    ref, ref_sc = oracle.topk(s, k, return_scores=True)
    pad = np.concatenate([s, rng.standard_normal((16384, n)).astype(np.float32)])   # >= 16384 rows: the throughput kernels
    cfg, bits, sc = ops.topk_bits(torch.from_numpy(pad).cuda(), k, want_scores=True)
    ok = np.array_equal(cfg[:4096].cpu().numpy(), ref) and np.array_equal(sc[:4096].cpu().numpy(), ref_sc)
    b = bits[:4096].cpu().numpy().astype(np.uint32)
    idx = np.arange(n)
    okb = np.array_equal(((b[:, :, idx // 32] >> (idx % 32)) & 1).astype(np.float32), ref)
    x = torch.randn(rows, n, device="cuda")
    ms_b = timeit(lambda: ops.topk_bits(x, k, want_configs=False))
    ms_c = timeit(lambda: ops.topk_bits(x, k))
    byt = rows * (4 * n + k * (4 * n + 4))
    print("rows=%d n=%d k=%d exact=%s bits=%s  bits-only %.3f ms  with fp32 configs %.3f ms (%.0f GB/s)" % (rows, n, k, ok, okb, ms_b, ms_c, byt / ms_c / 1e6), flush=True)
