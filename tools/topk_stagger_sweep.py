"""k-best kernel (topk_blk.cu): exactness of the first 4096 rows against the oracle, bits vs fp32 configurations on every row,
then timing with and without the fp32 configurations (development aid).  It produced profiles/r02_topk_stagger.log when the
kernel had a phase-stagger knob (SMLVM_TOPK_STAGGER, measured and removed: DESIGN.md 4.4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import numpy as np
import torch
import oracle
from lvmhelpers import ops
from tools.quick_bench import timeit
for rows, n, k in ((262144, 128, 10), (262144, 128, 16), (262144, 100, 8), (1 << 20, 128, 10)):
    rng = np.random.default_rng(n * 31 + k)
    s = rng.standard_normal((4096, n)).astype(np.float32)
    s[::7] = np.round(s[::7] * 2) / 2
    ref = oracle.topk(s, k)
    x = torch.randn(rows, n, device="cuda")
    x[:4096] = torch.from_numpy(s).cuda()
    cfg, bits = ops.topk_bits(x, k)[:2]
    ok = np.array_equal(cfg[:4096].cpu().numpy(), ref)
    idx = torch.arange(n, device="cuda")
    exp = ((bits.to(torch.int64)[:, :, idx // 32] >> (idx % 32)) & 1).to(torch.float32)
    okb = bool(torch.equal(exp, cfg))
    del exp
    ms_b = timeit(lambda: ops.topk_bits(x, k, want_configs=False))
    ms_c = timeit(lambda: ops.topk_bits(x, k))
    byt = rows * (4 * n + k * (4 * n + 4))
    print("STAGGER=%s rows=%d n=%d k=%d exact=%s bits==cfg=%s  bits-only %.3f ms  with fp32 configs %.3f ms (%.0f GB/s, %.3f of peak)"
          % (os.environ.get("SMLVM_TOPK_STAGGER", "default"), rows, n, k, ok, okb, ms_b, ms_c, byt / ms_c / 1e6, byt / ms_c / 1e6 / 6550.7), flush=True)
