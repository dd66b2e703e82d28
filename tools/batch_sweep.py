"""BASELINE configs[4]: the batch sweep 4K .. 1M at width 128 on one GPU -- categorical step (resident, eager
launches), k-best bit-vectors (k = 10, fp32 configurations + packed bits), SparseMAP solve (Bernoulli, budget-64)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops
from lvmhelpers.hotpath import SparsemaxMargStep


def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


out = {}
g = torch.Generator().manual_seed(0)
for rows in (4096, 16384, 65536, 262144, 1048576):
    x = torch.randn(rows, 128, generator=g).cuda()
    L = (torch.rand(rows, 128, generator=g) * 3).cuda()
    nnz = SparsemaxMargStep.size_capacity(x)
    step = SparsemaxMargStep(rows, 128, x.device, nnz + 1024, entropy_coeff=1.0)
    ms = timeit(lambda: step.run(x, L), 50)
    row = {"categorical_step_ms": ms, "categorical_samples_per_s": rows / ms * 1e3}
    del step
    if rows <= 262144:
        ms = timeit(lambda: ops.topk_bits(x, 10), 10)
        row.update(topk_ms=ms, topk_rows_per_s=rows / ms * 1e3)
    if rows <= 262144:
        for name, kind, budget in (("bernoulli", 0, 0), ("budget64", 1, 64)):
            ms = timeit(lambda: ops.sparsemap_solve(x, kind, budget=budget, max_iter=300, keep_S=False), 1)
            row["sparsemap_%s_ms" % name] = ms
            row["sparsemap_%s_samples_per_s" % name] = rows / ms * 1e3
    out[rows] = row
    print(rows, json.dumps({k: round(v, 3) if v < 1e4 else round(v) for k, v in row.items()}), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "batch_sweep.json"), "w"), indent=1)
