"""Ad-hoc kernel timings on one GPU (CUDA events, warm-up, L2-exceeding inputs).
Not the contract benchmark (that is bench.py); used while developing."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
from lvmhelpers import ops

PEAK = 6550.7
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    res = {}
    dev = "cuda"
    which = sys.argv[1:] or ["sparsemax", "compact", "wloss", "topk", "smap", "ragged"]
    if "sparsemax" in which:
        for rows, K, scale in [(1 << 20, 128, 1.0), (1 << 20, 128, 0.3), (1 << 19, 256, 0.1), (1 << 22, 10, 1.0),
                               (1 << 16, 1024, 1.0), (1 << 21, 64, 1.0), (1 << 22, 32, 1.0), (1 << 20, 128, 3.0), (1 << 20, 128, 0.1)]:
            x = torch.randn(rows, K, device=dev) * scale
            g = torch.randn(rows, K, device=dev)
            out = ops.sparsemax_fwd(x)
            ms_f = timeit(lambda: ops.sparsemax_fwd(x))
            ms_b = timeit(lambda: ops.sparsemax_bwd(out["p"], out["supp"], dp=g))
            bf, bb = rows * K * 8, rows * K * 12
            res["sparsemax_%dx%d_s%g" % (rows, K, scale)] = dict(
                fwd_ms=ms_f, bwd_ms=ms_b, fwd_gbs=bf / ms_f / 1e6, bwd_gbs=bb / ms_b / 1e6,
                fwd_frac=bf / ms_f / 1e6 / PEAK, bwd_frac=bb / ms_b / 1e6 / PEAK,
                rows_per_s=rows / ((ms_f + ms_b) * 1e-3), mean_nnz=float(out["nnz"].float().mean()))
            print(list(res.items())[-1], flush=True)
    if "compact" in which:
        rows, K = 1 << 20, 128
        x = torch.randn(rows, K, device=dev)
        out = ops.sparsemax_fwd(x)
        off = ops.scan_counts(out["nnz"])
        total = int(off[-1])
        ms_s = timeit(lambda: ops.scan_counts(out["nnz"]))
        ms_f = timeit(lambda: ops.compact_fill(out["p"], off, total))
        byt = rows * K * 4 + total * 12 + rows * 8
        res["compact_%dx%d" % (rows, K)] = dict(scan_ms=ms_s, fill_ms=ms_f, fill_gbs=byt / ms_f / 1e6,
                                                fill_frac=byt / ms_f / 1e6 / PEAK, total=total)
        print(list(res.items())[-1], flush=True)
    if "wloss" in which:
        rows, K = 1 << 20, 128
        P = torch.rand(rows, K, device=dev)
        L = torch.rand(rows, K, device=dev)
        ms = timeit(lambda: ops.wloss_dense_fwd(P, L, 1.0 / rows))
        res["wloss_dense_%dx%d" % (rows, K)] = dict(ms=ms, gbs=rows * K * 8 / ms / 1e6,
                                                    frac=rows * K * 8 / ms / 1e6 / PEAK)
        print(list(res.items())[-1], flush=True)
    if "topk" in which:
        for rows, n, k in [(1 << 18, 128, 10), (1 << 20, 32, 10)]:
            x = torch.randn(rows, n, device=dev)
            ms = timeit(lambda: ops.topk_bits(x, k, want_bits=True), reps=5)
            byt = rows * (4 * n + k * (4 * n + 4))
            res["topk_%dx%d_k%d" % (rows, n, k)] = dict(ms=ms, rows_per_s=rows / (ms * 1e-3), gbs=byt / ms / 1e6,
                                                        frac=byt / ms / 1e6 / PEAK)
            print(list(res.items())[-1], flush=True)
            ms2 = timeit(lambda: ops.topk_bits(x, k, want_configs=False, want_bits=True), reps=5)
            res["topk_bitsonly_%dx%d_k%d" % (rows, n, k)] = dict(ms=ms2, rows_per_s=rows / (ms2 * 1e-3))
            print(list(res.items())[-1], flush=True)
    if "smap" in which:
        for rows, n, kind, budget in [(1 << 15, 32, 0, 0), (1 << 15, 32, 1, 16), (1 << 12, 128, 0, 0),
                                      (1 << 12, 128, 1, 64)]:
            x = torch.randn(rows, n, device=dev)
            st = ops.sparsemap_solve(x, kind, budget=budget, max_iter=300)
            ms = timeit(lambda: ops.sparsemap_solve(x, kind, budget=budget, max_iter=300), warm=1, reps=3)
            res["smap_%dx%d_kind%d" % (rows, n, kind)] = dict(
                ms=ms, samples_per_s=rows / (ms * 1e-3), mean_A=float(st["counts"].float().mean()),
                mean_iters=float(st["iters"].float().mean()), max_iters=int(st["iters"].max()),
                status_hist=torch.bincount(st["status"], minlength=4).tolist())
            print(list(res.items())[-1], flush=True)
    if "ragged" in which:
        # ragged weighted loss / entropy, bits re-scoring, solver expansion + backward
        N = 1 << 24
        pr = torch.rand(N, device=dev) + 1e-3
        lo = torch.rand(N, device=dev)
        g1 = torch.ones((), device=dev)
        ms = timeit(lambda: ops.wloss_ragged_fwd(pr, lo, scale=1.0, want_total=True, want_entropy=True))
        res["wloss_ragged_fwd_%d" % N] = dict(ms=ms, gbs=N * 8 / ms / 1e6, frac=N * 8 / ms / 1e6 / PEAK)
        print(list(res.items())[-1], flush=True)
        dp, dl = torch.empty_like(pr), torch.empty_like(lo)
        from lvmhelpers._native import lib, ptr, stream
        ms = timeit(lambda: lib.smlvm_wloss_ragged_bwd(ptr(pr), ptr(lo), ptr(g1), ptr(g1), 1.0, ptr(dp), ptr(dl), N,
                                                       stream()))
        res["wloss_ragged_bwd_%d" % N] = dict(ms=ms, gbs=N * 16 / ms / 1e6, frac=N * 16 / ms / 1e6 / PEAK)
        print(list(res.items())[-1], flush=True)
        rows, n, k = 1 << 20, 128, 10
        x = torch.randn(rows, n, device=dev)
        _, bits, _ = ops.topk_bits(x, k, want_configs=False)
        ms = timeit(lambda: ops.bits_score(x, bits))
        byt = rows * (4 * n + k * 4 * (n // 32) + 4 * k)
        res["bits_score_fwd_%dx%d_k%d" % (rows, n, k)] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
        print(list(res.items())[-1], flush=True)
        ds = torch.randn(rows, k, device=dev)
        dxo = torch.empty(rows, n, device=dev)
        ms = timeit(lambda: lib.smlvm_bits_score_bwd(ptr(bits), ptr(ds), ptr(dxo), rows, n, k, stream()))
        res["bits_score_bwd_%dx%d_k%d" % (rows, n, k)] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
        print(list(res.items())[-1], flush=True)
        rows, n = 1 << 14, 128
        xs = torch.randn(rows, n, device=dev)
        st = ops.sparsemap_solve(xs, 0, max_iter=300)
        off = ops.scan_counts(st["kept"])
        tot = int(off[-1])
        ms = timeit(lambda: ops.sparsemap_expand(st, True, off, tot))
        byt = tot * (4 * n + 8) + rows * (n + 1) * (4 + 4 * (n // 32))
        res["smap_expand_%dx%d" % (rows, n)] = dict(ms=ms, items=tot, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
        print(list(res.items())[-1], flush=True)
        dpp = torch.randn(tot, device=dev)
        ms = timeit(lambda: ops.sparsemap_bwd(st, dpp, off, True))
        mA = st["counts"].double()
        byt = float((mA * mA * 8).sum()) + tot * 4 + rows * n * 4
        res["smap_bwd_%dx%d" % (rows, n)] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK,
                                                 note="bytes = S (fp64 |A|^2) + dp + dx")
        print(list(res.items())[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "quick_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
