#!/usr/bin/env python
"""Benchmark of the sparse-marginalization hot path (BASELINE.json metric:
"latent-marginalized samples/sec fwd+bwd at 1/2/4/8 B200; % HBM roofline").

  python bench.py [--gpus N] [--steps K] [--warmup W]            our sm_100a path
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's CPU path

Workload (BASELINE.json configs[4], the config the metric is quoted on): synthetic scaling
sweep, batched sparsemax marginalisation at latent width 128, 1,048,576 rows per GPU
(weak scaling: every rank owns its own shard, no data-path collective).  One step = one pass
of the categorical hot path over the batch, forward and backward:
  sparsemax fwd (+entropy, argmax, nnz) -> scan -> support compaction -> probability-weighted
  loss -> fused backward (sparsemax Jacobian of the loss + decoder-side gradient).
Inputs (scores X, per-latent losses L: 512 MiB each, >> the 126 MB L2) are resident in HBM
for `value`; `e2e` repeats the step with X and L coming from pinned host memory and the
gradient + loss going back to the host inside the timed region.

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))

import torch  # noqa: E402

METRIC = "latent-marginalized samples/sec fwd+bwd"
UNIT = "samples/s"


_JSON_FD = None


def keep_stdout_for_the_json_line():
    """Only the JSON line may reach stdout: native libraries print there too (NCCL's version banner under
    torchrun).  File descriptor 1 is pointed at stderr for the run; emit() writes to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_inputs(rows, cols, seed):
    """Scores ~ N(0,1), per-latent losses ~ U(0,3); CPU generator so that
    the CPU baseline and the GPU path see identical bits."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, cols, generator=g)
    losses = torch.rand(rows, cols, generator=g) * 3
    return x, losses


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.stop_flag, self.ok = [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                rs = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.time(), mhz, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, t0, t1):
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = [name for bit, name in self.REASONS.items() if bits & bit]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------------------------
# CPU reference path (oracle): what the reference does on host cores for the same step
# --------------------------------------------------------------------------------------
ENTROPY_COEFF = 1.0   # encoder entropy coefficient of the VAE experiments (experiments/bit_vector-vae/train.py:106)


def cpu_step(x, losses, coeff=ENTROPY_COEFF):
    """What the reference does on the host for one step: entmax.sparsemax restated with torch CPU ops
    (marg.py:51), entropy (marg.py:6-28), boolean-mask compaction (structmarg.py:116-126), weighted
    loss (marg.py:125-126), full objective mean loss - coeff * mean entropy (marg.py:119-126) and its
    autograd backward through SparsemaxFunction.backward.  Returns everything the GPU step produces."""
    import oracle
    xr = x.detach().clone().requires_grad_(True)
    lr = losses.detach().clone().requires_grad_(True)
    P = oracle.sparsemax(xr)
    ent = oracle.entropy(P)
    mask = P.detach() > 0
    vals = P.detach()[mask]
    idx = torch.nonzero(mask)
    loss = P.unsqueeze(1).bmm(lr.unsqueeze(-1)).squeeze().mean()
    (loss - coeff * ent.mean()).backward()
    return {"loss": float(loss.detach()), "P": P.detach(), "entropy": ent.detach(), "dx": xr.grad, "dl": lr.grad,
            "vals": vals, "idx": idx}


def time_cpu(x, losses, steps, warmup):
    """-> (rows/s, seconds per step, outputs of the last pass)."""
    out = None
    for _ in range(warmup):
        out = cpu_step(x, losses)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = cpu_step(x, losses)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return x.shape[0] / dt, dt, out


def parity_categorical(step, total_nnz, cpu):
    """GPU step (resident buffers of `step`, already run on the same bits) against the CPU arm's outputs at the
    FULL benchmark size.  Tolerances are BASELINE.json's: supports / compaction indices bit-exact, probabilities
    <= 1e-6 (bit-equal in practice), gradients <= 1e-6 of their scale, loss <= 1e-5."""
    P = step.p.cpu()
    n = int(step.offsets[-1].item())
    idx = cpu["idx"]
    dx, dl = step.dx.cpu(), step.dloss.cpu()
    sx, sl = float(cpu["dx"].abs().max()), float(cpu["dl"].abs().max())
    res = {
        "rows": int(P.shape[0]),
        "p_bit_equal": bool(torch.equal(P, cpu["P"])),
        "p_max_abs_diff": float((P - cpu["P"]).abs().max()),
        "support_sizes_equal": bool(torch.equal(step.nnz.cpu().long(), (cpu["P"] > 0).sum(-1))),
        "compaction_total_equal": n == int(idx.shape[0]) == int(total_nnz),
        "compaction_rows_equal": bool(n == idx.shape[0] and torch.equal(step.ri[:n].cpu().long(), idx[:, 0])),
        "compaction_cols_equal": bool(n == idx.shape[0] and torch.equal(step.ci[:n].cpu().long(), idx[:, 1])),
        "compaction_vals_equal": bool(n == idx.shape[0] and torch.equal(step.vals[:n].cpu(), cpu["vals"])),
        "entropy_max_abs_diff": float((step.ent.cpu() - cpu["entropy"]).abs().max()),
        "dx_max_abs_diff_over_scale": float((dx - cpu["dx"]).abs().max()) / sx,
        "dloss_max_abs_diff_over_scale": float((dl - cpu["dl"]).abs().max()) / sl,
        "loss_abs_diff": abs(float(step.total.item()) - cpu["loss"]),
    }
    res["ok"] = bool(res["p_bit_equal"] and res["support_sizes_equal"] and res["compaction_total_equal"]
                     and res["compaction_rows_equal"] and res["compaction_cols_equal"] and res["compaction_vals_equal"]
                     and res["entropy_max_abs_diff"] <= 1e-5 and res["dx_max_abs_diff_over_scale"] <= 1e-6
                     and res["dloss_max_abs_diff_over_scale"] <= 1e-6 and res["loss_abs_diff"] <= 1e-5)
    return res


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, BEFORE the pinned host buffers
    are allocated (first touch puts them on that NUMA node).  With 8 ranks streaming 1.6 GB per step
    each through host memory, remote-socket buffers halve the end-to-end rate.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_reference(args, rank, world):
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm gets every host thread
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    x, losses = make_inputs(65536, args.cols, 42)
    rate, _, _ = time_cpu(x, losses, 1, 1)  # calibrate
    budget_s = 120.0
    rows = int(min(args.rows, max(4096, rate * budget_s / max(args.steps + args.warmup, 1))))
    rows = 1 << (rows.bit_length() - 1)
    x, losses = make_inputs(rows, args.cols, 42)
    rate, dt, _ = time_cpu(x, losses, args.steps, args.warmup)
    sample = "%d of %d rows per step (%d steps, %d warm-up), torch CPU ops, %d threads" % (
        rows, args.rows, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * (args.rows / rows),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "entmax is not installable offline; the CPU arm is the oracle's torch-CPU "
                "restatement of entmax.sparsemax + the reference's own torch expressions",
    }
    emit(line)


def workload_config(args, world):
    return {"workload": "C5a synthetic sweep: sparsemax marginalisation fwd+bwd, latent width %d, "
                        "%d rows per GPU, scores N(0,1), losses U(0,3)" % (args.cols, args.rows),
            "rows_per_gpu": args.rows, "cols": args.cols, "global_rows": args.rows * world,
            "sharding": "rows, %d shard(s), no data-path collective" % world,
            "l2": "inputs (2 x %d MiB per GPU) exceed the 126 MB L2; no explicit flush"
                  % (args.rows * args.cols * 4 >> 20)}


# --------------------------------------------------------------------------------------
# our path
# --------------------------------------------------------------------------------------
def dist_max(value, world, device):
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        torch.distributed.barrier()


def timed(fn, steps, warmup, world, device):
    """warm-up, barrier + sync, K steps between two CUDA events, sync + barrier; max over ranks (ms)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.time()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t_host1 = time.time()
    barrier(world)
    return dist_max(e0.elapsed_time(e1), world, device), t_host0, t_host1


def small_config_latency(device):
    """The launch-bound experiment shapes (BASELINE configs 1-2, B = 64): microseconds per step of
    the categorical pipeline, eager launches vs one CUDA-graph replay."""
    from lvmhelpers.hotpath import SparsemaxMargStep
    out = {}
    for name, K, scale in (("C1_ssvae_64x10", 10, 3.0), ("C2_signal_game_64x256", 256, 0.1)):
        g = torch.Generator().manual_seed(42)
        x = (torch.randn(64, K, generator=g) * scale).to(device)
        L = (torch.rand(64, K, generator=g) * 3).to(device)
        step = SparsemaxMargStep(64, K, device, 64 * K + 16, entropy_coeff=1.0)
        graph = step.capture(x, L)
        res = {}
        for label, fn in (("eager_us", lambda: step.run(x, L)), ("graph_us", graph.replay)):
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[label] = e0.elapsed_time(e1) / 200 * 1e3
        out[name] = res

    # configs 3-4 (bit-vector VAE, 32 bits, batch 64) through the public wrapper API, forward + backward
    from lvmhelpers import ops
    from lvmhelpers.structmarg import TopKSparsemaxWrapper
    g = torch.Generator().manual_seed(43)
    xb = torch.randn(64, 32, generator=g).to(device).requires_grad_(True)
    wrap = TopKSparsemaxWrapper(torch.nn.Identity(), k=10)
    cl = (torch.rand(640, generator=g) * 3).to(device)

    def topk_step():
        xb.grad = None
        _, distr, ent = wrap(xb)
        (torch.dot(distr.view(-1), cl) * (1.0 / 64) - ent).backward()   # (p . loss) / B - entropy, structmarg.py:140-142

    def smap_step(kind, budget):
        def f():
            xb.grad = None
            p, _, _ = ops.sparsemap_batch(xb, kind, budget=budget, max_iter=300, only_positive=True)
            (ops.ragged_loss(p, torch.ones_like(p), 1.0 / 64) - ops.ragged_entropy(p, 1.0 / 64)).backward()
        return f

    def smap_static_step(kind, budget):
        def f():
            xb.grad = None
            p, _, _ = ops.sparsemap_batch(xb, kind, budget=budget, max_iter=300, only_positive=True, static=True)
            (ops.ragged_loss(p, ones_pad, 1.0 / 64) - ops.ragged_entropy(p, 1.0 / 64)).backward()
        return f

    ones_pad = torch.ones(64 * 33, device=device)

    def graphed(fn):
        """fn (forward + backward, no host synchronisation) captured once, replayed as one launch."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        gr = torch.cuda.CUDAGraph()
        xb.grad = None
        with torch.cuda.graph(gr):
            fn()
        return gr.replay

    for name, fn, static_fn in (("C3_topk_sparsemax_64x32_k10", topk_step, topk_step),
                                ("C4_sparsemap_bernoulli_64x32", smap_step(0, 0), smap_static_step(0, 0)),
                                ("C4_sparsemap_budget16_64x32", smap_step(1, 16), smap_static_step(1, 16))):
        res = {}
        for label, f in (("eager_us", fn), ("graph_us", graphed(static_fn))):
            for _ in range(10):
                f()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100):
                f()
            e1.record()
            torch.cuda.synchronize()
            res[label] = e0.elapsed_time(e1) / 100 * 1e3
            if label == "eager_us":
                res["host_wall_us"] = (time.perf_counter() - t0) / 100 * 1e6
        if name.startswith("C4"):
            # init=True (sparsemap.py:25-27): the 2n uniform draws per sample come from the CPU generator, as in the
            # reference, and seed the active set -- one more H2D copy per step, eager launches only
            from lvmhelpers.sparsemap import draw_init_scores
            kind_b = (0, 0) if "bernoulli" in name else (1, 16)

            def init_step():
                xb.grad = None
                p, _, _ = ops.sparsemap_batch(xb, kind_b[0], budget=kind_b[1], max_iter=300, only_positive=True,
                                              init_scores=draw_init_scores(64, 32))
                (ops.ragged_loss(p, torch.ones_like(p), 1.0 / 64) - ops.ragged_entropy(p, 1.0 / 64)).backward()

            for _ in range(5):
                init_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(50):
                init_step()
            torch.cuda.synchronize()
            res["eager_init_true_host_wall_us"] = (time.perf_counter() - t0) / 50 * 1e6
        res["note"] = ("public wrapper API, forward + backward: eager launches (host-launch bound; the ragged SparseMAP "
                       "outputs read their sizes back) vs one CUDA-graph replay (SparseMAP: static-shape outputs, no host read)")
        out[name] = res
    return out


def extra_paths(args, world, device, rank):
    """The other named shapes of configs[4]: top-k sparsemax and SparseMAP at width 128 (public
    wrapper API, forward + backward, CUDA events, max over ranks)."""
    from lvmhelpers import ops
    from lvmhelpers.structmarg import TopKSparsemaxWrapper
    out = {}
    if rank == 0:
        out["small_configs"] = small_config_latency(device)
    g = torch.Generator().manual_seed(1000 + rank)

    # the categorical step with the decoder on the SUPPORT only (ragged losses): same rows, same scores
    from lvmhelpers.hotpath import SparsemaxMargStep, SparsemaxMargCompactStep
    crow = args.rows
    xc = torch.randn(crow, args.cols, generator=g).to(device)
    n_items = SparsemaxMargStep.size_capacity(xc)
    lr = (torch.rand(n_items, generator=g) * 3).to(device)
    cstep = SparsemaxMargCompactStep(crow, args.cols, device, n_items + 1024, entropy_coeff=1.0)
    marks = []

    def rec(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, ev))

    ms, _, _ = timed(lambda: cstep.run(xc, lr, n_items, rec), 20, 5, world, device)
    per_step = len(SparsemaxMargCompactStep.KERNELS) + 1
    marks = marks[5 * per_step:]
    dur = {k: 0.0 for k in SparsemaxMargCompactStep.KERNELS}
    for i in range(len(marks) - 1):
        if marks[i][0] is not None:
            dur[marks[i][0]] += marks[i][1].elapsed_time(marks[i + 1][1]) / 20
    ab = cstep.algorithmic_bytes(n_items)
    peak = measured_hbm_peak()[0]
    out["sparsemax_marg_compact_support"] = {
        "rows_per_gpu": crow, "cols": args.cols, "support_items": n_items, "ms_per_step": ms / 20,
        "value": crow * world / (ms / 20 * 1e-3), "unit": UNIT,
        "note": "decoder on the support only: ragged losses [N], ragged weighted loss, backward over the support",
        "kernels": {k: {"ms": dur[k], "algorithmic_bytes": ab[k], "frac": ab[k] / (dur[k] * 1e-3) / 1e9 / peak}
                    for k in dur}}
    del cstep, xc, lr

    # (e) the one collective the path has in training: the all-reduce of the wrapped encoder's parameter gradients
    # (lvmhelpers/dist.py::GradAllReducer over NCCL).  The encoder is the bit-vector VAE's inference net shape,
    # Linear(784, 128) -> ReLU -> Linear(128, n) (experiments/bit_vector-vae/train.py:56-60), feeding the sharded
    # categorical step; rows sharded, losses scaled by 1 / B_global.  Three timings: no collective, all-reduce
    # launched after backward, all-reduce launched from backward's own hooks (the last layer's bucket goes out
    # while the first layer's gradients are still being computed).
    from lvmhelpers.dist import GradAllReducer
    erow = min(args.rows, 262144)
    enc = torch.nn.Sequential(torch.nn.Linear(784, 128), torch.nn.ReLU(), torch.nn.Linear(128, args.cols)).to(device)
    with torch.no_grad():
        for i, prm in enumerate(enc.parameters()):     # same weights on every rank
            prm.copy_(torch.randn(prm.shape, generator=torch.Generator().manual_seed(5 + i)) * 0.05)
    einp = torch.randn(erow, 784, generator=g).to(device)
    eloss = (torch.rand(erow, args.cols, generator=g) * 3).to(device)
    estep = SparsemaxMargStep(erow, args.cols, device, 8 * erow, entropy_coeff=1.0, global_rows=erow * world,
                              use_slots=True)
    reducer = GradAllReducer(enc.parameters(), bucket_bytes=1 << 17)     # two buckets: one per layer

    def enc_step(mode):
        enc.zero_grad(set_to_none=True)
        scores = enc(einp)
        estep.run(scores.detach(), eloss)
        scores.backward(estep.dx)
        if mode == "after":
            reducer.start()
        if mode != "none":
            reducer.finish()

    ms_local, _, _ = timed(lambda: enc_step("none"), 10, 3, world, device)
    ms_red, _, _ = timed(lambda: enc_step("after"), 10, 3, world, device)
    reducer.attach()
    ms_hook, _, _ = timed(lambda: enc_step("hooks"), 10, 3, world, device)
    reducer.detach()
    out["sharded_step_with_encoder_allreduce"] = {
        "rows_per_gpu": erow, "encoder": "Linear(784, 128) -> ReLU -> Linear(128, %d), fp32" % args.cols,
        "allreduce_bytes": sum(prm.numel() * 4 for prm in enc.parameters()), "buckets": len(reducer.buckets),
        "backend": "nccl" if world > 1 else "none (1 rank: the reducer is a no-op)",
        "ms_per_step_without_allreduce": ms_local / 10, "ms_per_step_allreduce_after_backward": ms_red / 10,
        "ms_per_step": ms_hook / 10,
        "note": "ms_per_step = all-reduce launched from backward hooks (overlapped)",
        "value": erow * world / (ms_hook / 10 * 1e-3), "unit": UNIT}
    del estep, einp, eloss, enc

    # multi-rank correctness on the GPU: per-rank partial losses with 1 / B_global SUM (NCCL all-reduce) to the
    # single-process CPU loss over the concatenated shards (rank 0 regenerates every shard from its seed)
    if world > 1:
        from lvmhelpers.dist import allreduce_stats
        crow_chk = 8192
        xk, lk = make_inputs(crow_chk, args.cols, 900 + rank)
        kstep = SparsemaxMargStep(crow_chk, args.cols, device, 8 * crow_chk, entropy_coeff=1.0,
                                  global_rows=crow_chk * world, use_slots=True)
        part = kstep.run(xk.to(device), lk.to(device))
        tot = allreduce_stats({"loss": part, "nnz": kstep.nnz.sum()})
        if rank == 0:
            import oracle
            xs, ls = zip(*[make_inputs(crow_chk, args.cols, 900 + r) for r in range(world)])
            Pc, _ = oracle.sparsemax_fwd(torch.cat(xs))
            ref_loss = float((Pc * torch.cat(ls)).sum(-1).mean())
            out.setdefault("_parity", {})["rank_sum"] = {
                "ranks": world, "rows_per_rank": crow_chk, "loss_sum_over_ranks": tot["loss"], "loss_single_process_cpu": ref_loss,
                "support_total_equal": int(tot["nnz"]) == int((Pc > 0).sum()),
                "ok": bool(abs(tot["loss"] - ref_loss) <= 1e-5 and int(tot["nnz"]) == int((Pc > 0).sum()))}
        del kstep

    # the work of the categorical step depends on the score scale (support size).  The contract line is
    # N(0,1) scores (support 3.5 of 128); here the same step at scales 0.3 / 0.1 / 3 and at the signal game's regime
    # (K = 256, scores / tau_s = 10 -> scale 0.1, experiments/signal-game/archs.py:42), per kernel, plus a batch sweep.
    def categorical_variant(vrows, K, scale, steps=10):
        gv = torch.Generator().manual_seed(77)
        xv = (torch.randn(vrows, K, generator=gv) * scale).to(device)
        lv = (torch.rand(vrows, K, generator=gv) * 3).to(device)
        nn_ = SparsemaxMargStep.size_capacity(xv)
        stp = SparsemaxMargStep(vrows, K, device, nn_ + 1024, entropy_coeff=1.0)
        mk = []

        def rec_v(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            mk.append((name, ev))

        msv, _, _ = timed(lambda: stp.run(xv, lv, rec_v), steps, 3, world, device)
        per = len(stp.kernels) + 1
        mk = mk[3 * per:]
        dv = {k: 0.0 for k in stp.kernels}
        for i in range(len(mk) - 1):
            if mk[i][0] is not None:
                dv[mk[i][0]] += mk[i][1].elapsed_time(mk[i + 1][1]) / steps
        abv = stp.algorithmic_bytes(nn_)
        pk = measured_hbm_peak()[0]
        return {"rows_per_gpu": vrows, "cols": K, "score_scale": scale, "support_mean": nn_ / vrows,
                "support_slots": stp.use_slots, "ms_per_step": msv / steps, "value": vrows * world / (msv / steps * 1e-3),
                "unit": UNIT, "step_frac": sum(abv.values()) / (msv / steps * 1e-3) / 1e9 / pk,
                "kernels": {k: {"ms": dv[k], "frac": abv[k] / (dv[k] * 1e-3) / 1e9 / pk if dv[k] > 0 else None} for k in dv}}

    out["score_scale_variants"] = [categorical_variant(args.rows, args.cols, sc) for sc in (3.0, 0.3, 0.1, 0.03)]
    out["signal_game_regime_K256"] = categorical_variant(min(args.rows, 262144), 256, 0.1)
    if rank == 0 and world == 1:
        out["batch_sweep"] = [{k: v for k, v in categorical_variant(b, args.cols, 1.0, steps=20).items()
                               if k in ("rows_per_gpu", "ms_per_step", "value", "step_frac")}
                              for b in (4096, 16384, 65536, 262144) if b <= args.rows]

    rows = args.path_rows_topk
    x = torch.randn(rows, 128, generator=g).to(device).requires_grad_(True)
    wrap = TopKSparsemaxWrapper(torch.nn.Identity(), k=10)
    cand_loss = (torch.rand(rows * 10, generator=g) * 500 + 1500).to(device)

    def topk_step():
        x.grad = None
        bits, distr, ent = wrap(x)
        loss = torch.dot(distr.view(-1), cand_loss) * (1.0 / rows) - ent   # (p . loss) / B - entropy, structmarg.py:140-142
        loss.backward()

    ms, _, _ = timed(topk_step, 5, 2, world, device)
    # k-best search 4n + k(4n+4) B/row (scores in, fp32 configurations + scores out), sparsemax over
    # the k scores fwd+bwd 20k B/row, gradient of the scores 4n B/row
    topk_bytes = rows * (4 * 128 + 10 * (4 * 128 + 4) + 20 * 10 + 4 * 128)
    out["topk_sparsemax_n128_k10"] = {"rows_per_gpu": rows, "ms_per_step": ms / 5,
                                     "value": rows * world / (ms / 5 * 1e-3), "unit": UNIT,
                                     "hbm_frac_of_algorithmic_bytes": topk_bytes / (ms / 5 * 1e-3) / 1e9
                                     / measured_hbm_peak()[0]}

    # the same branch through its consumer, TopKSparsemaxMarg (structmarg.py:82-153): candidates handed over
    # packed, only the support expanded to fp32 for the (dummy) decoder -- against the dense [B,k,n] hand-over
    from lvmhelpers.structmarg import TopKSparsemaxMarg

    class _Opaque(torch.nn.Module):     # hides the wrapper's type: the Marg then takes the dense tensor
        def __init__(self, w):
            super().__init__()
            self.w = w

        def forward(self, *a):
            return self.w(*a)

    def _dec(z, dec_in):
        return z

    def _loss(enc_in, z, dec_in, dec_out, lab):
        return dec_out.sum(-1) * 3.0 + lab, {}

    class _AddBias(torch.nn.Module):    # a parameterised encoder, so that the input itself is data (no gradient)
        def __init__(self):
            super().__init__()
            self.bias = torch.nn.Parameter(torch.zeros(128, device=device))

        def forward(self, inp):
            return inp + self.bias

    dec_in = torch.zeros(rows, 1, device=device)
    lab = torch.rand(rows, generator=g).to(device)
    xin = x.detach()
    wrap_b = TopKSparsemaxWrapper(_AddBias(), k=10)
    res_marg = {}
    for label, enc_mod in (("ms_per_step", wrap_b), ("ms_per_step_dense_candidates", _Opaque(wrap_b))):
        marg = TopKSparsemaxMarg(enc_mod, _dec, _loss, encoder_entropy_coeff=1.0)

        def marg_step():
            wrap_b.agent.bias.grad = None
            marg(xin, dec_in, lab)["loss"].backward()

        ms, _, _ = timed(marg_step, 5, 2, world, device)
        res_marg[label] = ms / 5
    out["topk_sparsemax_marg_n128_k10"] = dict(res_marg, rows_per_gpu=rows, unit=UNIT,
                                               value=rows * world / (res_marg["ms_per_step"] * 1e-3))
    del dec_in, lab

    smap_cases = [("sparsemap_bernoulli_n128", 0, 0, args.path_rows_smap), ("sparsemap_budget64_n128", 1, 64, args.path_rows_smap)]
    if world == 1 and args.rows >= (1 << 20):
        # the sweep's largest point, one GPU: 1,048,576 samples (inverses bump-allocated from one arena, ~30 GB)
        smap_cases += [("sparsemap_bernoulli_n128_1M", 0, 0, 1 << 20), ("sparsemap_budget64_n128_1M", 1, 64, 1 << 20)]
    for name, kind, budget, rows in smap_cases:
        xs = torch.randn(rows, 128, generator=g).to(device).requires_grad_(True)
        stats = {}

        def smap_step():
            xs.grad = None
            p, aset, aux = ops.sparsemap_batch(xs, kind, budget=budget, max_iter=300, only_positive=True)
            w = torch.ones_like(p)
            loss = ops.ragged_loss(p, w, 1.0 / rows) - ops.ragged_entropy(p, 1.0 / rows)
            loss.backward()
            stats["N"] = int(p.numel())
            stats["st"] = aux["state"]

        ms, _, _ = timed(smap_step, 3, 3 if rows <= (1 << 16) else 1, world, device)
        st = stats["st"]
        n_items = stats["N"]
        fwd_bytes = rows * 4 * 128 + n_items * (4 * 128 + 8) + 4 * rows
        bwd_bytes = 4 * n_items + 4 * 128 * n_items + 4 * 128 * rows
        out[name] = {"rows_per_gpu": rows, "ms_per_step": ms / 3, "value": rows * world / (ms / 3 * 1e-3),
                     "unit": UNIT, "mean_support": n_items / rows,
                     "mean_iterations": float(st["iters"].float().mean()),
                     "status_hist": torch.bincount(st["status"], minlength=4).tolist(),
                     "max_iterations": int(st["iters"].max()),
                     "capacity_tiers": [{"capacity": c, "smem_bytes_per_cta": b, "ctas_per_sm": o,
                                         "samples_final_size_in_tier": int(((st["counts"] <= c) & (st["counts"] > lo)).sum())}
                                        for (c, b, o), lo in zip(ops.sparsemap_plan(128, kind),
                                                                 [0] + [t[0] for t in ops.sparsemap_plan(128, kind)][:-1])],
                     "algorithmic_bytes": fwd_bytes + bwd_bytes,
                     "hbm_frac_of_algorithmic_bytes": (fwd_bytes + bwd_bytes) / (ms / 3 * 1e-3) / 1e9
                     / measured_hbm_peak()[0]}
        del xs, st
        stats.clear()
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu:
        # the reference's CPU implementations of the same branches on the box's host cores, bounded
        # samples, single-threaded as in the reference (pbinary_topk.pyx:31-34 and the per-sample
        # solver loop structmarg.py:181-196 are serial): oracle/_ref where the reference compiled,
        # else the oracle port
        import numpy as np
        import oracle
        rng = np.random.default_rng(7)
        xs = rng.standard_normal((20000, 128)).astype(np.float32)
        which = "ref_topk" if oracle.have_ref("ref_topk") else "oracle"
        t0 = time.perf_counter()
        cfg_cpu = oracle.topk(xs, 10, which=which)
        out["topk_sparsemax_n128_k10"]["cpu_topk_only"] = {
            "value": 20000 / (time.perf_counter() - t0), "unit": "rows/s", "cores": 1,
            "kind": "reference" if which == "ref_topk" else "port", "sample": "20000 rows, top-k search alone"}
        par = {}
        # parity of the k-best branch on those 20000 rows: configurations bit-exact against the reference's own
        # binary_topk.h compiled here (oracle/_ref), then the wrapper's distribution over them (sparsemax of the
        # re-scored candidates) against the CPU restatement
        xt = torch.from_numpy(xs).to(device)
        cfg_gpu, bits_gpu, _ = ops.topk_bits(xt, 10)
        sk = ops.bits_score(xt, bits_gpu)
        distr = ops.sparsemax(sk)
        cfg_t = torch.from_numpy(cfg_cpu)
        sk_cpu = torch.einsum("bkj,bj->bk", cfg_t, torch.from_numpy(xs))
        distr_cpu, _ = oracle.sparsemax_fwd(sk_cpu)
        same_support = bool(torch.equal(distr.cpu() > 0, distr_cpu > 0))
        par["topk_n128_k10"] = {
            "rows": 20000, "against": "oracle/_ref (reference header)" if which == "ref_topk" else "oracle port",
            "configs_bit_equal": bool(np.array_equal(cfg_gpu.cpu().numpy(), cfg_cpu)),
            "rescored_max_abs_diff": float((sk.cpu() - sk_cpu).abs().max()),
            "distr_support_equal": same_support,
            "distr_max_abs_diff": float((distr.cpu() - distr_cpu).abs().max())}
        # the fp32 einsum of the reference rounds differently from the fp64-accumulated re-scoring here: the support
        # of the k-sparsemax can flip for an entry within rounding of the threshold, so it is reported, not required
        par["topk_n128_k10"]["ok"] = bool(par["topk_n128_k10"]["configs_bit_equal"]
                                          and par["topk_n128_k10"]["rescored_max_abs_diff"] <= 1e-4
                                          and par["topk_n128_k10"]["distr_max_abs_diff"] <= 1e-4)
        del xt, cfg_gpu, bits_gpu, sk, distr
        for name, kind, budget in (("sparsemap_bernoulli_n128", 0, 0), ("sparsemap_budget64_n128", 1, 64)):
            xs = rng.standard_normal((384, 128)).astype(np.float32)
            t0 = time.perf_counter()
            cnt_cpu, it_cpu = oracle.sparsemap_batch_counts(xs, kind, budget, 300)
            out[name]["cpu_solver_only"] = {
                "value": 384 / (time.perf_counter() - t0), "unit": UNIT, "cores": 1, "kind": "port",
                "sample": "384 samples, forward solve alone, tight C++ loop (no Python per-sample overhead)"}
            # parity on those 384 samples: active-set sizes and iteration counts of every sample, and the full
            # (p, aset, status) of the first 48 -- identical active sets in solver order, p within 1e-5
            st = ops.sparsemap_solve(torch.from_numpy(xs).to(device), kind, budget=budget, max_iter=300, keep_S=False)
            cnt_gpu, it_gpu = st["counts"].cpu().numpy(), st["iters"].cpu().numpy()
            stat_gpu = st["status"].cpu().numpy()
            p_pad, bits = st["p_pad"].cpu().numpy(), st["bits"].cpu().numpy().view(np.uint32)
            worst_p, aset_ok, status_ok = 0.0, True, True
            for r in range(48):
                ref = oracle.sparsemap_fwd(xs[r], kind, budget, max_iter=300)
                m = ref["p"].shape[0]
                if m != cnt_gpu[r]:
                    aset_ok = False
                    continue
                vert = ((bits[r, :m, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(m, -1)[:, :128]
                aset_ok = aset_ok and np.array_equal(vert.astype(np.int32), ref["aset"])
                status_ok = status_ok and int(stat_gpu[r]) == int(ref["status"])
                worst_p = max(worst_p, float(np.abs(p_pad[r, :m] - ref["p"]).max()))
            par[name] = {"samples": 384, "active_set_sizes_equal": bool(np.array_equal(cnt_gpu, cnt_cpu)),
                         "iterations_equal": bool(np.array_equal(it_gpu, it_cpu)),
                         "full_state_samples": 48, "active_sets_identical": bool(aset_ok),
                         "status_equal": bool(status_ok), "p_max_abs_diff": worst_p,
                         "status_hist": np.bincount(stat_gpu, minlength=4).tolist()}
            par[name]["ok"] = bool(par[name]["active_set_sizes_equal"] and par[name]["iterations_equal"]
                                   and aset_ok and status_ok and worst_p <= 1e-5)
        out["_parity"] = par
    return out


def sparsemap_roofline(paths, peak):
    """The SparseMAP branch of the same config (north star: "sparsemax and SparseMAP marginalization") in the terms
    of the contract's `roofline` block: algorithmic bytes of SURVEY 8(d) over the measured step time, next to what
    actually bounds the solver -- serial fp64 work per sample (warp instructions, issue utilisation and resident
    CTAs of the first capacity tier; ncu, profiles/r02_ncu_kernels.md) -- and the measured DRAM bytes of one step."""
    if not paths:
        return None
    tj = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
    out = {}
    for name, key in (("sparsemap_bernoulli_n128", "sparsemap_step_bernoulli_65536x128"),
                      ("sparsemap_budget64_n128", "sparsemap_step_budget64_65536x128")):
        pth = paths.get(name)
        if not pth:
            continue
        meas = tj.get(key, {})
        ab = pth["algorithmic_bytes"]
        ach = ab / (pth["ms_per_step"] * 1e-3) / 1e9
        out[name] = {"bound": "serial fp64 instruction issue per sample (not HBM): the pivoting of libad3 fixes the "
                              "summation order of every sum that feeds z",
                     "rows_per_gpu": pth["rows_per_gpu"], "ms_per_step": pth["ms_per_step"], "value": pth["value"],
                     "unit": UNIT, "achieved": ach, "peak": peak, "achieved_unit": "GB/s", "frac": ach / peak,
                     "algorithmic_bytes": ab, "traffic": meas.get("dram_bytes"),
                     "mean_iterations": pth["mean_iterations"], "mean_support": pth["mean_support"],
                     "status_hist": pth["status_hist"], "capacity_tiers": pth["capacity_tiers"],
                     "ncu_first_tier": meas.get("tier1"), "ncu_kernels_ms": meas.get("kernels_ms")}
        big = paths.get(name + "_1M")
        if big:
            out[name]["at_1048576_rows"] = {"ms_per_step": big["ms_per_step"], "value": big["value"],
                                            "frac": big["hbm_frac_of_algorithmic_bytes"], "status_hist": big["status_hist"]}
    return out


def run_native(args, rank, local_rank, world):
    from lvmhelpers.hotpath import SparsemaxMargStep
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    peak, peak_src = measured_hbm_peak()
    rows, K = args.rows, args.cols

    # pinned host buffers are allocated while bound to the GPU's NUMA node; a single rank gets its CPUs back
    # afterwards (the CPU baseline of the same run uses every host thread), sharded ranks stay bound
    affinity0 = os.sched_getaffinity(0)
    if world == 1:
        torch.ones(1 << 22).add_(1.0).sum()  # start the intra-op thread pool BEFORE binding: its threads keep every CPU
    numa_cpus = bind_to_gpu_numa_node(physical_gpu_index(local_rank))
    x_host, l_host = make_inputs(rows, K, 42 + rank)
    x_pin, l_pin = x_host.pin_memory(), l_host.pin_memory()
    dx_pin = torch.empty(rows, K).pin_memory()
    if world == 1:
        try:
            os.sched_setaffinity(0, affinity0)
        except Exception:
            pass
    x, losses = x_pin.to(device), l_pin.to(device)
    total_nnz = SparsemaxMargStep.size_capacity(x)
    step = SparsemaxMargStep(rows, K, device, capacity=total_nnz + 1024, entropy_coeff=1.0)

    # per-kernel CUDA events, recorded on the launching stream inside the timed region
    marks = []

    def rec(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, ev))

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    ms, th0, th1 = timed(lambda: step.run(x, losses, rec), args.steps, args.warmup, world, device)
    clocks = sampler.summary(th0, th1)

    # kernel durations over the timed steps only (drop the warm-up marks)
    per_step = len(step.kernels) + 1
    marks = marks[args.warmup * per_step:]
    dur = {k: 0.0 for k in step.kernels}
    for i in range(len(marks) - 1):
        name, ev = marks[i]
        if name is not None:
            dur[name] += ev.elapsed_time(marks[i + 1][1])
    dur = {k: v / args.steps for k, v in dur.items()}
    abytes = step.algorithmic_bytes(total_nnz)
    kernels = {k: {"ms": dur[k], "algorithmic_bytes": abytes[k],
                   "gbs": abytes[k] / (dur[k] * 1e-3) / 1e9 if dur[k] > 0 else None,
                   "frac": abytes[k] / (dur[k] * 1e-3) / 1e9 / peak if dur[k] > 0 else None}
               for k in dur}
    top = max(dur, key=dur.get)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        suffix = "" if step.use_slots else "_without_slots"
        traffic = tj.get(top + suffix, tj.get(top))
        for k in kernels:   # measured DRAM bytes per launch (ncu), next to the algorithmic figure
            kernels[k]["traffic"] = tj.get(k + suffix, tj.get(k))
    step_bytes = sum(abytes.values())
    ms_per_step = ms / args.steps
    value = rows * world / (ms_per_step * 1e-3)

    # the same step as a user runs it: no events between the kernels, the compaction branch on the step's side
    # stream under the loss / backward kernels, and as one CUDA-graph replay (informational: `value` above stays
    # the instrumented, serial region the roofline numbers come from)
    ms_plain, _, _ = timed(lambda: step.run(x, losses), args.steps, args.warmup, world, device)
    graph = step.capture(x, losses)
    ms_graph, _, _ = timed(graph.replay, args.steps, args.warmup, world, device)
    uninstrumented = {"ms_per_step": ms_plain / args.steps, "ms_per_step_cuda_graph": ms_graph / args.steps,
                      "value": rows * world / (ms_plain / args.steps * 1e-3),
                      "value_cuda_graph": rows * world / (ms_graph / args.steps * 1e-3), "unit": UNIT,
                      "note": "no per-kernel events between the launches (hotpath.SparsemaxMargStep.run: forward -> scan -> fused step)"}
    del graph

    # end to end through host buffers: H2D of X and L, the step, D2H of gradient + loss, every
    # step, through the host-buffer entry point of the package (micro-batched over 3 streams)
    from lvmhelpers.hotpath import HostPipelinedMargStep
    e2e_steps = max(3, min(args.steps, 20))
    # (the loss matrix stays in its pinned host buffer: the fused step kernel reads it over PCIe on the support only --
    # one read per non-zero of P instead of an upload of all rows x K losses; scores up, dense gradient + loss down)
    zero_copy = step.fused and step.use_slots
    pipe = HostPipelinedMargStep(rows, K, device, capacity_per_row=total_nnz / rows * 1.05 + 0.5,
                                 entropy_coeff=1.0, chunks=args.e2e_chunks, zero_copy_losses=zero_copy)
    ms_e2e, _, _ = timed(lambda: pipe.run(x_pin, l_pin, dx_pin), e2e_steps, 3, world, device)
    loss_value = pipe.loss()
    e2e_dx_err = float((dx_pin[:4096].to(device) - step.dx[:4096]).abs().max())
    # round 1 / early round 2 form, for comparison: both inputs uploaded in full
    pipe_up = HostPipelinedMargStep(rows, K, device, capacity_per_row=total_nnz / rows * 1.05 + 0.5,
                                    entropy_coeff=1.0, chunks=8) if zero_copy else pipe
    ms_e2e_up, _, _ = timed(lambda: pipe_up.run(x_pin, l_pin, dx_pin), e2e_steps, 3, world, device) if zero_copy \
        else (ms_e2e, None, None)
    loss_value_up = pipe_up.loss()
    # the same with the gradient left on the device (the encoder that consumes it runs there): only the loss is read back
    ms_e2e_lo, _, _ = timed(lambda: pipe.run(x_pin, l_pin, None), e2e_steps, 3, world, device)
    # the ceiling of an upload-bound step: a plain pinned-memory H2D copy of the same two inputs, nothing else
    ms_copy, _, _ = timed(lambda: (x.copy_(x_pin, non_blocking=True), losses.copy_(l_pin, non_blocking=True)), 5, 2,
                          world, device)
    gather_wire = 64 * total_nnz if zero_copy else 0    # one PCIe read per non-zero; 64 B each is what the step time implies
    e2e = {"value": rows * world / (ms_e2e / e2e_steps * 1e-3), "unit": UNIT,
           "h2d_gbs_achieved_per_rank": (pipe.h2d_bytes + gather_wire) / (ms_e2e / e2e_steps * 1e-3) / 1e9,
           "h2d_gbs_plain_pinned_copy_per_rank": 2 * rows * K * 4 / (ms_copy / 5 * 1e-3) / 1e9,
           "bound": "PCIe / host memory: every rank uploads rows x K x 4 B of scores per step and reads one loss per "
                    "non-zero of P in place; all ranks of a box share the host's DRAM and PCIe root complexes, so e2e "
                    "scales with host bandwidth, not with GPUs",
           "h2d_bytes_per_step": pipe.h2d_bytes + gather_wire, "d2h_bytes_per_step": pipe.d2h_bytes,
           "h2d_bytes_copied": pipe.h2d_bytes, "h2d_zero_copy_reads": total_nnz if zero_copy else 0,
           "h2d_zero_copy_bytes_useful": 4 * total_nnz if zero_copy else 0,
           "losses_read_in_place": bool(zero_copy),
           "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
           "api": "lvmhelpers.hotpath.HostPipelinedMargStep.run over pinned host buffers (zero_copy_losses=%s), "
                  "%d micro-batches, H2D / kernels / D2H on three streams" % (zero_copy, args.e2e_chunks),
           "check_dx_vs_resident": e2e_dx_err,
           "both_inputs_uploaded": {"value": rows * world / (ms_e2e_up / e2e_steps * 1e-3), "unit": UNIT,
                                    "ms_per_step": ms_e2e_up / e2e_steps, "h2d_bytes_per_step": 2 * rows * K * 4,
                                    "loss": loss_value_up,
                                    "note": "the round-1 form of the same call (zero_copy_losses=False, its best setting of "
                                            "8 micro-batches): the loss matrix uploaded in full before the step"},
           "d2h_contents": "dX [rows, K] + one partial loss per micro-batch; P, dL and the compacted (value, row, col) "
                           "lists stay on the device (the decoder that consumes them runs there)",
           "loss_only_readback": {"value": rows * world / (ms_e2e_lo / e2e_steps * 1e-3), "unit": UNIT,
                                  "ms_per_step": ms_e2e_lo / e2e_steps, "d2h_bytes_per_step": 4 * args.e2e_chunks,
                                  "note": "informational: the same call with dX left on the device; `value` above "
                                          "also returns the dense gradient to the host every step"},
           "host_cpus_bound_per_rank": numa_cpus}

    paths = extra_paths(args, world, device, rank) if not args.no_paths else None
    sampler.stop_flag = True

    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = torch.get_num_threads()
        rate, dt, cpu_out = time_cpu(x_host, l_host, 3, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "host_cpus": os.cpu_count(),
               "sample": "3 timed passes (1 warm-up) over the full %dx%d workload, torch CPU ops" % (rows, K)}
        # parity at the benchmarked size: the resident step's buffers against the CPU arm on the same bits
        step.run(x, losses)
        torch.cuda.synchronize()
        parity = {"categorical_step": parity_categorical(step, total_nnz, cpu_out)}
        del cpu_out
        if paths is not None:
            parity.update(paths.pop("_parity", {}))
        parity["ok"] = all(v.get("ok", False) for v in parity.values() if isinstance(v, dict))
    elif paths is not None:
        extra_par = paths.pop("_parity", None)
        if rank == 0 and extra_par:
            parity = dict(extra_par)
            parity["ok"] = all(v.get("ok", False) for v in parity.values() if isinstance(v, dict))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": args.steps * len(step.kernels),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": kernels[top]["gbs"], "peak": peak,
                         "unit": "GB/s", "frac": kernels[top]["frac"], "traffic": traffic,
                         "peak_source": peak_src,
                         "step": {"algorithmic_bytes": step_bytes,
                                  "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                                  "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak}},
            "roofline_sparsemap": sparsemap_roofline(paths, peak),
            "kernels": kernels,
            "resident_uninstrumented": uninstrumented,
            "cpu_baseline": cpu,
            "paths": paths,
            "check": {"loss": loss_value, "support_mean": total_nnz / rows, "parity": parity},
        }
        emit(line)
        if parity is not None and not parity["ok"]:
            raise SystemExit("bench.py: GPU results differ from the CPU arm: %s" % json.dumps(parity))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--rows", type=int, default=1 << 20, help="rows per GPU")
    ap.add_argument("--cols", type=int, default=128)
    ap.add_argument("--path-rows-topk", type=int, default=1 << 18)
    ap.add_argument("--path-rows-smap", type=int, default=1 << 16)
    ap.add_argument("--e2e-chunks", type=int, default=32)
    ap.add_argument("--no-paths", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    keep_stdout_for_the_json_line()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the sm_100a path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_native(args, rank, local_rank, world)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
