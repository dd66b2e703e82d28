"""Randomised parity sweep of the batched SparseMAP solver against the CPU oracle's batch loop
(active-set sizes and iteration counts of every sample; full p / active sets on a sample of rows)
and of the k-best bit-vector kernel against the oracle's list merge."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import numpy as np
import torch
import oracle
from lvmhelpers import ops


def run_solver(seed=0, n_cases=24, rows=512, verbose=True):
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(n_cases):
        n = int(rng.choice([8, 32, 33, 64, 100, 128]))
        kind = int(rng.integers(0, 2))
        budget = int(rng.integers(1, n + 1)) if kind == 1 else 0
        scale = float(10 ** rng.uniform(-1.0, 0.7))
        x = (rng.standard_normal((rows, n)) * scale).astype(np.float32)
        if case % 3 == 0:
            x = np.round(x * 2) / 2          # ties between scores, coordinates exactly at the box faces
        st = ops.sparsemap_solve(torch.from_numpy(x).cuda(), kind, budget=budget, max_iter=300, keep_S=False)
        torch.cuda.synchronize()
        c_ref, it_ref = oracle.sparsemap_batch_counts(x, kind, budget, max_iter=300)
        c = st["counts"].cpu().numpy(); it = st["iters"].cpu().numpy()
        ok = np.array_equal(c, c_ref) and np.array_equal(it, it_ref)
        for r in rng.choice(rows, 6, replace=False):
            ref = oracle.sparsemap_fwd(x[r], kind, budget, max_iter=300)
            m = int(c[r])
            p = st["p_pad"][r, :m].cpu().numpy()
            b = st["bits"][r, :m].cpu().numpy().astype(np.uint32)
            idx = np.arange(n)
            aset = ((b[:, idx // 32] >> (idx % 32)) & 1).astype(np.int32)
            ok = ok and aset.shape == ref["aset"].shape and np.array_equal(aset, ref["aset"]) \
                and np.abs(p - ref["p"]).max() <= 1e-5 and int(st["status"][r]) == ref["status"]
        if not ok:
            bad += 1
            print("SOLVER MISMATCH case", case, "n", n, "kind", kind, "budget", budget, "scale %.3g" % scale,
                  "rows with different counts/iters:", np.nonzero((c != c_ref) | (it != it_ref))[0][:8], flush=True)
    if verbose:
        print("solver fuzz: %d cases, %d mismatches" % (n_cases, bad))
    return bad


def run_sequence(seed=0, n_cases=6, rows=48, verbose=True):
    """Sequence-binary factor (transition scores t, 2-state Viterbi oracle): every row against the oracle."""
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(n_cases):
        n = int(rng.choice([2, 5, 20, 33, 64, 100]))
        scale = float(10 ** rng.uniform(-0.7, 0.5))
        x = (rng.standard_normal((rows, n)) * scale).astype(np.float32)
        t = (rng.standard_normal((rows, n - 1)) * scale).astype(np.float32)
        if case % 2 == 0:
            x, t = np.round(x * 2) / 2, np.round(t * 2) / 2
        st = ops.sparsemap_solve(torch.from_numpy(x).cuda(), 2, t=torch.from_numpy(t).cuda(), max_iter=100, keep_S=False)
        torch.cuda.synchronize()
        ok = True
        for r in range(rows):
            ref = oracle.sparsemap_fwd(x[r], 2, t=t[r], max_iter=100)
            m = int(st["counts"][r])
            p = st["p_pad"][r, :m].cpu().numpy()
            b = st["bits"][r, :m].cpu().numpy().astype(np.uint32)
            idx = np.arange(n)
            aset = ((b[:, idx // 32] >> (idx % 32)) & 1).astype(np.int32)
            ok = ok and aset.shape == ref["aset"].shape and np.array_equal(aset, ref["aset"]) \
                and np.abs(p - ref["p"]).max() <= 1e-5 and int(st["status"][r]) == ref["status"] \
                and int(st["iters"][r]) == ref["iters"]
        if not ok:
            bad += 1
            print("SEQUENCE MISMATCH case", case, "n", n, "scale %.3g" % scale, flush=True)
    if verbose:
        print("sequence fuzz: %d cases, %d mismatches" % (n_cases, bad))
    return bad


def run_topk(seed=0, n_cases=40, verbose=True):
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(n_cases):
        n = int(rng.choice([4, 11, 32, 33, 64, 100, 128, 200]))
        k = int(rng.choice([2, 3, 5, 10, 16, 17, 32]))
        k = min(k, 2 ** min(n, 20))
        rows = int(rng.choice([7, 300, 16384 + 5]))
        scale = float(10 ** rng.uniform(-2, 1))
        x = (rng.standard_normal((rows, n)) * scale).astype(np.float32)
        if case % 3 == 0:
            x = np.round(x * 2) / 2          # tied scores: the strict ">" of the merge decides
        if case % 5 == 0:
            x[:, : n // 2] = 0.0             # zero scores: many configurations with equal score
        ref = oracle.topk(x, k)
        cfg = ops.topk_bits(torch.from_numpy(x).cuda(), k, want_bits=False)[0].cpu().numpy()
        if not np.array_equal(cfg, ref):
            bad += 1
            print("TOPK MISMATCH case", case, "rows", rows, "n", n, "k", k, flush=True)
    if verbose:
        print("topk fuzz: %d cases, %d mismatches" % (n_cases, bad))
    return bad


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    b = run_topk(seed) + run_solver(seed) + run_sequence(seed)
    sys.exit(1 if b else 0)
