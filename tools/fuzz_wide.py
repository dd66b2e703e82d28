"""Randomised parity sweep of the round-2 kernels against the CPU oracle: cooperative sparsemax forward (no slots), the
dense fused step (loss, lists, gradients), the k-sparsemax of the top-k path and entmax15, over random shapes, score
scales and adversarial rows (ties, constants, gaps of exactly 1, -inf masks, huge magnitudes, one-hot rows)."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
import oracle
from lvmhelpers import ops, _native as N
from lvmhelpers.hotpath import SparsemaxMargStep
from lvmhelpers.marg import _Entmax15


def adversarial(x, K):
    rows = x.shape[0]
    for r in random.sample(range(rows), min(rows, 40)):
        kind = random.randrange(8)
        if kind == 0: x[r] = torch.round(x[r] * 4) / 4
        elif kind == 1: x[r] = 0.0
        elif kind == 2: x[r] = torch.tensor([1.0] + [0.0] * (K - 1))[torch.randperm(K)]
        elif kind == 3: x[r, : K // 2] = float("-inf")
        elif kind == 4: x[r] = x[r] * 1e4
        elif kind == 5: x[r] = torch.linspace(-1e-4, 1e-4, K)[torch.randperm(K)]
        elif kind == 6: x[r] = (torch.arange(K) % 3).float() * 0.5
        elif kind == 7: x[r, random.randrange(K)] += 50.0
    return x


def run(seed=0, n_cases=100, verbose=True):
    random.seed(seed)
    torch.manual_seed(random.randrange(1 << 30))
    bad = 0
    for case in range(n_cases):
        K = random.choice([36, 64, 100, 128, 128, 200, 256, 300, 512, 40])
        rows = random.choice([4096, 4096 + random.randrange(1, 300), 6000])
        scale = 10 ** random.uniform(-2.5, 1.0)
        x = adversarial(torch.randn(rows, K) * scale, K)
        P_ref, k_ref = oracle.sparsemax_fwd(x)
        xd = x.cuda()
        out = ops.sparsemax_fwd(xd, want_entropy=True, want_argmax=True)            # no slots: cooperative kernel
        ok = torch.equal(out["p"].cpu(), P_ref) and torch.equal(out["supp"].cpu().long(), k_ref)
        ok = ok and torch.equal(out["nnz"].cpu().long(), (P_ref > 0).sum(-1)) and torch.equal(out["argmax"].cpu(), x.argmax(-1))
        ok = ok and torch.allclose(out["entropy"].cpu(), oracle.entropy(P_ref), rtol=1e-5, atol=1e-6)
        msg = "" if ok else "fwd "
        if K % 4 == 0:
            finite = torch.isfinite(x).all(-1)
            L = torch.rand(rows, K) * 3
            nz = torch.nonzero(P_ref > 0)
            n = nz.shape[0]
            step = SparsemaxMargStep(rows, K, xd.device, n + 16, entropy_coeff=0.5, use_slots=False)
            assert step.fused
            tot = step.run(xd, L.cuda())
            xr = x.clone().requires_grad_(True)
            Pr = oracle.sparsemax(xr)
            ((Pr * L).sum(-1).mean() - 0.5 * oracle.entropy(Pr).mean()).backward()
            ok2 = int(step.offsets[-1]) == n and torch.equal(step.ri[:n].cpu().long(), nz[:, 0]) and torch.equal(step.ci[:n].cpu().long(), nz[:, 1])
            ok2 = ok2 and torch.equal(step.vals[:n].cpu(), P_ref[P_ref > 0])
            ok2 = ok2 and abs(tot.item() - (P_ref * L).sum(-1).mean().item()) <= 1e-5 * max(1.0, abs(tot.item()))
            g_ref = xr.grad
            ok2 = ok2 and (step.dx.cpu()[finite] - g_ref[finite]).abs().max().item() <= 4e-6 * max(1.0, g_ref[finite].abs().max().item())
            ok2 = ok2 and (step.dloss.cpu() - P_ref / rows).abs().max().item() <= 1e-7
            if not ok2:
                msg += "fused "
            ok = ok and ok2
        # k-sparsemax of the top-k path
        k = random.randrange(2, 17)
        sk = adversarial(torch.randn(rows, k) * scale * 3, k) if k >= 4 else torch.randn(rows, k) * scale * 3
        Pk_ref, kk_ref = oracle.sparsemax_fwd(sk)
        skd = sk.cuda()
        distr = torch.empty_like(skd)
        supp = torch.empty(rows, dtype=torch.int32, device="cuda")
        N.check(N.lib.smlvm_ksparsemax_fwd(N.ptr(skd), N.ptr(distr), N.ptr(supp), None, None, None, 1.0, rows, k, None, 0, N.stream()), "ks")
        ok3 = torch.equal(distr.cpu(), Pk_ref) and torch.equal(supp.cpu().long(), kk_ref)
        if not ok3:
            msg += "ksparsemax "
        # entmax15 against the sort-based definition (fp64)
        xe = torch.randn(rows, K) * scale * 3
        pe = ops.entmax15(xe.cuda()).cpu().double()
        ok4 = (pe - _Entmax15.apply(xe.double())).abs().max().item() <= 1e-6
        if not ok4:
            msg += "entmax15 "
        ok = ok and ok3 and ok4
        if not ok:
            bad += 1
        if verbose and (not ok or case % 20 == 0):
            print("case %d K=%d rows=%d scale=%.3g k=%d: %s" % (case, K, rows, scale, k, "ok" if ok else "MISMATCH " + msg), flush=True)
    print("fuzz_wide: %d cases, %d mismatches" % (n_cases, bad))
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 100) else 0)
