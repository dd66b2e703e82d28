"""Randomised parity sweep: sparsemax forward (tile and register paths), slots, compaction, slot-driven
loss / backward against the CPU oracle over random shapes, scales and adversarial row patterns."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
import torch
import oracle
from lvmhelpers import ops


def run(seed=0, n_cases=120, verbose=True):
  random.seed(seed)
  torch.manual_seed(random.randrange(1 << 30))
  bad = 0
  for case in range(n_cases):
      K = random.choice([32, 64, 128, 128, 128, 256, 10, 100, 36, 512])
      rows = random.choice([4096, 4096 + random.randrange(1, 200), 6000, 64, 1000])
      scale = 10 ** random.uniform(-2.5, 1.0)
      x = torch.randn(rows, K) * scale
      # adversarial rows
      for r in random.sample(range(rows), min(rows, 40)):
          kind = random.randrange(8)
          if kind == 0: x[r] = torch.round(x[r] * 4) / 4                      # many exact ties
          elif kind == 1: x[r] = 0.0                                         # constant row
          elif kind == 2: x[r] = torch.tensor([1.0] + [0.0] * (K - 1))[torch.randperm(K)]  # gap exactly 1
          elif kind == 3: x[r, : K // 2] = float("-inf")                     # masked
          elif kind == 4: x[r] = x[r] * 1e4                                  # huge magnitudes
          elif kind == 5: x[r] = torch.linspace(-1e-4, 1e-4, K)[torch.randperm(K)]  # nearly constant
          elif kind == 6: x[r] = (torch.arange(K) % 3).float() * 0.5         # periodic ties at thresholds
          elif kind == 7: x[r, random.randrange(K)] += 50.0                  # one-hot
      P_ref, k_ref = oracle.sparsemax_fwd(x)
      xd = x.cuda()
      out = ops.sparsemax_fwd(xd, want_entropy=True, want_argmax=True, want_slots=True)
      P = out["p"].cpu()
      ok = torch.equal(P, P_ref) and torch.equal(out["supp"].cpu().long(), k_ref)
      ok = ok and torch.equal(out["nnz"].cpu().long(), (P_ref > 0).sum(-1))
      ok = ok and torch.equal(out["argmax"].cpu(), x.argmax(-1))
      comp = ops.compact(out["p"], out["nnz"], slots=out["slots"])
      nz = torch.nonzero(P_ref > 0)
      ok = ok and comp["total"] == nz.shape[0] and torch.equal(comp["rows"].cpu().long(), nz[:, 0]) \
          and torch.equal(comp["cols"].cpu().long(), nz[:, 1]) and torch.equal(comp["vals"].cpu(), P_ref[P_ref > 0])
      L = torch.rand(rows, K).cuda()
      dent = torch.randn(rows).cuda()
      a = ops.sparsemax_bwd(out["p"], out["supp"], loss=L, loss_scale_host=0.5, dent=dent)
      b = ops.sparsemax_bwd(out["p"], out["supp"], loss=L, loss_scale_host=0.5, dent=dent, slots=out["slots"])
      fin = torch.isfinite(a) & torch.isfinite(b)
      ok = ok and bool(((a - b)[fin].abs().max() if fin.any() else torch.zeros(())) <= 1e-5 * max(1.0, float(a[fin].abs().max()) if fin.any() else 1.0))
      r0, _ = ops.wloss_dense_fwd(out["p"], L)
      r1, _ = ops.wloss_dense_fwd(out["p"], L, slots=out["slots"])
      ok = ok and bool((r0 - r1).abs().max() <= 1e-5)
      if not ok:
          bad += 1
          diff = (P != P_ref).any(-1).nonzero().flatten()[:5].tolist()
          print("MISMATCH case", case, "rows", rows, "K", K, "scale %.4g" % scale, "rows differing in P:", diff, flush=True)
  if verbose:
    print("fuzz done: %d cases, %d mismatches" % (n_cases, bad))
  return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 120) else 0)

