"""GPU parity against the committed golden fixtures (tests/golden/*.npz, written in the build
container by tests/golden/make_golden.py; nothing here reads /root/reference).

* kernels.npz : kernel-level inputs/outputs (top-k from the reference's own compiled module;
  sparsemax / SparseMAP from the oracle driving the reference's factor headers).
* wrappers.npz: losses, logs and parameter gradients produced by the REFERENCE'S OWN Python
  wrappers (marg.py, structmarg.py, sparsemap.py, unmodified) on small seeded problems.  Our
  drop-in classes must reproduce them through the sm_100a kernels.

Tolerances (BASELINE.json north_star): supports / indices bit-exact, fp32 probabilities 1e-6,
SparseMAP (p, aset) 1e-5; scalar losses and gradients get the fp32 accumulation slack stated
next to each assert.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
from toy import toy_problem, make_modules, loss_fun, CatDecoder, BitDecoder  # noqa: E402


@pytest.fixture(scope="module")
def kern():
    return np.load(os.path.join(GOLD, "kernels.npz"))


@pytest.fixture(scope="module")
def wrap():
    return np.load(os.path.join(GOLD, "wrappers.npz"))


# ----------------------------------------------------------------------------------------
# kernel level
# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "c2", "c5", "ties"])
def test_sparsemax_kernel_vs_golden(kern, name):
    from lvmhelpers import ops
    x = torch.from_numpy(kern[f"sparsemax_{name}_x"]).cuda()
    g = torch.from_numpy(kern[f"sparsemax_{name}_g"]).cuda()
    out = ops.sparsemax_fwd(x, want_entropy=True)
    P = out["p"].cpu().numpy()
    assert np.array_equal(P > 0, kern[f"sparsemax_{name}_p"] > 0)                 # supports bit-exact
    assert np.array_equal(out["supp"].cpu().numpy(), kern[f"sparsemax_{name}_k"])
    assert np.abs(P - kern[f"sparsemax_{name}_p"]).max() <= 1e-6                  # probabilities 1e-6
    dx = ops.sparsemax_bwd(out["p"], out["supp"], dp=g).cpu().numpy()
    assert np.abs(dx - kern[f"sparsemax_{name}_dx"]).max() <= 2e-6
    assert np.allclose(out["entropy"].cpu().numpy(), kern[f"sparsemax_{name}_entropy"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,k", [("c3", 10), ("n128", 10), ("small", 4), ("k32", 32), ("ties", 8)])
def test_topk_kernel_vs_reference_build(kern, name, k):
    from lvmhelpers import ops, pbinary_topk
    x = kern[f"topk_{name}_x"]
    cfg, _, _ = ops.topk_bits(torch.from_numpy(x).cuda(), k)
    assert np.array_equal(cfg.cpu().numpy(), kern[f"topk_{name}_cfg"])           # bit-exact
    out = np.empty_like(kern[f"topk_{name}_cfg"])                                 # Cython-style in-place API
    assert pbinary_topk.batched_topk(x, out, k) is None
    assert np.array_equal(out, kern[f"topk_{name}_cfg"])


@pytest.mark.parametrize("name,kind,budget", [("bern32", 0, 0), ("budget32", 1, 16), ("budget32_3", 1, 3),
                                              ("bern128", 0, 0), ("budget128", 1, 64), ("seq24", 2, 0)])
def test_sparsemap_kernel_vs_golden(kern, name, kind, budget):
    from lvmhelpers import ops
    x = kern[f"smap_{name}_x"]
    rows, n = x.shape
    t = torch.from_numpy(kern[f"smap_{name}_t"]).cuda() if kind == 2 else None
    xs = torch.from_numpy(x).cuda().requires_grad_(True)
    tt = None if t is None else t.requires_grad_(True)
    p, aset, aux = ops.sparsemap_batch(xs, kind, budget=budget, t=tt, max_iter=300, only_positive=False)
    sizes = aux["sizes"].numpy()
    assert np.array_equal(sizes, kern[f"smap_{name}_count"])
    assert np.array_equal(aux["state"]["status"].cpu().numpy(), kern[f"smap_{name}_status"])
    off = np.concatenate([[0], np.cumsum(sizes)])
    w = torch.cat([torch.cos(torch.arange(int(m), dtype=torch.float64)).float() for m in sizes]).cuda()
    (p * w).sum().backward()
    for r in range(rows):
        m = int(sizes[r])
        assert np.array_equal(aset[off[r]:off[r + 1]].cpu().numpy(), kern[f"smap_{name}_aset"][r, :m])
        assert np.abs(p[off[r]:off[r + 1]].detach().cpu().numpy() - kern[f"smap_{name}_p"][r, :m]).max() <= 1e-5
    assert np.abs(xs.grad.cpu().numpy() - kern[f"smap_{name}_dx"]).max() <= 1e-4   # fp32 dp in, fp32 dx out


def test_sparsemap_demo_vector(kern):
    """The reference's print-only demo (lvmhelpers/sparsemap.py:163-178) through the drop-in API."""
    from lvmhelpers.sparsemap import bernoulli_smap, budget_smap
    x = torch.from_numpy(kern["smap_demo_x"]).cuda().requires_grad_(True)
    p, aset = bernoulli_smap(x, init=False)
    assert np.array_equal(aset.cpu().numpy(), kern["smap_demo_bern_aset"].astype(np.float32))
    assert np.abs(p.detach().cpu().numpy() - kern["smap_demo_bern_p"]).max() <= 1e-5
    p, aset = budget_smap(x, budget=3, init=False)
    assert np.array_equal(aset.cpu().numpy(), kern["smap_demo_budget3_aset"].astype(np.float32))
    assert np.abs(p.detach().cpu().numpy() - kern["smap_demo_budget3_p"]).max() <= 1e-5
    assert not aset.requires_grad
    g, = torch.autograd.grad(p[0], x)
    assert torch.isfinite(g).all()


# ----------------------------------------------------------------------------------------
# wrapper level: the reference's own Python wrappers produced these numbers
# ----------------------------------------------------------------------------------------
def _cuda_case(seed, latent, categorical):
    prob = toy_problem(seed, 64, 20, latent, 6)
    enc, dec = make_modules(prob, categorical=categorical)
    return prob["x"].cuda(), prob["labels"].cuda(), enc.cuda(), dec.cuda()


def _check(wrap, name, res, params, rtol=2e-5, gtol=1e-5):
    loss = res["loss"].item()
    assert abs(loss - float(wrap[f"{name}_loss"])) <= rtol * max(1.0, abs(loss))
    for key in ("loss", "encoder_entropy", "acc"):
        assert abs(float(res["log"][key].detach()) - float(wrap[f"{name}_log_{key}"])) <= 2e-5 * max(
            1.0, abs(float(wrap[f"{name}_log_{key}"])))
    sup = res["log"]["support"]
    assert np.array_equal(np.asarray(sup.cpu() if hasattr(sup, "cpu") else sup, dtype=np.float32),
                          wrap[f"{name}_log_support"])
    for pname, prm in params.items():
        ref = wrap[f"{name}_grad_{pname}"]
        tol = gtol * max(1.0, np.abs(ref).max())       # fp32 sums in a different order
        assert np.abs(prm.grad.cpu().numpy() - ref).max() <= tol, pname


def test_marginalizer_vs_reference_wrapper(wrap):
    from lvmhelpers.marg import ExplicitWrapper, Marginalizer
    from lvmhelpers.utils import DeterministicWrapper
    x, labels, enc, dec = _cuda_case(1, 10, True)
    m = Marginalizer(ExplicitWrapper(enc, normalizer="sparsemax"), DeterministicWrapper(CatDecoder(dec)),
                     loss_fun, encoder_entropy_coeff=0.1)
    res = m(x, x, labels)
    res["loss"].backward()
    distr = res["log"]["distr"].detach().cpu().numpy()
    assert np.array_equal(distr > 0, wrap["marg_log_distr"] > 0)
    assert np.abs(distr - wrap["marg_log_distr"]).max() <= 1e-6
    _check(wrap, "marg", res, {"enc_w": enc.weight, "enc_b": enc.bias, "dec": dec.weight})


def test_topk_marg_vs_reference_wrapper(wrap):
    from lvmhelpers.structmarg import TopKSparsemaxWrapper, TopKSparsemaxMarg
    from lvmhelpers.utils import DeterministicWrapper
    x, labels, enc, dec = _cuda_case(2, 32, False)
    m = TopKSparsemaxMarg(TopKSparsemaxWrapper(enc, k=10), DeterministicWrapper(BitDecoder(dec)), loss_fun,
                          encoder_entropy_coeff=1.0)
    res = m(x, x, labels)
    res["loss"].backward()
    distr = res["log"]["distr"].detach().cpu().numpy()
    assert np.array_equal(distr > 0, wrap["topk_log_distr"] > 0)                  # compaction mask bit-exact
    # The sparsemax INPUTS differ from the reference's by the rounding of ONE contraction: scores_k =
    # einsum(bits, Linear(x)) (structmarg.py:47) has |scores_k| ~ 10-46 (fp32 ulp ~ 4e-6); the reference
    # accumulates it in fp32 on the CPU, the re-scoring kernel in fp64 with one final rounding.  The goldens hold
    # the reference's own pipeline BOTH ways (tests/golden/make_golden.py: `topk` = as shipped, `topk64` = its
    # einsum accumulated in fp64), so the tolerance is measured, not asserted in a comment:
    #   * against the fp64-einsum reference the usual bars hold (probabilities 1e-6 of kernel-level parity plus
    #     one ulp of the scores, gradients 1e-5 of their scale);
    #   * against the shipped fp32-einsum reference the result lies inside the band the reference itself spans
    #     between its two variants (x1.5 + the bar above).
    d32, d64 = wrap["topk_log_distr"], wrap["topk64_log_distr"]
    assert np.array_equal(d32 > 0, d64 > 0)
    assert np.abs(distr - d64).max() <= 2e-6
    assert np.abs(distr - d32).max() <= 1.5 * np.abs(d32 - d64).max() + 2e-6
    lo = res["log"]["loss_output"].detach().cpu().numpy()
    assert lo.shape == wrap["topk_log_loss_output"].shape                         # same ragged total, same order
    assert np.allclose(lo, wrap["topk_log_loss_output"], rtol=1e-5, atol=1e-5)
    params = {"enc_w": enc.weight, "enc_b": enc.bias, "dec": dec.weight}
    _check(wrap, "topk64", res, params, gtol=2e-5)
    for pname, prm in params.items():
        r32, r64 = wrap[f"topk_grad_{pname}"], wrap[f"topk64_grad_{pname}"]
        band = 1.5 * np.abs(r32 - r64).max() + 2e-5 * max(1.0, np.abs(r32).max())
        assert np.abs(prm.grad.cpu().numpy() - r32).max() <= band, pname
    # the band itself is what the old hand-set 1e-3 stood for: the reference moves ~2e-4 in enc_w.grad
    assert 1e-5 < np.abs(wrap["topk_grad_enc_w"] - wrap["topk64_grad_enc_w"]).max() < 1e-3


@pytest.mark.parametrize("name,budget,init", [("smap_bern", 0, False), ("smap_budget", 16, False),
                                              ("smap_budget_init", 16, True)])
def test_sparsemap_marg_vs_reference_wrapper(wrap, name, budget, init):
    from lvmhelpers.structmarg import SparseMAPWrapper, SparseMAPMarg
    from lvmhelpers.utils import DeterministicWrapper
    x, labels, enc, dec = _cuda_case(3, 32, False)
    m = SparseMAPMarg(SparseMAPWrapper(enc, budget=budget, init=init, max_iter=300),
                      DeterministicWrapper(BitDecoder(dec)), loss_fun, encoder_entropy_coeff=1.0)
    torch.manual_seed(123)   # the init=True draws come from the CPU generator (sparsemap.py:26)
    res = m(x, x, labels)
    res["loss"].backward()
    assert list(res["log"]["idxs"]) == wrap[f"{name}_log_idxs"].tolist()          # ragged layout bit-exact
    distr = res["log"]["distr"].cpu().numpy()
    assert np.abs(distr - wrap[f"{name}_log_distr"]).max() <= 1e-5
    lo = res["log"]["loss_output"].cpu().numpy()
    assert np.allclose(lo, wrap[f"{name}_log_loss_output"], rtol=1e-5, atol=1e-5)  # same vertices, same order
    _check(wrap, name, res, {"enc_w": enc.weight, "enc_b": enc.bias, "dec": dec.weight}, rtol=5e-5)


def test_marginalizer_compact_support_equals_reference_loop(wrap):
    """decoder on the compacted (sample, latent) pairs only.  Same loss,
    logs and gradients as the reference's per-column loop (deterministic decoder), with
    decoder rows = total support instead of K_active * B."""
    from lvmhelpers.marg import ExplicitWrapper, Marginalizer
    from lvmhelpers.utils import DeterministicWrapper
    x, labels, enc, dec = _cuda_case(1, 10, True)
    m = Marginalizer(ExplicitWrapper(enc, normalizer="sparsemax"), DeterministicWrapper(CatDecoder(dec)),
                     loss_fun, encoder_entropy_coeff=0.1, compact_support=True)
    res = m(x, x, labels)
    res["loss"].backward()
    _check(wrap, "marg", res, {"enc_w": enc.weight, "enc_b": enc.bias, "dec": dec.weight})
    n_support = int((wrap["marg_log_distr"] > 0).sum())
    assert res["log"]["decoder_rows"] == n_support < 64 * 10


def test_slot_policy_follows_the_observed_supports_without_changing_results():
    """ops.sparsemax_stats asks the forward for support slots on large batches until a batch shows that more than
    10 % of its rows overflow them (observed through an asynchronous copy, never a wait); results are the same on
    either path."""
    from lvmhelpers import ops
    from lvmhelpers.marg import ExplicitWrapper, Marginalizer
    from lvmhelpers.utils import DeterministicWrapper
    torch.manual_seed(0)
    rows, K = 4096, 64
    ops._SlotPolicy._state.clear()
    dec = torch.nn.Embedding(K, 8).cuda()

    def loss_fun(enc_in, z, dec_in, dec_out, labels):
        return ((dec_out - labels) ** 2).sum(-1), {}

    labels = torch.randn(rows, 8, device="cuda")
    for scale in (0.05, 3.0):                 # flat rows (supports ~25), then peaked rows (supports ~2)
        x = (torch.randn(rows, K) * scale).cuda()
        res = []
        for it in range(3):
            xin = x.clone().requires_grad_(True)
            m = Marginalizer(ExplicitWrapper(torch.nn.Identity(), normalizer="sparsemax"),
                             DeterministicWrapper(lambda z, _inp: dec(z)), loss_fun, encoder_entropy_coeff=0.3)
            out = m(xin, None, labels)
            out["loss"].backward()
            torch.cuda.synchronize()
            res.append((out["loss"].item(), xin.grad.clone()))
        wants = ops._SlotPolicy.want(x.device, K)
        assert wants == (scale == 3.0)
        for loss, g in res[1:]:
            assert abs(loss - res[0][0]) <= 1e-5 * max(1.0, abs(res[0][0]))
            assert (g - res[0][1]).abs().max().item() <= 1e-6 * max(1.0, res[0][1].abs().max().item())
