"""CPU suite (no GPU): pins the oracle.

* against the committed golden vectors (tests/golden/kernels.npz, produced in the build
  container by tests/golden/make_golden.py from the REFERENCE'S OWN compiled code where it
  compiles -- pbinary_topk.pyx/binary_topk.h and the factor headers -- and from the
  restatement where the arithmetic is third-party and absent: entmax, libad3);
* against restatement-independent known answers: simplex / threshold
  identities for sparsemax, brute-force enumeration for top-k, closed-form marginals and
  the KKT conditions for SparseMAP.
"""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "kernels.npz"))


# ----------------------------------------------------------------------------------------
# sparsemax
# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "c2", "c5", "ties"])
def test_sparsemax_golden(oracle, gold, name):
    x = torch.from_numpy(gold[f"sparsemax_{name}_x"])
    g = torch.from_numpy(gold[f"sparsemax_{name}_g"])
    P, k = oracle.sparsemax_fwd(x)
    assert np.array_equal(P.numpy(), gold[f"sparsemax_{name}_p"])
    assert np.array_equal(k.numpy(), gold[f"sparsemax_{name}_k"])
    assert np.array_equal(oracle.sparsemax_bwd(P, k, g).numpy(), gold[f"sparsemax_{name}_dx"])
    assert np.allclose(oracle.entropy(P).numpy(), gold[f"sparsemax_{name}_entropy"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("K,scale", [(10, 3.0), (256, 0.1), (128, 1.0), (5, 1.0), (1, 1.0), (1000, 1.0)])
def test_sparsemax_properties(oracle, K, scale):
    g = torch.Generator().manual_seed(K)
    x = torch.randn(97, K, generator=g) * scale
    P, k = oracle.sparsemax_fwd(x)
    assert (P >= 0).all()
    assert torch.allclose(P.sum(-1), torch.ones(97), atol=2e-6)
    assert torch.equal((P > 0).sum(-1), k)
    # p_i > 0  <=>  x_i > tau, with tau recovered from the support in fp64
    xs = (x - x.max(-1, keepdim=True).values).double()
    tau = ((xs * (P > 0)).sum(-1) - 1) / k
    inside = xs > tau.unsqueeze(-1) + 1e-6
    outside = xs < tau.unsqueeze(-1) - 1e-6
    assert (P[inside] > 0).all() and (P[outside] == 0).all()
    # shift invariance (exact: the row max is subtracted first) and the closed form on the support
    P2, _ = oracle.sparsemax_fwd(x + 3.0)
    assert (P2 - P).abs().max() <= 2e-6
    assert ((xs - tau.unsqueeze(-1)).clamp(min=0) - P.double()).abs().max() <= 1e-6


def test_sparsemax_known_rows(oracle):
    K = 8
    x = torch.stack([torch.zeros(K), torch.tensor([5.0] + [0.0] * (K - 1)),
                     torch.tensor([1.0, 0.0] + [-9.0] * (K - 2)),
                     torch.tensor([0.5, 0.5] + [-9.0] * (K - 2))])
    P, k = oracle.sparsemax_fwd(x)
    assert torch.allclose(P[0], torch.full((K,), 1.0 / K))
    assert torch.equal(P[1], torch.tensor([1.0] + [0.0] * (K - 1)))
    assert torch.equal(P[2], torch.tensor([1.0] + [0.0] * (K - 1))) and int(k[2]) == 1  # gap == 1: strict >
    assert torch.equal(P[3, :2], torch.tensor([0.5, 0.5]))


def test_sparsemax_backward_is_autograd_of_composite(oracle):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(31, 20, generator=g, dtype=torch.float64).requires_grad_(True)
    w = torch.randn(31, 20, generator=g, dtype=torch.float64)
    # composite differentiable definition in fp64: sort, cumsum, threshold, clamp
    xs = x - x.max(-1, keepdim=True).values
    srt, _ = torch.sort(xs, -1, descending=True)
    cs = srt.cumsum(-1) - 1
    rho = torch.arange(1, 21, dtype=torch.float64)
    kk = (rho * srt > cs).sum(-1, keepdim=True)
    tau = cs.gather(-1, kk - 1) / kk
    P = (xs - tau).clamp(min=0)
    (P * w).sum().backward()
    Pf, kf = oracle.sparsemax_fwd(x.detach().float())
    dx = oracle.sparsemax_bwd(Pf, kf, w.float())
    assert (dx.double() - x.grad).abs().max() < 1e-5
    assert (dx * (Pf > 0)).sum(-1).abs().max() < 1e-5  # sum of dX over the support is 0


# ----------------------------------------------------------------------------------------
# top-k bit-vectors
# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,k", [("c3", 10), ("n128", 10), ("small", 4), ("k32", 32), ("ties", 8)])
def test_topk_golden_from_reference_build(oracle, gold, name, k):
    """The golden configs were written by the reference's own compiled pbinary_topk module."""
    cfg = oracle.topk(gold[f"topk_{name}_x"], k)
    assert np.array_equal(cfg, gold[f"topk_{name}_cfg"])


@pytest.mark.parametrize("n,k", [(4, 3), (8, 10), (12, 32), (14, 2)])
def test_topk_bruteforce(oracle, n, k):
    rng = np.random.default_rng(n * 100 + k)
    x = rng.standard_normal((20, n)).astype(np.float32)
    cfg, sc = oracle.topk(x, k, return_scores=True)
    for r in range(20):
        best, _ = oracle.topk_bruteforce(x[r], k)
        got = cfg[r].astype(np.float64) @ x[r].astype(np.float64)
        assert np.allclose(got, best, atol=1e-5)
        assert np.all(np.diff(sc[r]) <= 0)
        assert np.array_equal(cfg[r, 0], (x[r] > 0).astype(np.float32)) or abs(got[0] - best[0]) < 1e-6


def test_topk_live_reference_when_present(oracle):
    if not oracle.have_ref("ref_topk"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(9)
    for n, k in ((32, 10), (128, 10), (7, 5), (64, 32)):
        x = rng.standard_normal((50, n)).astype(np.float32)
        assert np.array_equal(oracle.topk(x, k), oracle.topk(x, k, which="ref_topk"))


def test_topk_rejects_bad_k(oracle):
    x = np.zeros((1, 2), dtype=np.float32)
    with pytest.raises(ValueError):
        oracle.topk(x, 1)
    with pytest.raises(ValueError):
        oracle.topk(x, 5)  # only 4 bit-vectors exist


# ----------------------------------------------------------------------------------------
# SparseMAP
# ----------------------------------------------------------------------------------------
SMAP_CASES = [("bern32", 0, 0), ("budget32", 1, 16), ("budget32_3", 1, 3), ("bern128", 0, 0),
              ("budget128", 1, 64), ("seq24", 2, 0)]


@pytest.mark.parametrize("name,kind,budget", SMAP_CASES)
def test_sparsemap_golden(oracle, gold, name, kind, budget):
    """Golden (p, aset, dx) came from the solver restatement driving the REFERENCE'S factor
    headers (oracle/_ref/libref_factors.so); here the restated factors must reproduce them."""
    x = gold[f"smap_{name}_x"]
    t = gold[f"smap_{name}_t"] if kind == 2 else None
    for r in range(x.shape[0]):
        res = oracle.sparsemap_fwd(x[r], kind, budget, t=None if t is None else t[r], max_iter=300)
        m = int(gold[f"smap_{name}_count"][r])
        assert len(res["p"]) == m and res["status"] == int(gold[f"smap_{name}_status"][r])
        assert np.array_equal(res["aset"], gold[f"smap_{name}_aset"][r, :m])
        assert np.abs(res["p"] - gold[f"smap_{name}_p"][r, :m]).max() <= 1e-12
        dx, _ = oracle.sparsemap_bwd(x[r], np.cos(np.arange(m)), kind, budget,
                                     t=None if t is None else t[r], max_iter=300)
        assert np.abs(dx - gold[f"smap_{name}_dx"][r]).max() <= 1e-9


def test_sparsemap_demo_vector(oracle, gold):
    """The reference's only self-check input (lvmhelpers/sparsemap.py:163-178)."""
    torch.manual_seed(42)
    xd = torch.randn(5).numpy()
    assert np.array_equal(xd, gold["smap_demo_x"])
    for nm, kind, budget in (("bern", 0, 0), ("budget3", 1, 3)):
        res = oracle.sparsemap_fwd(xd, kind, budget, max_iter=100)
        assert np.array_equal(res["aset"], gold[f"smap_demo_{nm}_aset"])
        assert np.abs(res["p"] - gold[f"smap_demo_{nm}_p"]).max() <= 1e-12


@pytest.mark.parametrize("n,kind,budget", [(32, 0, 0), (32, 1, 16), (32, 1, 3), (128, 0, 0), (128, 1, 64), (6, 1, 2)])
@pytest.mark.parametrize("scale", [0.3, 1.0, 3.0])
def test_sparsemap_closed_form_and_kkt(oracle, n, kind, budget, scale):
    rng = np.random.default_rng(n + 13 * kind + budget)
    for r in range(6):
        x = (rng.standard_normal(n) * scale).astype(np.float32)
        res = oracle.sparsemap_fwd(x, kind, budget, max_iter=300)
        p, aset = res["p"], res["aset"].astype(np.float64)
        assert abs(p.sum() - 1) < 1e-4 and p.min() > -1e-9
        u = aset.T @ p
        cf = oracle.bernoulli_marginals(x) if kind == 0 else oracle.budget_marginals(x, budget)
        if res["status"] == 0:
            assert np.abs(u - cf).max() < 1e-4
            # KKT: all active vertices share the reduced score tau; the MAP of the reduced
            # scores does not beat it by more than the solver's own 1e-9 slack (+ drift)
            eta = np.concatenate([x.astype(np.float64), np.zeros(n)])
            g = eta - np.concatenate([u, 1 - u])
            red = aset @ g[:n] + (1 - aset) @ g[n:]
            live = p > 1e-9
            assert np.ptp(red[live]) < 1e-4
            _, best = oracle.factor_map(g, kind, budget)
            assert best <= red[live].max() + 1e-4
        if kind == 1:
            assert aset.sum(1).max() <= budget


def test_sparsemap_init_changes_path_not_marginals(oracle):
    rng = np.random.default_rng(3)
    x = rng.standard_normal(32).astype(np.float32)
    a = oracle.sparsemap_fwd(x, 1, 16, max_iter=300)
    b = oracle.sparsemap_fwd(x, 1, 16, init=rng.random(64), max_iter=300)
    ua = a["aset"].T.astype(np.float64) @ a["p"]
    ub = b["aset"].T.astype(np.float64) @ b["p"]
    assert np.abs(ua - ub).max() < 1e-6


def test_sparsemap_backward_finite_differences(oracle):
    """d(u . w)/dx by central differences on the unique marginals vs M S dp."""
    rng = np.random.default_rng(11)
    n = 10
    x = rng.standard_normal(n) * 0.7
    res = oracle.sparsemap_fwd(x, 0, 0, max_iter=300)
    w = rng.standard_normal(n)
    dp = res["aset"].astype(np.float64) @ w       # d(u.w)/dp_j = w . y_j
    dx, _ = oracle.sparsemap_bwd(x, dp, 0, 0, max_iter=300)
    h = 1e-5
    fd = np.zeros(n)
    for i in range(n):
        e = np.zeros(n)
        e[i] = h
        fd[i] = (oracle.bernoulli_marginals(x + e) @ w - oracle.bernoulli_marginals(x - e) @ w) / (2 * h)
    assert np.abs(dx - fd).max() < 1e-5
    assert abs((res["S"] @ dp).sum()) < 1e-9      # S dp stays in the tangent space of the simplex


def test_factor_maps_live_reference_when_present(oracle):
    if not oracle.have_ref("ref_factors"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(21)
    for kind, budget, n in ((0, 0, 17), (1, 5, 17), (1, 16, 32), (2, 0, 12)):
        for _ in range(50):
            eta = rng.standard_normal(2 * n)
            if rng.random() < 0.3:
                eta = np.round(eta)   # ties
            t = rng.standard_normal(n - 1) if kind == 2 else None
            y0, v0 = oracle.factor_map(eta, kind, budget, t=t)
            y1, v1 = oracle.factor_map(eta, kind, budget, t=t, which="ref_factors")
            assert np.array_equal(y0, y1) and abs(v0 - v1) < 1e-12


def test_general_sequence_restatement_equals_reference_header():
    """oracle RSequence (factors_restated.h) against the reference's own FactorSequence (sequence.h compiled into
    oracle/_ref/libref_factors.so): Viterbi (ties included) and the whole solve, vertex for vertex."""
    import numpy as np
    import oracle
    oracle.build(ref=True)
    if not oracle.have_ref("ref_factors"):
        import pytest
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(3)
    for ns in ([3, 4, 2, 5, 3], [2] * 10, [4] * 6, [1, 3, 1, 2], [7]):
        V, E = sum(ns), oracle.seq_num_edges(ns)
        for trial in range(6):
            x, add = rng.standard_normal(V), rng.standard_normal(E) * 0.5
            if trial == 0:
                x, add = np.round(x), np.zeros(E)
            ya, va = oracle.seq_map(x, add, ns)
            yb, vb = oracle.seq_map(x, add, ns, which="ref_factors")
            assert np.array_equal(ya, yb) and va == vb
            a = oracle.seq_fwd(x, add, ns, max_iter=200)
            b = oracle.seq_fwd(x, add, ns, max_iter=200, which="ref_factors")
            assert np.array_equal(a["aset"], b["aset"]) and np.array_equal(a["p"], b["p"])
            assert a["iters"] == b["iters"] and a["status"] == b["status"]
            # marginals are a distribution over state sequences: every position sums to one
            assert abs(a["p"].sum() - 1) < 1e-9
