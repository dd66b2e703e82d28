"""GPU: the 1.5-entmax kernel and the Fenchel-Young loss kernels against the package's
sort-based definition restated in torch (lvmhelpers.marg._Entmax15, itself checked against the defining equations in
tests/test_entmax15_cpu.py).  Tolerance: probabilities within 1e-6, as the north star asks of fp32 probabilities."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,K,scale", [(64, 10, 3.0), (64, 256, 0.1), (5000, 128, 1.0), (300, 128, 0.03), (7, 3, 1.0),
                                          (40, 1000, 0.5), (33, 64, 10.0), (9, 1, 1.0), (100, 37, 2.0), (64, 500, 1.0),
                                          (4096, 256, 0.1), (4100, 64, 1.0), (4096, 500, 0.3), (4097, 100, 1.0)])
def test_entmax15_forward_backward(rows, K, scale):
    from lvmhelpers import ops
    from lvmhelpers.marg import _Entmax15, entropy
    g = torch.Generator().manual_seed(rows + K)
    x = torch.randn(rows, K, generator=g) * scale
    x[::5] = torch.round(x[::5] * 4) / 4            # ties, entries exactly at the threshold
    if rows > 8:
        x[3] = 0.5                                  # constant row: uniform
        x[4, 1:] = x[4, 0] - 5.0                    # one-hot
    w = torch.randn(rows, K, generator=g)
    xr = x.double().requires_grad_(True)
    p_ref = _Entmax15.apply(xr)
    (p_ref * w.double()).sum().backward()
    xc = x.cuda().requires_grad_(True)
    p, ent, amax = ops.entmax15(xc, want_stats=True)
    assert (p.detach().cpu().double() - p_ref.detach()).abs().max().item() <= 1e-6
    assert (p.detach().sum(-1) - 1).abs().max().item() <= 1e-5
    assert torch.equal(amax.cpu(), x.argmax(-1))
    # |d (p log p)| = |log p + 1| dp: an entry at the threshold (p ~ 1e-7 on one side, 0 on the other) moves the
    # entropy by ~2e-5 for the 1e-6 the probabilities are allowed
    assert torch.allclose(ent.detach().cpu(), entropy(p_ref.detach().float()), rtol=1e-4, atol=5e-5)
    (p * w.cuda()).sum().backward()
    tol = 2e-5 * max(1.0, xr.grad.abs().max().item())   # sqrt(p) amplifies 1e-6 on p near the threshold
    assert (xc.grad.cpu().double() - xr.grad).abs().max().item() <= tol
    # entropy gradient flows too (ExplicitWrapper returns it)
    xc2 = x.cuda().requires_grad_(True)
    _, ent2, _ = ops.entmax15(xc2, want_stats=True)
    ent2.sum().backward()
    xr2 = x.double().requires_grad_(True)
    pr2 = _Entmax15.apply(xr2)
    eps = torch.finfo(torch.float32).eps
    (-(torch.where(pr2 > 0, pr2.clamp(eps, 1 - eps) * pr2.clamp(eps, 1 - eps).log(), torch.zeros_like(pr2))).sum()).backward()
    assert (xc2.grad.cpu().double() - xr2.grad).abs().max().item() <= 5e-4 * max(1.0, xr2.grad.abs().max().item())


def test_explicit_wrapper_entmax_is_one_launch_and_matches():
    from lvmhelpers.marg import ExplicitWrapper, _Entmax15, entropy
    torch.manual_seed(0)
    agent = torch.nn.Linear(12, 10).cuda()
    w = ExplicitWrapper(agent)                       # the reference's default normalizer: 'entmax'
    x = torch.randn(64, 12, device="cuda")
    sample, distr, ent = w(x)
    scores = agent(x).detach()
    ref = _Entmax15.apply(scores.double().cpu())
    assert (distr.detach().cpu().double() - ref).abs().max().item() <= 1e-6
    assert torch.equal(sample.cpu(), scores.argmax(-1).cpu())
    assert torch.allclose(ent.detach().cpu(), entropy(ref.float()), rtol=1e-4, atol=2e-6)
    (distr.sum() + ent.sum()).backward()
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in agent.parameters())


@pytest.mark.parametrize("alpha", [1.5, 2.0])
@pytest.mark.parametrize("rows,K", [(64, 10), (500, 100), (9, 3)])
def test_fy_loss_kernels_match_definition(oracle, alpha, rows, K):
    from lvmhelpers.losses import SparsemaxLoss, Entmax15Loss
    from lvmhelpers.marg import _Entmax15
    g = torch.Generator().manual_seed(K)
    X = torch.randn(rows, K, generator=g) * 2
    y = torch.randint(0, K, (rows,), generator=g)
    P = (oracle.sparsemax_fwd(X)[0] if alpha == 2.0 else _Entmax15.apply(X.double())).double()
    ref = (1 - (P ** alpha).sum(1)) / (alpha * (alpha - 1)) + (P * X.double()).sum(1) - X.double().gather(1, y[:, None])[:, 0]
    cls = SparsemaxLoss if alpha == 2.0 else Entmax15Loss
    Xc = X.cuda().requires_grad_(True)
    out = cls(reduction="none")(Xc, y.cuda())
    assert (out.detach().cpu().double() - ref).abs().max().item() <= 2e-5
    wgt = torch.rand(rows, generator=g)
    (out * wgt.cuda()).sum().backward()
    onehot = torch.zeros(rows, K, dtype=torch.float64).scatter_(1, y[:, None], 1.0)
    assert (Xc.grad.cpu().double() - wgt.double()[:, None] * (P - onehot)).abs().max().item() <= 2e-6
