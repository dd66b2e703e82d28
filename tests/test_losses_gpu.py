"""GPU: SparsemaxLoss / Entmax15Loss (the SSVAE experiment's remaining entmax imports) against
their Fenchel-Young definition evaluated with the oracle's sparsemax in fp64, and gradcheck-style
finite differences."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sparsemax_loss_matches_definition(oracle):
    from lvmhelpers.losses import SparsemaxLoss
    g = torch.Generator().manual_seed(0)
    X = torch.randn(64, 10, generator=g) * 2
    y = torch.randint(0, 10, (64,), generator=g)
    P, _ = oracle.sparsemax_fwd(X)
    ref = (1 - (P.double() ** 2).sum(1)) / 2 + (P.double() * X.double()).sum(1) - X.double().gather(1, y[:, None])[:, 0]
    Xc = X.cuda().requires_grad_(True)
    for red, want in (("none", ref), ("sum", ref.sum()), ("elementwise_mean", ref.mean())):
        out = SparsemaxLoss(reduction=red)(Xc, y.cuda())
        # fp32 rows (|loss| ~ 10) and their fp32 sum (~200): a few ulp of the result
        assert (out.detach().cpu().double() - want).abs().max() <= 4e-7 * max(1.0, want.abs().max().item())
    SparsemaxLoss()(Xc, y.cuda()).backward()
    onehot = torch.zeros(64, 10).scatter_(1, y[:, None], 1.0)
    assert (Xc.grad.cpu() - (P - onehot) / 64).abs().max() < 1e-6
    assert (ref >= -1e-6).all()          # a Fenchel-Young loss is non-negative


def test_entmax15_loss_gradient():
    from lvmhelpers.losses import Entmax15Loss
    from lvmhelpers.marg import entmax15
    g = torch.Generator().manual_seed(1)
    X = (torch.randn(8, 6, generator=g, dtype=torch.float64)).cuda().requires_grad_(True)
    y = torch.randint(0, 6, (8,), generator=g).cuda()
    loss = Entmax15Loss(reduction="sum")(X, y)
    loss.backward()
    onehot = torch.zeros(8, 6, dtype=torch.float64, device="cuda").scatter_(1, y[:, None], 1.0)
    assert (X.grad - (entmax15(X.detach()) - onehot)).abs().max() < 1e-12
    # finite differences on one coordinate
    h = 1e-6
    Xp = X.detach().clone(); Xp[0, 0] += h
    Xm = X.detach().clone(); Xm[0, 0] -= h
    fd = (Entmax15Loss(reduction="sum")(Xp, y) - Entmax15Loss(reduction="sum")(Xm, y)) / (2 * h)
    assert abs(fd.item() - X.grad[0, 0].item()) < 1e-6
