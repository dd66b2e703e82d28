"""The sort-based definition of entmax15 (``lvmhelpers.marg._Entmax15``, the torch-op restatement the kernel is tested
against in tests/test_entmax15_gpu.py): checked here against its defining equations -- p = [(x/2 - tau)_+]^2 with
sum p = 1 -- solved by bisection, and against finite differences.  The product entry points have no CPU path."""
import torch


def _bisect(x):
    x = x.double() / 2
    lo = x.max(-1, keepdim=True).values - 1.0
    hi = x.max(-1, keepdim=True).values
    for _ in range(100):
        mid = (lo + hi) / 2
        f = (torch.clamp(x - mid, min=0) ** 2).sum(-1, keepdim=True)
        lo = torch.where(f >= 1, mid, lo)
        hi = torch.where(f >= 1, hi, mid)
    return torch.clamp(x - (lo + hi) / 2, min=0) ** 2


def test_entmax15_matches_definition():
    from lvmhelpers.marg import _Entmax15
    entmax15 = _Entmax15.apply
    g = torch.Generator().manual_seed(0)
    for K, scale in ((10, 3.0), (256, 0.1), (128, 1.0), (3, 1.0)):
        x = torch.randn(50, K, generator=g, dtype=torch.float64) * scale
        p = entmax15(x)
        assert torch.allclose(p.sum(-1), torch.ones(50, dtype=torch.float64), atol=1e-12)
        assert (p - _bisect(x)).abs().max() < 1e-10
        assert (p >= 0).all()


def test_entmax15_backward_finite_differences():
    from lvmhelpers.marg import _Entmax15
    entmax15 = _Entmax15.apply
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 7, generator=g, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(entmax15, (x,), eps=1e-6, atol=1e-5)


def test_explicit_wrapper_default_normalizer_has_no_cpu_path():
    import pytest
    from lvmhelpers.marg import ExplicitWrapper, entmax15
    w = ExplicitWrapper(torch.nn.Linear(5, 6))          # normalizer='entmax', the reference's default
    with pytest.raises(RuntimeError, match="CUDA"):
        w(torch.randn(9, 5))
    with pytest.raises(RuntimeError, match="CUDA"):
        entmax15(torch.randn(3, 4))
