"""Fenchel-Young losses the SSVAE experiment imports from ``entmax``
(experiments/semi_supervised-vae/train.py:9,312-325): ``SparsemaxLoss`` and ``Entmax15Loss``.

SURVEY.md §8(f) rank 4 -- outside the hot path proper, provided so that the experiment's last
``entmax`` import has a drop-in.  Definition (entmax ``_GenericLoss``): with p* = normalizer(X),

    L(X, y) = Omega(p*) + <p* - e_y, X>,       dL/dX = p* - e_y,

Omega(p) = (1 - sum_j p_j^alpha) / (alpha (alpha - 1)), alpha = 2 (sparsemax) or 1.5 (entmax15).
fp32 CUDA scores with at most 1024 classes run as kernels: the normalizer (``smlvm_sparsemax_fwd`` /
``smlvm_entmax15_fwd``), the loss rows with p - e_y (``smlvm_fy_loss_fwd``) and the backward (``smlvm_row_scale``).
Other tensors (fp64 checks) take the torch-op form below.
"""
import torch
import torch.nn as nn

from . import ops
from .marg import entmax15


class _FYLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, target, alpha):
        if alpha == 2.0:
            p_star = ops.sparsemax_fwd(X.contiguous(), want_nnz=False)["p"] if X.is_cuda else None
            if p_star is None:
                raise RuntimeError("SparsemaxLoss: scores must live on a CUDA device (no CPU implementation)")
        else:
            p_star = entmax15(X.detach(), dim=-1)
        loss = (1 - (p_star ** alpha).sum(dim=1)) / (alpha * (alpha - 1))
        p_star = p_star.clone()
        p_star.scatter_add_(1, target.unsqueeze(1), torch.full_like(p_star[:, :1], -1.0))
        loss = loss + torch.einsum("ij,ij->i", p_star, X)
        ctx.save_for_backward(p_star)
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        (p_minus_y,) = ctx.saved_tensors
        return grad_output.unsqueeze(1) * p_minus_y, None, None


class _FYLoss(nn.Module):
    alpha = None

    def __init__(self, k=None, ignore_index=-100, reduction="elementwise_mean"):
        super().__init__()
        if reduction not in ("elementwise_mean", "sum", "none"):
            raise ValueError(reduction)
        self.ignore_index, self.reduction = ignore_index, reduction   # k: accepted for API parity, unused

    def forward(self, X, target):
        if X.is_cuda and X.dtype == torch.float32 and X.dim() == 2 and X.shape[-1] <= ops.ENTMAX15_MAX_COLS:
            loss = ops.fy_loss(X, target.clamp(min=0), self.alpha == 1.5)
        else:
            loss = _FYLossFunction.apply(X, target.clamp(min=0), self.alpha)
        size = float(X.shape[0])
        if self.ignore_index >= 0:
            ignored = target == self.ignore_index
            size = float((X.shape[0] - ignored.sum()).item())
            loss = loss.masked_fill(ignored, 0.0)
        if self.reduction == "sum":
            return loss.sum()
        if self.reduction == "elementwise_mean":
            return loss.sum() / size
        return loss


class SparsemaxLoss(_FYLoss):
    alpha = 2.0


class Entmax15Loss(_FYLoss):
    alpha = 1.5
