"""Contrastive losses with the reference's call signatures (``src/openmatch/loss.py:7-38``); the
arithmetic (logits, log-softmax, loss and both gradients) runs in libopenmatch_b200.so (csrc/loss.cu) and is
exposed to autograd through one ``torch.autograd.Function``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from torch import distributed as dist

from . import _lib


class _FusedContrastive(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, y: Tensor, target: Optional[Tensor], reduction: str, want_scores: bool):
        if not (x.is_cuda and y.is_cuda):
            raise RuntimeError("openmatch_b200 contrastive loss runs on CUDA tensors only (no CPU path)")
        if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
            raise ValueError("expected x [nq, d] and y [np, d]")
        lib = _lib.load()
        dt = torch.bfloat16 if (x.dtype == torch.bfloat16 and y.dtype == torch.bfloat16) else torch.float32
        xc, yc = x.detach().to(dt).contiguous(), y.detach().to(dt).contiguous()
        nq, d = xc.shape
        n_p = yc.shape[0]
        tgt = target.to(torch.int64).contiguous() if target is not None else None
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        need_grad = x.requires_grad or y.requires_grad
        dx = torch.empty((nq, d), dtype=torch.float32, device=x.device) if need_grad else None
        dy = torch.empty((n_p, d), dtype=torch.float32, device=x.device) if need_grad else None
        scores = torch.empty((nq, n_p), dtype=torch.float32, device=x.device) if want_scores else None
        red = {"mean": _lib.OM_REDUCE_MEAN, "sum": _lib.OM_REDUCE_SUM}.get(reduction)
        if red is None:
            raise ValueError("reduction must be 'mean' or 'sum'")
        _lib.check(lib.om_contrastive_loss_fwd_bwd(
            xc.data_ptr(), yc.data_ptr(), _lib.OM_BF16 if dt == torch.bfloat16 else _lib.OM_F32, nq, n_p, d,
            tgt.data_ptr() if tgt is not None else None, red, 1.0, loss.data_ptr(),
            dx.data_ptr() if dx is not None else None, dy.data_ptr() if dy is not None else None,
            scores.data_ptr() if scores is not None else None, _lib.current_stream_ptr()))
        ctx.save_for_backward(dx, dy)
        ctx.in_dtypes = (x.dtype, y.dtype)
        if want_scores:
            ctx.mark_non_differentiable(scores)
            return loss, scores
        return loss, None

    @staticmethod
    def backward(ctx, grad_loss, _grad_scores):
        dx, dy = ctx.saved_tensors
        gx = (dx * grad_loss).to(ctx.in_dtypes[0]) if dx is not None else None
        gy = (dy * grad_loss).to(ctx.in_dtypes[1]) if dy is not None else None
        return gx, gy, None, None, None


def fused_contrastive_loss(x: Tensor, y: Tensor, target: Optional[Tensor] = None, reduction: str = "mean",
                           return_scores: bool = False):
    loss, scores = _FusedContrastive.apply(x, y, target, reduction, return_scores)
    return (loss, scores) if return_scores else loss


class SimpleContrastiveLoss:
    """``loss.py:7-15``: default target is ``i * (y.size(0) // x.size(0))``."""

    def __call__(self, x: Tensor, y: Tensor, target: Tensor = None, reduction: str = 'mean'):
        return fused_contrastive_loss(x, y, target, reduction)


class DistributedContrastiveLoss(SimpleContrastiveLoss):
    """``loss.py:18-38``: all-gather x and y (own slice keeps its autograd history), loss x world_size."""

    def __init__(self, n_target: int = 0, scale_loss: bool = True):
        assert dist.is_initialized(), "Distributed training has not been properly initialized."
        super().__init__()
        self.word_size = dist.get_world_size()
        self.rank = dist.get_rank()
        self.scale_loss = scale_loss

    def __call__(self, x: Tensor, y: Tensor, **kwargs):
        dist_x = self.gather_tensor(x)
        dist_y = self.gather_tensor(y)
        loss = super().__call__(dist_x, dist_y, **kwargs)
        if self.scale_loss:
            loss = loss * self.word_size
        return loss

    def gather_tensor(self, t):
        t = t.contiguous()
        gathered = [torch.empty_like(t) for _ in range(self.word_size)]
        dist.all_gather(gathered, t.detach())
        gathered[self.rank] = t
        return torch.cat(gathered, dim=0)
