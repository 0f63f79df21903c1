"""Loss surface of the warp stage, named as in the reference's models/losses.py.

``reprojection_loss`` / ``ReprojectionLoss`` (models/losses.py:6-29) run on the sfh kernel and
are differentiable w.r.t. ``inputs``; ``per_sample_weighted_criterion`` (models/losses.py:33-41)
keeps the reference's plain-broadcast weighting so the ``[B]*[B,1] -> [B,B]`` quirk
 is preserved when it is fed the fused per-sample losses.
"""
from __future__ import annotations

import torch

from . import _lib
from .warper import _ptr, _stream


class _ReprojFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets, nonzeros, num_nonzero):
        B, N = inputs.shape[:2]
        Rb = torch.empty(B, dtype=torch.float32, device=inputs.device)
        with torch.cuda.device(inputs.device):
            rc = _lib.lib().sfh_reproj_loss(inputs.data_ptr(), targets.data_ptr(), nonzeros.data_ptr(),
                                            num_nonzero.data_ptr(), B, N, Rb.data_ptr(), None, None, _stream())
        _lib.check(rc, "sfh_reproj_loss")
        ctx.save_for_backward(inputs, targets, nonzeros, num_nonzero)
        return Rb

    @staticmethod
    def backward(ctx, g):
        inputs, targets, nonzeros, num_nonzero = ctx.saved_tensors
        B, N = inputs.shape[:2]
        g = g.contiguous().to(torch.float32)
        din = torch.empty_like(inputs)
        with torch.cuda.device(inputs.device):
            rc = _lib.lib().sfh_reproj_loss(inputs.data_ptr(), targets.data_ptr(), nonzeros.data_ptr(),
                                            num_nonzero.data_ptr(), B, N, None, g.data_ptr(),
                                            din.data_ptr(), _stream())
        _lib.check(rc, "sfh_reproj_loss(bwd)")
        return din, None, None, None


def reprojection_per_sample(inputs, targets, nonzeros, num_nonzero):
    """[B]: sum_n ||targets-inputs|| * nonzeros / num_nonzero  (models/losses.py:10-11)."""
    for name, t in (("inputs", inputs), ("targets", targets), ("nonzeros", nonzeros), ("num_nonzero", num_nonzero)):
        if not isinstance(t, torch.Tensor) or t.dtype != torch.float32 or not t.is_cuda:
            raise TypeError(f"{name} must be a float32 CUDA tensor")
    if inputs.ndim != 3 or inputs.shape[-1] != 2 or targets.shape != inputs.shape:
        raise ValueError("inputs/targets must be [B,N,2]")
    if tuple(nonzeros.shape) != tuple(inputs.shape[:2]) or num_nonzero.numel() != inputs.shape[0]:
        raise ValueError("nonzeros must be [B,N] and num_nonzero [B]")
    return _ReprojFn.apply(inputs.contiguous(), targets.contiguous(), nonzeros.contiguous(),
                           num_nonzero.contiguous())


def reprojection_loss(inputs, targets, nonzeros, num_nonzero, reduction="mean"):
    """models/losses.py:6-18."""
    loss = reprojection_per_sample(inputs, targets, nonzeros, num_nonzero)
    if reduction == "mean":
        loss = torch.mean(loss)
    elif reduction == "sum":
        loss = torch.sum(loss)
    return loss


class ReprojectionLoss(torch.nn.Module):
    """models/losses.py:21-29."""

    def forward(self, inputs, targets, nonzeros, num_nonzero, reduction="mean"):
        return reprojection_loss(inputs, targets, nonzeros, num_nonzero, reduction)


def weight_and_reduce(per_sample_loss, per_sample_weights):
    """Last two lines of per_sample_weighted_criterion (models/losses.py:38-39) applied to an
    already-reduced per-sample loss: ``mean(L_b * w)`` with plain broadcasting."""
    return torch.mean(per_sample_loss * per_sample_weights)
