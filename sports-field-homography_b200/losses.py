"""Loss surface of the warp stage, named as in the reference's models/losses.py.

``reprojection_loss`` / ``ReprojectionLoss`` (models/losses.py:6-29) run on the sfh kernel and
are differentiable w.r.t. ``inputs``; ``per_sample_weighted_criterion`` (models/losses.py:33-41)
and its tail ``weight_and_reduce`` keep the reference's plain-broadcast weighting so the
``[B]*[B,1] -> [B,B]`` quirk is preserved when fed the fused per-sample losses.
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


def per_sample_weighted_criterion(criterion, inputs, targets, per_sample_weights):
    """models/losses.py:33-41, same signature: ``criterion`` is an ``nn.MSELoss(reduction='none')`` /
    ``nn.SmoothL1Loss(reduction='none')``-like callable; elementwise loss, mean over (1,2), plain-broadcast
    weighting (``[B]*[B,1] -> [B,B]`` quirk preserved), batch mean.  Runs on torch ops: it is the generic,
    unfused form; ``STNWarpStage.train_tail`` fuses it with the warp for the two reference criteria."""
    loss = criterion(inputs, targets)
    loss = torch.mean(loss, dim=(1, 2))
    loss = loss * per_sample_weights
    return torch.mean(loss)


def weight_and_reduce(per_sample_loss, per_sample_weights):
    """Last two lines of per_sample_weighted_criterion (models/losses.py:38-39) applied to an
    already-reduced per-sample loss: ``mean(L_b * w)`` with plain broadcasting."""
    return torch.mean(per_sample_loss * per_sample_weights)


# ---------------------------------------------------------------------------------------------
# Consistency loss: train.py:219-223, eval.py:201-203
# ---------------------------------------------------------------------------------------------
_CONSIST_WS = {}


def _consist_ws(device):
    """Small per-(device, stream) workspace: one ticket + one fp32 partial per CTA; allocated zeroed,
    left zeroed by the kernel."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _CONSIST_WS.get(key)
    if ws is None:
        ws = torch.zeros(int(_lib.lib().sfh_consist_workspace_bytes()), dtype=torch.uint8, device=device)
        _CONSIST_WS[key] = ws
    return ws


def _check_consist_args(logits, warp_mask, num_classes):
    for name, t in (("logits", logits), ("warp_mask", warp_mask)):
        if not isinstance(t, torch.Tensor) or t.dtype != torch.float32 or not t.is_cuda:
            raise TypeError(f"{name} must be a float32 CUDA tensor")
    if logits.ndim != 4 or logits.shape[1] != num_classes:
        raise ValueError("logits must be [B,num_classes,H,W]")
    B, _, H, W = logits.shape
    if warp_mask.numel() != B * H * W or tuple(warp_mask.shape[-2:]) != (H, W):
        raise ValueError("warp_mask must be [B,1,H,W] (or [B,H,W]) with the logits' H, W")
    if not 1 <= num_classes <= 8:
        raise ValueError("num_classes must be in 1..8")


def consistency_step(logits, warp_mask, num_classes, consist_lambda=1.0, need_grad=True, out=None,
                     criterion="CE", alpha=1.0, gamma=2.0):
    """``consist_lambda * CrossEntropyLoss()(logits, (warp_mask * num_classes).long())`` and its
    gradient w.r.t. ``logits`` in ONE streaming launch (no log_softmax tensor, no int64 mask).
    ``criterion='focal'`` selects train.py:133-134's ``kornia.losses.FocalLoss(alpha, gamma, 'mean')`` instead
    (kornia 0.5.x eps conventions, see include/sfh_b200.h).
    Returns ``{"loss": scalar tensor, "dlogits": [B,nc,H,W] or None}``; ``out`` reuses buffers."""
    if criterion not in ("CE", "focal"):
        raise NotImplementedError(criterion)
    _check_consist_args(logits, warp_mask, num_classes)
    B, nc, H, W = logits.shape
    logits = logits.contiguous()
    warp_mask = warp_mask.contiguous()
    out = {} if out is None else out
    loss = out.get("loss")
    if loss is None or loss.device != logits.device:
        loss = out["loss"] = torch.empty((), dtype=torch.float32, device=logits.device)
    dl = None
    if need_grad:
        dl = out.get("dlogits")
        if dl is None or dl.shape != logits.shape or dl.device != logits.device:
            dl = out["dlogits"] = torch.empty_like(logits)
    ws = _consist_ws(logits.device)
    with torch.cuda.device(logits.device):
        if criterion == "CE":
            rc = _lib.lib().sfh_consist_loss_fwd_bwd(warp_mask.data_ptr(), logits.data_ptr(), B, nc, H, W,
                                                     float(consist_lambda), loss.data_ptr(),
                                                     dl.data_ptr() if dl is not None else None,
                                                     ws.data_ptr(), ws.numel(), _stream())
        else:
            rc = _lib.lib().sfh_consist_focal_fwd_bwd(warp_mask.data_ptr(), logits.data_ptr(), B, nc, H, W,
                                                      float(alpha), float(gamma), float(consist_lambda), loss.data_ptr(),
                                                      dl.data_ptr() if dl is not None else None,
                                                      ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "sfh_consist_loss_fwd_bwd" if criterion == "CE" else "sfh_consist_focal_fwd_bwd")
    return {"loss": loss, "dlogits": dl}


class _ConsistFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, warp_mask, num_classes, consist_lambda, criterion, alpha, gamma):
        r = consistency_step(logits, warp_mask, num_classes, consist_lambda, need_grad=logits.requires_grad,
                             criterion=criterion, alpha=alpha, gamma=gamma)
        ctx.dl = r["dlogits"]
        return r["loss"].clone()

    @staticmethod
    def backward(ctx, g):
        if ctx.dl is None:                      # logits did not require grad at forward time
            return (None,) * 7
        return (ctx.dl * g.to(torch.float32),) + (None,) * 6        # out of place: backward may run twice


def consistency_loss(logits, warp_mask, num_classes, consist_lambda=1.0, criterion="CE", alpha=1.0, gamma=2.0):
    """Differentiable (w.r.t. ``logits``) form of train.py:221-222 for both of the reference's consistency
    criteria (train.py:131-134: 'CE' -> nn.CrossEntropyLoss(), 'focal' -> kornia FocalLoss(alpha=1, gamma=2,
    'mean')); no gradient reaches ``warp_mask`` — the reference's ``.to(dtype=torch.long)`` cuts it the same way."""
    return _ConsistFn.apply(logits, warp_mask.detach(), int(num_classes), float(consist_lambda),
                            criterion, float(alpha), float(gamma))
