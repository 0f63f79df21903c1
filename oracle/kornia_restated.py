"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — parity unpinned at the kornia boundary.

Pure-torch restatement of the third-party arithmetic the reference's STN warp stage bottoms
out in.  kornia (``requirements.txt:1`` of the reference: ``kornia>=0.5.0``, constrained to
0.5.x/0.6.x by ``torch==1.8.1``) is not vendored under /root/reference and not installable
here, so its published algorithm is restated from SURVEY.md Appendix A and anchored on the
reference's call sites:

* ``kornia.geometry.transform.HomographyWarper``  <- models/reconstructor.py:105,107,116
* ``kornia.geometry.linalg.transform_points``      <- models/reconstructor.py:4,124

Everything below runs on real ATen ops (``torch.bmm``, ``F.grid_sample``, ``torch.inverse``)
so that, executed on a device, it is bit-for-bit what the reference executes on that device.
fp32 is the parity target; fp64 is the truth used for noise-floor accounting.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- kornia.utils
def create_meshgrid(height: int, width: int, normalized_coordinates: bool = True,
                    device=None, dtype=torch.float32) -> torch.Tensor:
    """kornia.utils.create_meshgrid (0.5.x): [1,H,W,2], last dim (x, y).

    Normalisation is by (size-1): ``(xs / (W-1) - 0.5) * 2``  (SURVEY.md App. A).
    """
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    base = torch.stack(torch.meshgrid([xs, ys], indexing="ij")).transpose(1, 2)  # 2xHxW
    return base.unsqueeze(0).permute(0, 2, 3, 1)


# ------------------------------------------------------------------ kornia.geometry.conversions
def convert_points_from_homogeneous(points: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """0.5.x masked form: scale = 1/z where |z| > eps else 1; result = scale * p[..., :-1]."""
    z_vec = points[..., -1:]
    mask = torch.abs(z_vec) > eps
    # masked_scatter_ form of kornia 0.5.x, written with torch.where so that it is
    # differentiable without a host sync; forward values are identical (IEEE 1/z).
    safe_z = torch.where(mask, z_vec, torch.ones_like(z_vec))
    scale = torch.where(mask, torch.ones_like(z_vec) / safe_z, torch.ones_like(z_vec))
    return scale * points[..., :-1]


def convert_points_to_homogeneous(points: torch.Tensor) -> torch.Tensor:
    return F.pad(points, [0, 1], "constant", 1.0)


# ----------------------------------------------------------------------- kornia.geometry.linalg
def transform_points(trans_01: torch.Tensor, points_1: torch.Tensor) -> torch.Tensor:
    """kornia.geometry.linalg.transform_points (reshape/bmm version, 0.5.x).

    trans_01: [B,3,3] or [B,1,3,3] (any leading dims); points_1: [B,N,2] or [B,H,W,2].
    """
    if not trans_01.device == points_1.device:
        raise TypeError("Tensor must be in the same device")
    if not trans_01.shape[0] == points_1.shape[0] and trans_01.shape[0] != 1:
        raise ValueError("Input batch size must be the same for both tensors or 1")
    if not trans_01.shape[-1] == (points_1.shape[-1] + 1):
        raise ValueError("Last input dimensions must differ by one unit")
    shape_inp = list(points_1.shape)
    points_1 = points_1.reshape(-1, points_1.shape[-2], points_1.shape[-1])
    trans_01 = trans_01.reshape(-1, trans_01.shape[-2], trans_01.shape[-1])
    trans_01 = torch.repeat_interleave(trans_01, repeats=points_1.shape[0] // trans_01.shape[0], dim=0)
    points_1_h = convert_points_to_homogeneous(points_1)
    points_0_h = torch.bmm(points_1_h, trans_01.permute(0, 2, 1))
    points_0_h = torch.squeeze(points_0_h, dim=-1)
    points_0 = convert_points_from_homogeneous(points_0_h)
    shape_inp[-2] = points_0.shape[-2]
    shape_inp[-1] = points_0.shape[-1]
    return points_0.reshape(shape_inp)


# -------------------------------------------------------------------- kornia.geometry.transform
def warp_grid(grid: torch.Tensor, src_homo_dst: torch.Tensor) -> torch.Tensor:
    batch_size = src_homo_dst.shape[0]
    _, height, width, _ = grid.shape
    grid = grid.expand(batch_size, -1, -1, -1)
    if len(src_homo_dst.shape) == 3:
        src_homo_dst = src_homo_dst.view(batch_size, 1, 3, 3)
    flow = transform_points(src_homo_dst, grid.to(src_homo_dst))
    return flow.view(batch_size, height, width, 2)


class HomographyWarper(torch.nn.Module):
    """Restated kornia.geometry.transform.HomographyWarper (normalized_coordinates=True)."""

    def __init__(self, height: int, width: int, mode: str = "bilinear", padding_mode: str = "zeros",
                 normalized_coordinates: bool = True, align_corners: bool = False, grid_dtype=None) -> None:
        super().__init__()
        self.width, self.height = width, height
        self.mode, self.padding_mode = mode, padding_mode
        self.normalized_coordinates = normalized_coordinates
        self.align_corners = align_corners
        # kornia 0.5.x creates the grid in __init__ with the default dtype (fp32) and casts it to the homography's
        # dtype in warp_grid (`grid.to(src_homo_dst)`); later versions build it from the input.  For fp32 inputs the
        # two coincide; for the fp64 Warper of utils/transform.py the choice is stated by the caller (None = input dtype).
        self.grid_dtype = grid_dtype

    def flow(self, patch_src: torch.Tensor, src_homo_dst: torch.Tensor) -> torch.Tensor:
        grid = create_meshgrid(self.height, self.width, self.normalized_coordinates,
                               device=patch_src.device, dtype=self.grid_dtype or patch_src.dtype)
        return warp_grid(grid, src_homo_dst)

    def forward(self, patch_src: torch.Tensor, src_homo_dst: torch.Tensor) -> torch.Tensor:
        if not src_homo_dst.device == patch_src.device:
            raise TypeError("Patch and homography must be on the same device.")
        flow = self.flow(patch_src, src_homo_dst)
        return F.grid_sample(patch_src, flow, mode=self.mode, padding_mode=self.padding_mode,
                             align_corners=self.align_corners)


# ------------------------------------------------------- models/reconstructor.py tails (restated)
def warp(theta: torch.Tensor, court_img: torch.Tensor, height: int, width: int,
         mode: str = "bilinear") -> torch.Tensor:
    """Reconstructor.warp — models/reconstructor.py:109-118."""
    bs = theta.shape[0]
    return HomographyWarper(height, width, mode=mode)(court_img[0:bs], theta).squeeze(1)


def transform_poi(theta: torch.Tensor, court_poi: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """Reconstructor.transform_poi — models/reconstructor.py:120-130."""
    bs = theta.shape[0]
    poi = transform_points(torch.inverse(theta[:bs]), court_poi[:bs])
    if normalize:
        poi = poi / 2.0 + 0.5
    return poi


def predict_tail(theta, court_img, logits, court_poi, mask_classes: int, height: int, width: int,
                 mode: str = "nearest", consistency: bool = True, project_poi: bool = True) -> dict:
    """Reconstructor.predict, warp-stage part — models/reconstructor.py:221-245."""
    ret = {"theta": theta}
    ret["warp_mask"] = warp(theta, court_img, height, width, mode) * mask_classes
    if consistency:
        warp_mask = ret["warp_mask"]
        if logits.shape[2:4] != warp_mask.shape[1:3]:
            h, w = logits.shape[2:4]
            warp_mask = F.interpolate(warp_mask.unsqueeze(1), size=(h, w), mode="nearest").squeeze(1)
        scores = F.cross_entropy(logits, warp_mask.type(torch.int64), reduction="none")
        ret["consist_score"] = torch.mean(scores, dim=(1, 2))
    ret["warp_mask"] = ret["warp_mask"].type(torch.int32)
    if project_poi:
        ret["poi"] = transform_poi(theta, court_poi)
    return ret


# ------------------------------------------------------------------ models/losses.py (restated)
def reprojection_loss(inputs, targets, nonzeros, num_nonzero, reduction: str = "mean"):
    """models/losses.py:6-18."""
    dist = torch.sqrt(torch.sum(torch.pow(targets - inputs, 2), dim=2))
    loss = torch.sum(dist * nonzeros, dim=1) / num_nonzero
    if reduction == "mean":
        loss = torch.mean(loss)
    elif reduction == "sum":
        loss = torch.sum(loss)
    return loss


def per_sample_weighted_criterion(criterion, inputs, targets, per_sample_weights):
    """models/losses.py:33-41 — plain broadcasting of the weights is part of the contract."""
    import types
    if isinstance(criterion, types.FunctionType):
        loss = criterion(inputs, targets, reduction="none")
    else:
        loss = criterion(inputs, targets)
    loss = torch.mean(loss, dim=(1, 2)) * per_sample_weights
    return torch.mean(loss)


def rec_loss_per_sample(warp_mask, gt_masks, mask_classes: int, kind: str = "MSE"):
    """L_b of train.py:194-197 / eval.py:157,186-188 before the weight multiply."""
    gt_f = gt_masks.to(dtype=warp_mask.dtype) / float(mask_classes)
    if kind == "MSE":
        ell = F.mse_loss(warp_mask, gt_f, reduction="none")
    elif kind == "SmoothL1":
        ell = F.smooth_l1_loss(warp_mask, gt_f, reduction="none")
    else:
        raise NotImplementedError(kind)
    return torch.mean(ell, dim=(1, 2))


def one_hot(labels: torch.Tensor, num_classes: int, dtype=None, eps: float = 1e-6) -> torch.Tensor:
    """kornia.utils.one_hot (0.5.x/0.6.x): scatter 1.0 along dim 1, then ``+ eps``."""
    shape = labels.shape
    oh = torch.zeros((shape[0], num_classes) + tuple(shape[1:]), device=labels.device, dtype=dtype)
    return oh.scatter_(1, labels.unsqueeze(1), 1.0) + eps


def focal_loss(input: torch.Tensor, target: torch.Tensor, alpha: float, gamma: float = 2.0,
               reduction: str = "none", eps: float = 1e-8) -> torch.Tensor:
    """kornia.losses.focal_loss as published in kornia 0.5.x/0.6.x (the versions torch 1.8.1 admits,
    requirements.txt:1,5) — used by train.py:106,134 as FocalLoss(alpha=1.0, gamma=2.0, ...).
    softmax + eps, one-hot targets carrying kornia.utils.one_hot's own eps, plain log."""
    input_soft = F.softmax(input, dim=1) + eps
    target_one_hot = one_hot(target, num_classes=input.shape[1], dtype=input.dtype)
    weight = torch.pow(-input_soft + 1.0, gamma)
    focal = -alpha * weight * torch.log(input_soft)
    loss_tmp = torch.sum(target_one_hot * focal, dim=1)
    if reduction == "none":
        return loss_tmp
    if reduction == "mean":
        return torch.mean(loss_tmp)
    if reduction == "sum":
        return torch.sum(loss_tmp)
    raise NotImplementedError(reduction)


def consistency_loss_focal(logits, warp_mask, num_classes: int, consist_lambda: float = 1.0,
                           alpha: float = 1.0, gamma: float = 2.0):
    """train.py:133-134 + 219-223 with ``--consist_loss focal``."""
    rec_masks_int = (warp_mask * num_classes).to(dtype=torch.long)
    if rec_masks_int.ndim == 4:
        rec_masks_int = rec_masks_int[:, 0]
    return focal_loss(logits, rec_masks_int, alpha, gamma, "mean") * consist_lambda


def consistency_loss(logits, warp_mask, num_classes: int, consist_lambda: float = 1.0):
    """train.py:219-223 (and the eval metric, eval.py:201-203): CE between the segmentation logits and
    the class mask obtained by truncating the warped template, ``nn.CrossEntropyLoss()`` defaults
    (reduction 'mean' over every pixel of the batch).  ``warp_mask`` is [B,1,H,W] as the warper
    returns it; train.py indexes it the same way through the squeeze in Reconstructor.forward."""
    rec_masks_int = (warp_mask * num_classes).to(dtype=torch.long)
    if rec_masks_int.ndim == 4:
        rec_masks_int = rec_masks_int[:, 0]
    return F.cross_entropy(logits, rec_masks_int) * consist_lambda
