"""Frame <-> court point mapping of the reference's ``utils/transform.py:25-55`` (``transform_poi``,
``map_frame_to_court``, ``map_court_to_frame``: numpy + ``cv2.perspectiveTransform`` on the host), batched
on the device through the library's ``transform_points`` kernel (third consumer).

The reference divides pixel locations by the frame / court size, maps them to [-1, 1], applies the 3x3
homography and returns points in [0, 1] (``normalize=True``).  Here ``theta`` may be one matrix ``[3,3]``
or a batch ``[B,3,3]`` and locations ``[N,2]`` / ``[B,N,2]``; everything is fp32 on the CUDA device."""
from __future__ import annotations

import torch

from .warper import transform_points


def _as_batched(theta, loc):
    if not isinstance(theta, torch.Tensor) or not isinstance(loc, torch.Tensor):
        raise TypeError("theta and locations must be torch tensors")
    if theta.shape[-2:] != (3, 3):
        raise ValueError("theta must be [3,3] or [B,3,3]")
    if loc.shape[-1] != 2:
        raise ValueError("locations must be [N,2] or [B,N,2]")
    th = theta.reshape(-1, 3, 3).to(torch.float32)
    pts = loc.to(torch.float32)
    squeeze = pts.ndim == 2
    if squeeze:
        pts = pts.unsqueeze(0)
    if th.shape[0] not in (1, pts.shape[0]):
        raise ValueError("theta batch must be 1 or match the locations")
    return th, pts, squeeze


def transform_poi(theta, poi, normalize=False):
    """utils/transform.py:25-33: ``cv2.perspectiveTransform(poi, theta)``, optionally ``/ 2 + 0.5``."""
    th, pts, squeeze = _as_batched(theta, poi)
    out = transform_points(th, pts.contiguous())
    if normalize:
        out = out / 2.0 + 0.5
    return out[0] if squeeze else out


def _map(theta, loc, size):
    th, pts, squeeze = _as_batched(theta, loc)
    if size is not None:
        scale = torch.tensor([float(size[0]), float(size[1])], dtype=torch.float32, device=pts.device)
        pts = (pts / scale - 0.5) * 2.0
    out = transform_points(th, pts.contiguous()) / 2.0 + 0.5
    return out[0] if squeeze else out


def map_frame_to_court(theta_f2c, frame_loc, frame_size=None):
    """utils/transform.py:36-44: frame pixel locations -> court coordinates in [0,1]."""
    return _map(theta_f2c, frame_loc, frame_size)


def map_court_to_frame(theta_c2f, court_loc, court_size=None):
    """utils/transform.py:47-55: court pixel locations -> frame coordinates in [0,1]."""
    return _map(theta_c2f, court_loc, court_size)
