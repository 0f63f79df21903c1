"""The reference's offline consumers of the warp, on the device.

``Warper``                     == utils/transform.py:7-20 (fp64 multi-channel nearest HomographyWarper)
``warp_perspective_nearest``   == cv2.warpPerspective(src, M, size, flags=cv2.INTER_NEAREST) as the dataset tooling
                                  uses it (dataset_utils/football_dataset.ipynb cell 11), batched over many M
``rescale_theta``              == dataset_utils/preparation.py:129-137
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .warper import _stream


def rescale_theta(src_size, dst_size, theta):
    """dataset_utils/preparation.py:129-137: diag(dst_w, dst_h, 1) @ theta @ diag(1/src_w, 1/src_h, 1), fp64
    (theta may carry leading batch dimensions)."""
    theta = np.asarray(theta, dtype=np.float64)
    left = np.array([dst_size[0], dst_size[1], 1.0], dtype=np.float64)[:, None]
    right = np.array([1.0 / src_size[0], 1.0 / src_size[1], 1.0], dtype=np.float64)[None, :]
    return theta * left * right


def meshgrid_factors_f64(height: int, width: int, device, grid_dtype=torch.float32):
    """fp64 meshgrid factors for the fp64 warper.  kornia 0.5.x builds HomographyWarper's grid once in ``__init__``
    with the default dtype (fp32) and casts it to the homography's dtype at call time — ``grid_dtype=torch.float32``
    (default) reproduces that; ``torch.float64`` builds the factors in double (later kornia versions build the grid
    from the input's dtype).  The two differ by ~1e-8, which moves a nearest pick only at an exact tie."""
    xs = torch.linspace(0, width - 1, width, dtype=grid_dtype)
    ys = torch.linspace(0, height - 1, height, dtype=grid_dtype)
    xs = (xs / (width - 1) - 0.5) * 2
    ys = (ys / (height - 1) - 0.5) * 2
    return xs.to(torch.float64).to(device).contiguous(), ys.to(torch.float64).to(device).contiguous()


class Warper:
    """utils/transform.py:7-20, same constructor / method signatures: ``Warper(size).warp(theta, proj)`` takes a
    numpy ``theta`` [3,3] and ``proj`` [H,W,C] and returns the nearest-warped ``proj`` as fp64 numpy [H,W,C].
    There is no CPU path: ``cuda=False`` raises."""

    def __init__(self, size, cuda=True, grid_dtype=torch.float32):
        if not cuda or not torch.cuda.is_available():
            raise RuntimeError("sfh_b200.Warper runs on a CUDA device only (no CPU fallback)")
        self.width, self.height = int(size[0]), int(size[1])
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.xs, self.ys = meshgrid_factors_f64(self.height, self.width, self.device, grid_dtype)

    def warp_tensor(self, theta: torch.Tensor, proj: torch.Tensor) -> torch.Tensor:
        """Device form: theta [B,3,3] fp64, proj [B|1,C,Hc,Wc] fp64 -> [B,C,H,W] fp64."""
        for name, t in (("theta", theta), ("proj", proj)):
            if not isinstance(t, torch.Tensor) or t.dtype != torch.float64 or not t.is_cuda:
                raise TypeError(f"{name} must be a float64 CUDA tensor")
        if theta.ndim != 3 or tuple(theta.shape[1:]) != (3, 3) or proj.ndim != 4 or proj.shape[0] not in (1, theta.shape[0]):
            raise ValueError("theta must be [B,3,3] and proj [B|1,C,Hc,Wc]")
        theta, proj = theta.contiguous(), proj.contiguous()
        B, (_, C, Hc, Wc) = theta.shape[0], proj.shape
        out = torch.empty((B, C, self.height, self.width), dtype=torch.float64, device=theta.device)
        with torch.cuda.device(theta.device):
            rc = _lib.lib().sfh_warp_nearest_f64(theta.data_ptr(), proj.data_ptr(), 0 if proj.shape[0] == 1 else C * Hc * Wc,
                                                 self.xs.data_ptr(), self.ys.data_ptr(), B, C, Hc, Wc,
                                                 self.height, self.width, out.data_ptr(), _stream())
        _lib.check(rc, "sfh_warp_nearest_f64")
        return out

    def warp(self, theta, proj):
        proj = torch.from_numpy(np.ascontiguousarray(proj)).type(torch.DoubleTensor).permute(2, 0, 1).unsqueeze(0)
        theta = torch.from_numpy(np.asarray(theta, dtype=np.float64)).reshape(1, 3, 3)
        out = self.warp_tensor(theta.to(self.device), proj.to(self.device))[0]
        return out.permute(1, 2, 0).cpu().numpy()


def warp_perspective_nearest(src, M, dsize, out: torch.Tensor = None) -> torch.Tensor:
    """``cv2.warpPerspective(src, M, dsize, flags=cv2.INTER_NEAREST)`` (border constant 0) for one or many ``M``.

    src: CUDA tensor [Hs,Ws] or [Hs,Ws,C] of any fixed-size dtype (uint8 BGR masks, fp64 UV maps, ...);
    M: [3,3] or [B,3,3] source->destination homographies in PIXEL coordinates (what ``rescale_theta`` returns; numpy
    or tensor); dsize (W,H).  Returns [B,H,W(,C)] of src's dtype (a single M gives [H,W(,C)])."""
    if not isinstance(src, torch.Tensor) or not src.is_cuda:
        raise TypeError("src must be a CUDA tensor (sfh_b200 has no CPU path)")
    if src.ndim not in (2, 3):
        raise ValueError("src must be [Hs,Ws] or [Hs,Ws,C]")
    Mn = np.asarray(M.detach().cpu().numpy() if isinstance(M, torch.Tensor) else M, dtype=np.float64)
    single = Mn.ndim == 2
    Mn = Mn.reshape(-1, 3, 3)
    minv = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(Mn))).to(src.device)      # cv2 inverts M itself (in double)
    src = src.contiguous()
    Hs, Ws = src.shape[:2]
    pix = src.element_size() * (src.shape[2] if src.ndim == 3 else 1)
    W, H = int(dsize[0]), int(dsize[1])
    shape = (Mn.shape[0], H, W) + tuple(src.shape[2:])
    if out is None or tuple(out.shape) != shape or out.dtype != src.dtype or out.device != src.device or not out.is_contiguous():
        out = torch.empty(shape, dtype=src.dtype, device=src.device)
    with torch.cuda.device(src.device):
        rc = _lib.lib().sfh_warp_perspective_nearest(minv.data_ptr(), Mn.shape[0], src.data_ptr(), Hs, Ws, pix, H, W,
                                                     out.data_ptr(), _stream())
    _lib.check(rc, "sfh_warp_perspective_nearest")
    return out[0] if single else out
