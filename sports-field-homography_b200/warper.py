"""kornia-compatible front end of the B200 warp kernels.

``HomographyWarper`` and ``transform_points`` keep the constructor / call signatures and the
error conventions of the kornia objects the reference uses
(models/reconstructor.py:4,105,107,116,124): ``TypeError`` for device / dtype problems,
``ValueError`` for shape problems, ``RuntimeError`` for CUDA failures.  All compute happens in
libsfh_b200.so; torch only owns the memory and the stream.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .court import CourtTemplate


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def check_theta(theta: torch.Tensor, device=None) -> torch.Tensor:
    """[B,1,3,3] or [B,3,3] fp32 CUDA -> contiguous [B,9] view (kornia accepts both ranks)."""
    if not isinstance(theta, torch.Tensor):
        raise TypeError(f"Input type is not a torch.Tensor. Got {type(theta)}")
    if theta.dtype != torch.float32:
        raise TypeError(f"theta must be float32, got {theta.dtype}")
    if not theta.is_cuda:
        raise TypeError("theta must live on a CUDA device (sfh_b200 has no CPU path)")
    if device is not None and theta.device != device:
        raise TypeError("Patch and homography must be on the same device.")
    if not (theta.shape[-2:] == (3, 3) and theta.ndim in (3, 4) and (theta.ndim == 3 or theta.shape[1] == 1)):
        raise ValueError(f"Invalid input homography shape, we expect Bx3x3 or Bx1x3x3. Got: {tuple(theta.shape)}")
    return theta.reshape(theta.shape[0], 9).contiguous()


class _Workspace:
    """Per-device scratch for the per-sample reductions; zero-filled once, kernels re-zero it."""

    def __init__(self):
        self.buf = {}

    def get(self, device, B, H, W):
        need = int(_lib.lib().sfh_workspace_bytes(B, H, W))
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        cur = self.buf.get(key)
        if cur is None or cur.numel() < need:
            cur = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=device)
            self.buf[key] = cur
        return cur


_WS = _Workspace()


def meshgrid_factors(height: int, width: int, device, dtype=torch.float32, grid_source: str = "device"):
    """1-D factors of kornia's create_meshgrid(normalized_coordinates=True), built with the same
    torch ops kornia uses.  ATen's CUDA ``tensor / python_scalar`` multiplies by the reciprocal
    while its CPU one divides, so the two devices disagree by an ulp on some columns (measured on
    B200: that is the ONLY difference between the reference run on GPU and on CPU; everything
    downstream is replayed bit for bit by the kernels).  ``grid_source='device'`` (default)
    reproduces the reference executed on this GPU, ``'cpu'`` the reference executed on the host."""
    if grid_source not in ("device", "cpu"):
        raise ValueError("grid_source must be 'device' or 'cpu'")
    build_on = device if grid_source == "device" else torch.device("cpu")
    xs = torch.linspace(0, width - 1, width, device=build_on, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=build_on, dtype=dtype)
    xs = (xs / (width - 1) - 0.5) * 2
    ys = (ys / (height - 1) - 0.5) * 2
    return xs.to(device).contiguous(), ys.to(device).contiguous()


class _WarpFn(torch.autograd.Function):
    """out = warp(template, theta); backward -> dtheta only (the template carries no gradient)."""

    @staticmethod
    def forward(ctx, theta9, tmpl: CourtTemplate, xs, ys, H, W, mode, shortcut):
        B = theta9.shape[0]
        out = torch.empty((B, tmpl.C, H, W), dtype=torch.float32, device=theta9.device)
        d = tmpl.desc(shortcut)
        with torch.cuda.device(theta9.device):
            rc = _lib.lib().sfh_warp_fwd(theta9.data_ptr(), d, _ptr(xs), _ptr(ys), B, H, W,
                                         _lib.MODE[mode], out.data_ptr(), _stream())
        _lib.check(rc, "sfh_warp_fwd")
        ctx.save_for_backward(theta9)
        ctx.tmpl, ctx.xs, ctx.ys, ctx.mode, ctx.hw, ctx.shortcut = tmpl, xs, ys, mode, (H, W), shortcut
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (theta9,) = ctx.saved_tensors
        B = theta9.shape[0]
        H, W = ctx.hw
        if ctx.mode == "nearest":           # piecewise-constant in theta: autograd gives zeros too
            return torch.zeros_like(theta9), None, None, None, None, None, None, None
        grad_out = grad_out.contiguous()
        if grad_out.dtype != torch.float32:
            raise TypeError("grad_out must be float32")
        dth = torch.empty_like(theta9)
        ws = _WS.get(theta9.device, B, H, W)
        d = ctx.tmpl.desc(ctx.shortcut)
        with torch.cuda.device(theta9.device):
            rc = _lib.lib().sfh_warp_bwd(theta9.data_ptr(), d, _ptr(ctx.xs), _ptr(ctx.ys),
                                         grad_out.data_ptr(), B, H, W, dth.data_ptr(),
                                         ws.data_ptr(), ws.numel(), _stream())
        _lib.check(rc, "sfh_warp_bwd")
        return dth, None, None, None, None, None, None, None


class HomographyWarper(torch.nn.Module):
    """Drop-in for ``kornia.geometry.transform.HomographyWarper`` (normalized coordinates,
    padding 'zeros', align_corners False — the configuration models/reconstructor.py:105,107
    constructs).  ``forward(patch_src, src_homo_dst) -> [B,C,H,W]``.

    A constant template can be staged once with ``set_template`` (packed palette format); a
    ``patch_src`` passed to ``forward`` that is not that template is sampled as plain fp32.
    """

    def __init__(self, height: int, width: int, mode: str = "bilinear", padding_mode: str = "zeros",
                 normalized_coordinates: bool = True, align_corners: bool = False,
                 grid_source: str = "device", exact: bool = False) -> None:
        super().__init__()
        # exact=True: every pixel is evaluated in ATen's fp32 operation order (bit-identical floats).
        # exact=False (default): 16x8 output patches whose sampling footprint contains no class edge
        # of a staged palette template are written as the class constant — bit-exact for 'nearest',
        # within 2 ulp (2e-7) of ATen's interpolated constant for 'bilinear', gradients unchanged.
        self.exact = bool(exact)
        if grid_source not in ("device", "cpu"):
            raise ValueError("grid_source must be 'device' or 'cpu'")
        self.grid_source = grid_source
        if mode not in ("bilinear", "nearest"):
            raise NotImplementedError(f"mode={mode!r}: only 'bilinear' and 'nearest' are implemented")
        if padding_mode != "zeros" or not normalized_coordinates or align_corners:
            raise NotImplementedError("only padding_mode='zeros', normalized_coordinates=True, "
                                      "align_corners=False (the reference's configuration) are implemented")
        self.height, self.width = int(height), int(width)
        self.mode, self.padding_mode = mode, padding_mode
        self.normalized_coordinates, self.align_corners = normalized_coordinates, align_corners
        self._grid = {}          # device -> (xs, ys); plain attributes, nothing enters state_dict
        self._staged = None      # (key, CourtTemplate)

    @property
    def edge_shortcut(self) -> bool:
        return self.mode == "nearest" or not self.exact

    def grid_factors(self, device):
        g = self._grid.get(device)
        if g is None:
            g = meshgrid_factors(self.height, self.width, device, grid_source=self.grid_source)
            self._grid[device] = g
        return g

    def set_template(self, court_img: torch.Tensor) -> CourtTemplate:
        """Stage a constant template (the reference's court_img) once for this warper."""
        t = CourtTemplate(court_img, shared=True, pack=True)
        self._staged = ((court_img.data_ptr(), court_img._version, tuple(court_img.shape)), t)
        return t

    def _template_for(self, patch_src: torch.Tensor, B: int) -> CourtTemplate:
        if self._staged is not None:
            (ptr, ver, shape), t = self._staged
            if patch_src.data_ptr() == ptr and patch_src._version == ver and tuple(patch_src.shape[1:]) == shape[1:]:
                return t
        if patch_src.shape[0] != B:
            raise ValueError(f"batch size mismatch: patch_src {patch_src.shape[0]} vs homography {B}")
        return CourtTemplate(patch_src, shared=False, pack=False)

    def forward(self, patch_src: torch.Tensor, src_homo_dst: torch.Tensor) -> torch.Tensor:
        if not isinstance(patch_src, torch.Tensor):
            raise TypeError(f"Input type is not a torch.Tensor. Got {type(patch_src)}")
        if not isinstance(src_homo_dst, torch.Tensor):
            raise TypeError(f"Input type is not a torch.Tensor. Got {type(src_homo_dst)}")
        if not src_homo_dst.device == patch_src.device:
            raise TypeError("Patch and homography must be on the same device.")
        if patch_src.ndim != 4:
            raise ValueError(f"Invalid input shape, we expect BxCxHxW. Got: {tuple(patch_src.shape)}")
        if patch_src.requires_grad:
            raise NotImplementedError("gradient w.r.t. the template is not implemented "
                                      "(the reference's court_img carries no gradient)")
        theta9 = check_theta(src_homo_dst, patch_src.device)
        tmpl = self._template_for(patch_src, theta9.shape[0])
        xs, ys = self.grid_factors(patch_src.device)
        return _WarpFn.apply(theta9, tmpl, xs, ys, self.height, self.width, self.mode, self.edge_shortcut)


class _TransformPointsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trans9, pts):
        Bt, (B, N) = trans9.shape[0], pts.shape[:2]
        out = torch.empty_like(pts)
        with torch.cuda.device(pts.device):
            rc = _lib.lib().sfh_transform_points_fwd(trans9.data_ptr(), Bt, pts.data_ptr(), B, N,
                                                     out.data_ptr(), _stream())
        _lib.check(rc, "sfh_transform_points_fwd")
        ctx.save_for_backward(trans9, pts)
        return out

    @staticmethod
    def backward(ctx, g):
        trans9, pts = ctx.saved_tensors
        Bt, (B, N) = trans9.shape[0], pts.shape[:2]
        g = g.contiguous()
        dtr = torch.empty_like(trans9) if ctx.needs_input_grad[0] else None
        dpt = torch.empty_like(pts) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(pts.device):
            rc = _lib.lib().sfh_transform_points_bwd(trans9.data_ptr(), Bt, pts.data_ptr(), g.data_ptr(),
                                                     B, N, _ptr(dtr), _ptr(dpt), _stream())
        _lib.check(rc, "sfh_transform_points_bwd")
        return dtr, dpt


def transform_points(trans_01: torch.Tensor, points_1: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``kornia.geometry.linalg.transform_points`` for 2-D points:
    trans_01 [B,3,3] / [B,1,3,3] / [1,3,3], points_1 [B,N,2] -> [B,N,2]."""
    if not isinstance(trans_01, torch.Tensor) or not isinstance(points_1, torch.Tensor):
        raise TypeError("Input type is not a torch.Tensor")
    if not trans_01.device == points_1.device:
        raise TypeError("Tensor must be in the same device")
    if not points_1.is_cuda:
        raise TypeError("points must live on a CUDA device (sfh_b200 has no CPU path)")
    if trans_01.dtype != torch.float32 or points_1.dtype != torch.float32:
        raise TypeError("transform_points: float32 tensors expected")
    if not trans_01.shape[0] == points_1.shape[0] and trans_01.shape[0] != 1:
        raise ValueError("Input batch size must be the same for both tensors or 1")
    if not trans_01.shape[-1] == (points_1.shape[-1] + 1) or points_1.shape[-1] != 2:
        raise ValueError("Last input dimensions must differ by one unit")
    shape = points_1.shape
    pts = points_1.reshape(shape[0], -1, 2).contiguous()
    tr = trans_01.reshape(-1, 9).contiguous()
    if tr.shape[0] not in (1, shape[0]):
        raise ValueError("Input batch size must be the same for both tensors or 1")
    return _TransformPointsFn.apply(tr, pts).reshape(shape)
