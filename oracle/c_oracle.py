"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): ctypes/numpy front end of warp_oracle.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsfh_oracle.so")
_lib = None

MODE = {"bilinear": 0, "nearest": 1}
LOSS = {"MSE": 0, "SmoothL1": 1}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "warp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def meshgrid(n: int) -> np.ndarray:
    out = np.empty(n, np.float32)
    lib().sfh_oracle_meshgrid(C.c_int(n), _p(out))
    return out


def _tmpl(tmpl, B):
    tmpl = _f(tmpl)
    assert tmpl.ndim == 4
    bstride = 0 if tmpl.shape[0] == 1 else tmpl[0].size
    assert tmpl.shape[0] in (1, B) or tmpl.shape[0] >= B
    return tmpl, bstride


def warp_fwd(theta, tmpl, H, W, mode="bilinear", xs=None, ys=None):
    theta = _f(theta).reshape(-1, 9)
    B = theta.shape[0]
    tmpl, bs = _tmpl(tmpl, B)
    _, Cc, Hc, Wc = tmpl.shape
    out = np.empty((B, Cc, H, W), np.float32)
    xs = None if xs is None else _f(xs)
    ys = None if ys is None else _f(ys)
    lib().sfh_oracle_warp_fwd(_p(theta), _p(tmpl), C.c_long(bs), B, Cc, Hc, Wc, H, W, MODE[mode],
                              _p(xs), _p(ys), _p(out))
    return out


def warp_bwd(theta, tmpl, grad_out, xs=None, ys=None):
    theta = _f(theta).reshape(-1, 9)
    B = theta.shape[0]
    tmpl, bs = _tmpl(tmpl, B)
    _, Cc, Hc, Wc = tmpl.shape
    grad_out = _f(grad_out)
    _, _, H, W = grad_out.shape
    dth = np.empty((B, 9), np.float32)
    xs = None if xs is None else _f(xs)
    ys = None if ys is None else _f(ys)
    lib().sfh_oracle_warp_bwd(_p(theta), _p(tmpl), C.c_long(bs), _p(grad_out), B, Cc, Hc, Wc, H, W,
                              _p(xs), _p(ys), _p(dth))
    return dth.reshape(B, 3, 3)


def warp_loss(theta, tmpl, gt, nc, kind="MSE", xs=None, ys=None):
    theta = _f(theta).reshape(-1, 9)
    B = theta.shape[0]
    tmpl, bs = _tmpl(tmpl, B)
    _, Cc, Hc, Wc = tmpl.shape
    assert Cc == 1
    gt = np.ascontiguousarray(gt, dtype=np.int64)
    _, H, W = gt.shape
    warp = np.empty((B, H, W), np.float32)
    Lb = np.empty(B, np.float32)
    J = np.empty((B, 9), np.float32)
    xs = None if xs is None else _f(xs)
    ys = None if ys is None else _f(ys)
    lib().sfh_oracle_warp_loss(_p(theta), _p(tmpl), C.c_long(bs), _p(gt), nc, LOSS[kind], B, Hc, Wc,
                               H, W, _p(xs), _p(ys), _p(warp), _p(Lb), _p(J))
    return warp, Lb, J.reshape(B, 3, 3)


def predict_tail(theta, tmpl, logits, nc, H, W, mode="nearest", xs=None, ys=None, score=True):
    theta = _f(theta).reshape(-1, 9)
    B = theta.shape[0]
    tmpl, bs = _tmpl(tmpl, B)
    _, Cc, Hc, Wc = tmpl.shape
    assert Cc == 1
    logits = _f(logits)
    _, ncl, h, w = logits.shape
    assert ncl == nc
    warp = np.empty((B, H, W), np.int32)
    sc = np.empty(B, np.float32) if score else None
    xs = None if xs is None else _f(xs)
    ys = None if ys is None else _f(ys)
    lib().sfh_oracle_predict_tail(_p(theta), _p(tmpl), C.c_long(bs), _p(logits), nc, h, w, B, Hc, Wc,
                                  H, W, MODE[mode], _p(xs), _p(ys), _p(warp), _p(sc))
    return warp, sc


def poi_fwd(theta, court_poi, normalize=True):
    theta = _f(theta).reshape(-1, 9)
    B = theta.shape[0]
    court_poi = _f(court_poi)
    N = court_poi.shape[1]
    out = np.empty((B, N, 2), np.float32)
    lib().sfh_oracle_poi_fwd(_p(theta), _p(court_poi), B, N, int(normalize), _p(out))
    return out


def reproj_per_sample(poi, gt_poi, nonzeros, num_nonzero):
    poi, gt_poi, nonzeros, num_nonzero = _f(poi), _f(gt_poi), _f(nonzeros), _f(num_nonzero)
    B, N, _ = poi.shape
    Lb = np.empty(B, np.float32)
    lib().sfh_oracle_reproj_per_sample(_p(poi), _p(gt_poi), _p(nonzeros), _p(num_nonzero), B, N, _p(Lb))
    return Lb
