"""ctypes binding of libsfh_b200.so (C ABI declared in include/sfh_b200.h).

There is no CPU fallback: if the library is missing or a tensor is not on a CUDA device the
call fails loudly.  The library itself never sees a torch type — only raw device pointers,
sizes and the current stream handle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("SFH_LIB_PATH") or os.path.join(_HERE, "libsfh_b200.so")   # override: development A/B only
CSRC = [os.path.join(_HERE, "csrc", f) for f in ("sfh_warp.cu", "sfh_poi.cu", "sfh_consist.cu", "sfh_post.cu", "sfh_aux.cu")]
HDRS = [os.path.join(_HERE, "csrc", f) for f in ("sfh_device.cuh", "sfh_poi.cuh")] + \
       [os.path.join(_ROOT, "include", "sfh_b200.h")]

MODE = {"bilinear": 0, "nearest": 1}
LOSS = {"MSE": 0, "SmoothL1": 1}
TMPL_F32, TMPL_Q2, TMPL_Q4 = 0, 1, 2


class SfhTemplate(C.Structure):
    """struct sfh_template (include/sfh_b200.h)."""
    _fields_ = [("data", C.c_void_p), ("fmt", C.c_int32), ("channels", C.c_int32),
                ("height", C.c_int32), ("width", C.c_int32), ("pitch", C.c_int32),
                ("n_palette", C.c_int32), ("batch_stride", C.c_int64), ("palette", C.c_float * 16),
                ("sat", C.c_void_p), ("sat_pitch", C.c_int32)]


class SfhTrainTailArgs(C.Structure):
    """struct sfh_train_tail_args (include/sfh_b200.h)."""
    _fields_ = [("theta", C.c_void_p), ("xs", C.c_void_p), ("ys", C.c_void_p), ("gt", C.c_void_p),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("nc", C.c_int32),
                ("kind", C.c_int32), ("N", C.c_int32),
                ("warp_out", C.c_void_p), ("L_b", C.c_void_p), ("dLb_dtheta", C.c_void_p),
                ("court_poi", C.c_void_p), ("court_poi_bstride", C.c_int64),
                ("gt_poi", C.c_void_p), ("nonzeros", C.c_void_p), ("num_nonzero", C.c_void_p),
                ("poi_out", C.c_void_p), ("R_b", C.c_void_p), ("dRb_dtheta", C.c_void_p),
                ("weights", C.c_void_p), ("weights_f64", C.c_int32), ("weights_outer", C.c_int32),
                ("gt_dtype", C.c_int32), ("rec_lambda", C.c_float), ("reproj_lambda", C.c_float),
                ("loss_out", C.c_void_p), ("dtheta_total", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


class SfhPredictTailArgs(C.Structure):
    """struct sfh_predict_tail_args (include/sfh_b200.h)."""
    _fields_ = [("theta", C.c_void_p), ("xs", C.c_void_p), ("ys", C.c_void_p),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("mode", C.c_int32),
                ("nc", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("N", C.c_int32),
                ("logits", C.c_void_p), ("warp_out", C.c_void_p), ("score", C.c_void_p),
                ("court_poi", C.c_void_p), ("court_poi_bstride", C.c_int64), ("poi_out", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("mask_dtype", C.c_int32)]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_T = C.POINTER(SfhTemplate)

# name -> (restype, argtypes); must list every symbol include/sfh_b200.h declares
SIGNATURES = {
    "sfh_abi_version": (_I, []),
    "sfh_build_info": (C.c_char_p, []),
    "sfh_error_string": (C.c_char_p, [_I]),
    "sfh_workspace_bytes": (_L, [_I, _I, _I]),
    "sfh_template_pack": (_I, [_P, _I, _I, C.POINTER(C.c_float), _I, _P, _I, _I, _P, _P, _I, _P]),
    "sfh_warp_fwd": (_I, [_P, _T, _P, _P, _I, _I, _I, _I, _P, _P]),
    "sfh_forward_tail": (_I, [_P, _T, _P, _P, _I, _I, _I, _I, _P, _P, _L, _I, _P, _P]),
    "sfh_warp_bwd": (_I, [_P, _T, _P, _P, _P, _I, _I, _I, _P, _P, _L, _P]),
    "sfh_warp_loss_fwd_bwd": (_I, [_T, C.POINTER(SfhTrainTailArgs), _P]),
    "sfh_predict_tail": (_I, [_T, C.POINTER(SfhPredictTailArgs), _P]),
    "sfh_poi_fwd": (_I, [_P, _P, _L, _I, _I, _I, _P, _P]),
    "sfh_poi_bwd": (_I, [_P, _P, _L, _P, _I, _I, _I, _P, _P]),
    "sfh_transform_points_fwd": (_I, [_P, _I, _P, _I, _I, _P, _P]),
    "sfh_transform_points_bwd": (_I, [_P, _I, _P, _P, _I, _I, _P, _P, _P]),
    "sfh_reproj_loss": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "sfh_consist_workspace_bytes": (_L, []),
    "sfh_consist_loss_fwd_bwd": (_I, [_P, _P, _I, _I, _I, _I, C.c_float, _P, _P, _P, _L, _P]),
    "sfh_consist_focal_fwd_bwd": (_I, [_P, _P, _I, _I, _I, _I, C.c_float, C.c_float, C.c_float, _P, _P, _P, _L, _P]),
    "sfh_postprocess": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "sfh_warp_nearest_f64": (_I, [_P, _P, _L, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "sfh_warp_perspective_nearest": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    "sfh_selftest_rcp": (_I, [_P, _P]),
    "sfh_debug_stream_cast": (_I, [_P, _P, _L, _I, _P]),
}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = CSRC + HDRS
    if not force and os.path.exists(LIB_PATH) and \
            os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", os.path.join(_ROOT, "include"), "-o", LIB_PATH] + CSRC
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libsfh_b200.so")
    return LIB_PATH


_lib = None


def lib():
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
                " sfh_b200 has no CPU or eager fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError => header/ABI mismatch, fail loudly
            fn.restype, fn.argtypes = res, args
        if l.sfh_abi_version() != 1:
            raise RuntimeError("libsfh_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().sfh_error_string(rc).decode()
        if rc < 0:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: CUDA error {rc}: {msg}")
