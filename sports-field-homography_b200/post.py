"""GPU post-processing of the predict outputs: what predict.py does on the CPU
after the device->host copy — ``preds_to_masks`` (utils/postprocess.py:7-18), ``.astype(np.uint8)``
(predict.py:99), mask_type conversion (predict.py:288-299, utils/postprocess.py:21-58) and the
``cv2.resize(..., INTER_NEAREST)`` to ``out_size`` (predict.py:303-315) — done in one launch on the
device so that only uint8 at the output size crosses PCIe."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .warper import _stream

MASK_TYPES = {"gray": 0, "bin": 1, "rgb": 2}
_TABLES = {}


def cv2_nearest_table(src: int, dst: int) -> np.ndarray:
    """Source index of every destination index under cv2.INTER_NEAREST (cv::resizeNN):
    ``min(floor(x * (1 / (dst / src))), src - 1)`` evaluated in double."""
    inv_scale = float(dst) / float(src)
    ifx = 1.0 / inv_scale
    return np.minimum(np.floor(np.arange(dst, dtype=np.float64) * ifx).astype(np.int64), src - 1).astype(np.int32)


def _table(src, dst, device):
    if src == dst:
        return None
    key = (src, dst, device.index)
    t = _TABLES.get(key)
    if t is None:
        t = _TABLES[key] = torch.from_numpy(cv2_nearest_table(src, dst)).to(device)
    return t


def postprocess_masks(src, mask_type="gray", out_size=None, n_classes=4, out=None):
    """``src``: logits ``[B,nc,h,w]`` fp32 (-> segmentation mask by argmax) or a class mask ``[B,h,w]``
    int32 / uint8 (the warp mask of ``predict_tail``).  ``out_size`` is ``(width, height)`` as in
    predict.py's ``args.out_size``; None keeps the size.  Returns uint8 ``[B,oh,ow]`` ('gray', 'bin') or
    ``[B,oh,ow,3]`` ('rgb'), on the device."""
    if not isinstance(src, torch.Tensor) or not src.is_cuda:
        raise TypeError("src must be a CUDA tensor")
    if mask_type not in MASK_TYPES:
        raise NotImplementedError(mask_type)                 # predict.py:300-301
    if src.ndim == 4:
        if src.dtype != torch.float32:
            raise TypeError("logits must be float32")
        kind, (B, nc, h, w) = 0, src.shape
        if nc != n_classes:
            raise ValueError("logits must have n_classes channels")
    elif src.ndim == 3:
        if src.dtype == torch.int32:
            kind = 1
        elif src.dtype == torch.uint8:
            kind = 2
        else:
            raise TypeError("class masks must be int32 or uint8")
        (B, h, w), nc = src.shape, int(n_classes)
    else:
        raise ValueError("src must be [B,nc,h,w] logits or a [B,h,w] class mask")
    if mask_type == "rgb" and nc not in (4, 7, 8):
        raise NotImplementedError("onehot_to_image knows 4, 7 or 8 classes")   # utils/postprocess.py:56-57
    ow, oh = (w, h) if out_size is None else (int(out_size[0]), int(out_size[1]))
    src = src.contiguous()
    shape = (B, oh, ow, 3) if mask_type == "rgb" else (B, oh, ow)
    if out is None or tuple(out.shape) != shape or out.dtype != torch.uint8 or out.device != src.device \
            or not out.is_contiguous():
        out = torch.empty(shape, dtype=torch.uint8, device=src.device)
    xo, yo = _table(w, ow, src.device), _table(h, oh, src.device)
    with torch.cuda.device(src.device):
        rc = _lib.lib().sfh_postprocess(src.data_ptr(), kind, B, nc, h, w, MASK_TYPES[mask_type],
                                        xo.data_ptr() if xo is not None else None,
                                        yo.data_ptr() if yo is not None else None, oh, ow,
                                        out.data_ptr(), _stream())
    _lib.check(rc, "sfh_postprocess")
    return out
