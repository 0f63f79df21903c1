"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's output post-processing
 for the parity tests of sfh_postprocess.  Never imported by the product path.

  preds_to_masks      utils/postprocess.py:7-18   softmax -> argmax -> IntTensor -> uint8
  warp mask           predict.py:99               .cpu().numpy().astype(np.uint8)
  mask_type           predict.py:288-299          'rgb' = onehot_to_image (utils/postprocess.py:21-58),
                                                  'bin' = (mask > 0) * 255, 'gray' = as is
  resize              predict.py:303-315          cv2.resize(m, out_size, interpolation=cv2.INTER_NEAREST)

The nearest-neighbour rule is restated from cv::resizeNN (OpenCV 4.x imgproc/resize.cpp):
x_ofs[x] = min(cvFloor(x * ifx), src_w - 1) with ifx = 1 / (dst_w / src_w) in double.  It is pinned
against the real cv2 (4.13 in this image) by tests/test_oracle.py and by the tables committed in
tests/golden/cv2_nearest_tables.npz (tools/make_golden_post.py)."""
import numpy as np
import torch
import torch.nn.functional as F

COLOURS = {1: (0, 255, 0), 2: (255, 0, 0), 3: (0, 0, 255), 4: (255, 255, 255),
           5: (255, 0, 255), 6: (0, 255, 255), 7: (255, 255, 0)}


def preds_to_masks(preds: torch.Tensor, n_classes: int) -> np.ndarray:
    probs = F.softmax(preds, dim=1)
    masks = torch.argmax(probs, dim=1)
    return masks.type(torch.IntTensor).cpu().numpy().astype(np.uint8)


def onehot_to_image(masks: np.ndarray, n_classes: int) -> np.ndarray:
    if n_classes not in (4, 7, 8):
        raise NotImplementedError
    rgb = np.zeros(masks.shape + (3,), dtype=np.uint8)
    for cid in range(1, n_classes):
        rgb[masks == cid] = COLOURS[cid]
    return rgb


def nearest_table(src: int, dst: int) -> np.ndarray:
    ifx = 1.0 / (float(dst) / float(src))
    return np.minimum(np.floor(np.arange(dst, dtype=np.float64) * ifx).astype(np.int64), src - 1)


def resize_nearest(m: np.ndarray, out_size) -> np.ndarray:
    """m: [h,w] or [h,w,3]; out_size = (width, height)."""
    ow, oh = out_size
    yo, xo = nearest_table(m.shape[0], oh), nearest_table(m.shape[1], ow)
    return m[yo][:, xo]


def postprocess(src, kind: str, mask_type: str, out_size, n_classes: int) -> np.ndarray:
    if kind == "logits":
        masks = preds_to_masks(src, n_classes)
    else:
        masks = src.cpu().numpy().astype(np.uint8)
    if mask_type == "rgb":
        masks = onehot_to_image(masks, n_classes)
    elif mask_type == "bin":
        masks = ((masks > 0) * 255).astype(np.uint8)
    elif mask_type != "gray":
        raise NotImplementedError
    if out_size is None:
        return masks
    return np.stack([resize_nearest(m, out_size) for m in masks], axis=0)
