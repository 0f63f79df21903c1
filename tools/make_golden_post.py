#!/usr/bin/env python
"""Writes tests/golden/cv2_nearest_tables.npz: the source-index tables cv2.resize(INTER_NEAREST) really
uses for a set of (src, dst) sizes, recovered by resizing an index ramp with the cv2 of this image.
Run in the build container (cv2 4.13)."""
import os
import cv2
import numpy as np

pairs = [(640, 1280), (360, 720), (1280, 640), (720, 360), (640, 1920), (360, 1080), (1280, 1920), (720, 1080),
         (640, 100), (360, 77), (77, 360), (100, 640), (640, 641), (641, 640), (1280, 854), (720, 480), (3, 7), (7, 3)]
out = {"cv2_version": np.array(cv2.__version__)}
for s, d in pairs:
    ramp = np.arange(s, dtype=np.float32)[None, :].repeat(2, 0)          # values are their own column index
    out[f"{s}_{d}"] = cv2.resize(ramp, (d, 2), interpolation=cv2.INTER_NEAREST)[0].astype(np.int32)
    col = np.arange(s, dtype=np.float32)[:, None].repeat(2, 1)
    assert np.array_equal(cv2.resize(col, (2, d), interpolation=cv2.INTER_NEAREST)[:, 0].astype(np.int32), out[f"{s}_{d}"])
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cv2_nearest_tables.npz"), **out)
print("wrote", len(pairs), "tables, cv2", cv2.__version__)
