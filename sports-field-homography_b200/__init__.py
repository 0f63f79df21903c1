"""sfh_b200 — B200-native (sm_100a) STN warp stage for darkAlert/sports-field-homography.

Only the Reconstructor's warp stage is implemented here (models/reconstructor.py:100-130,
185-192, 221-245 + models/losses.py of the reference): template warp by predicted homographies,
the warp-mask MSE/SmoothL1 loss with dL/dtheta, POI reprojection + RRMSE, and the consistency
score.  Compute lives in libsfh_b200.so (hand-written CUDA behind a C ABI, include/sfh_b200.h);
this package is the Python mirror of the reference's call surface.  There is no CPU fallback.
"""
from . import _lib
from .court import CourtTemplate, load_bundled, open_court_poi, open_court_template
from .dist import global_means, reduce_sums, shard_range
from .losses import (ReprojectionLoss, consistency_loss, consistency_step, per_sample_weighted_criterion,
                     reprojection_loss, reprojection_per_sample, weight_and_reduce)
from .mapping import map_court_to_frame, map_frame_to_court
from .post import cv2_nearest_table, postprocess_masks
from .stage import STNWarpStage, patch_reconstructor
from .tooling import Warper, rescale_theta, warp_perspective_nearest
from .warper import HomographyWarper, meshgrid_factors, transform_points

__all__ = [
    "HomographyWarper", "transform_points", "meshgrid_factors", "STNWarpStage", "patch_reconstructor",
    "CourtTemplate", "open_court_template", "open_court_poi", "load_bundled",
    "reprojection_loss", "reprojection_per_sample", "ReprojectionLoss", "weight_and_reduce",
    "per_sample_weighted_criterion",
    "consistency_loss", "consistency_step", "postprocess_masks", "cv2_nearest_table",
    "map_frame_to_court", "map_court_to_frame", "Warper", "warp_perspective_nearest", "rescale_theta",
    "shard_range", "reduce_sums", "global_means",
]
__version__ = "0.1.0"
