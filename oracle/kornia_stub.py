"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Installs a stand-in ``kornia`` package into ``sys.modules`` whose ``HomographyWarper`` and
``transform_points`` are the restatements in ``oracle/kornia_restated.py``.  With it, the
UNMODIFIED reference files (``/root/reference/models/reconstructor.py`` etc., which do
``import kornia`` at import time, reconstructor.py:3-4) can be imported and executed in this
container to generate the golden vectors under ``tests/golden/`` (tools/make_golden.py) and to
run the drop-in test of ``patch_reconstructor`` against the real ``Reconstructor`` class.
"""
from __future__ import annotations

import sys
import types

from . import kornia_restated as _kr


def install() -> types.ModuleType:
    if "kornia" in sys.modules and not getattr(sys.modules["kornia"], "__sfh_stub__", False):
        return sys.modules["kornia"]          # a real kornia is present: use it
    kornia = types.ModuleType("kornia")
    kornia.__sfh_stub__ = True
    geometry = types.ModuleType("kornia.geometry")
    transform = types.ModuleType("kornia.geometry.transform")
    linalg = types.ModuleType("kornia.geometry.linalg")
    losses = types.ModuleType("kornia.losses")
    utils = types.ModuleType("kornia.utils")

    transform.HomographyWarper = _kr.HomographyWarper
    linalg.transform_points = _kr.transform_points
    utils.create_meshgrid = _kr.create_meshgrid

    class FocalLoss:  # placeholder: segmentation loss, outside the hot path
        def __init__(self, *a, **k):
            raise NotImplementedError("kornia.losses.FocalLoss is out of scope for the stub")
    losses.FocalLoss = FocalLoss

    kornia.geometry, kornia.losses, kornia.utils = geometry, losses, utils
    geometry.transform, geometry.linalg = transform, linalg
    geometry.transform_points = _kr.transform_points
    for m in (kornia, geometry, transform, linalg, losses, utils):
        sys.modules[m.__name__] = m
    return kornia
