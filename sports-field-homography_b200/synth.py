"""Seeded synthetic inputs for tests and bench: random homographies injected at
the warp boundary, because a randomly initialised Reconstructor emits theta == identity
(models/resnet.py:207-208).  CPU generators so the same bits are used on every device."""
from __future__ import annotations

import numpy as np
import torch


def theta_family_a(B: int, seed: int = 1234, amp: float = 0.15) -> torch.Tensor:
    """I + U(-amp, amp) per entry, [B,1,3,3] fp32 (Z >= 0.55, ~70 % template coverage)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.eye(3)[None, None] + (torch.rand(B, 1, 3, 3, generator=g) * 2 - 1) * amp).float()


def _dlt(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    A = []
    for (x, y), (u, v) in zip(src, dst):
        A.append([-x, -y, -1, 0, 0, 0, u * x, u * y, u])
        A.append([0, 0, 0, -x, -y, -1, v * x, v * y, v])
    _, _, vt = np.linalg.svd(np.asarray(A, np.float64))
    Hm = vt[-1].reshape(3, 3)
    return Hm / Hm[2, 2]


def theta_family_b(B: int, seed: int = 1234) -> torch.Tensor:
    """'Broadcast camera': fp64 DLT from the frame square to a random convex court quadrilateral
    covering 40-100 % of the court, times a scalar U(1,20) (utils/mapping_example.py has
    h22 ~ 13-17), cast to fp32, [B,1,3,3]."""
    rng = np.random.default_rng(seed)
    src = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float64)
    out = []
    while len(out) < B:
        s = rng.uniform(0.63, 1.0)
        c = rng.uniform(-(1 - s), 1 - s, size=2)
        dst = c + s * src + rng.uniform(-0.12, 0.12, size=(4, 2)) * s
        Hm = _dlt(src, dst)
        zs = Hm[2, 0] * src[:, 0] + Hm[2, 1] * src[:, 1] + Hm[2, 2]
        if zs.min() < 0.3:
            continue
        out.append(Hm * rng.uniform(1.0, 20.0))
    return torch.from_numpy(np.stack(out)).float()[:, None]


def perturb(theta: torch.Tensor, sigma: float = 0.01, seed: int = 99) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    scale = theta.reshape(theta.shape[0], -1).abs().amax(dim=1).reshape(-1, 1, 1, 1)
    return theta + torch.randn(theta.shape, generator=g) * sigma * scale
