"""Batch sharding of the warp stage across the GPUs of one box.

The path shards by frame with no data-path collective.  The only exchange is one all-reduce of
a short vector of loss / metric numerators per step, because the reference's losses and metrics
are batch means (models/losses.py:13-14,38-39; eval.py:219-224).  With the ``[B]*[B,1]`` weight
quirk the global loss is ``mean(w) * mean(L_b)``, so sum(L_b), sum(w), sum(L_b*w) and the counts
are reduced separately and the caller picks the form it needs.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for `rank` (first n%world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_sums(vec: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a short fp64 vector of numerators/counts over ranks (NCCL on GPU, gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def global_means(rec_per_sample=None, weights=None, reproj_per_sample=None, consist_score=None, group=None):
    """Global batch statistics from local per-sample terms.

    Returns a dict with (when the inputs are given)
      rec_mean        mean_b L_b                         (eval.py:196 path, unweighted)
      rec_weighted    mean_b (L_b * w_b)                 (models/losses.py:38-39, w of shape [B])
      rec_quirk       mean(w) * mean(L_b)                (same lines, w of shape [B,1])
      reproj_mean     mean_b R_b                         (models/losses.py:13-14)
      consist_mean    mean_b score_b
      n               global number of frames
    """
    ref = next(t for t in (rec_per_sample, reproj_per_sample, consist_score) if t is not None)
    z = torch.zeros((), dtype=torch.float64, device=ref.device)
    f = lambda t: t.double().sum() if t is not None else z
    n = torch.tensor(float(ref.shape[0]), dtype=torch.float64, device=ref.device)
    w = None if weights is None else weights.reshape(-1).to(ref.device)
    lw = (rec_per_sample.double() * w.double()).sum() if (w is not None and rec_per_sample is not None) else z
    vec = torch.stack([f(rec_per_sample), f(w), lw, f(reproj_per_sample), f(consist_score), n])
    vec = reduce_sums(vec, group)
    n = vec[5]
    out = {"n": n}
    if rec_per_sample is not None:
        out["rec_mean"] = vec[0] / n
        if w is not None:
            out["rec_weighted"] = vec[2] / n
            out["rec_quirk"] = (vec[1] / n) * (vec[0] / n)
    if reproj_per_sample is not None:
        out["reproj_mean"] = vec[3] / n
    if consist_score is not None:
        out["consist_mean"] = vec[4] / n
    return out
