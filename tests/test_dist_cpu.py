"""world_size-2 gloo test of the multi-GPU plumbing: batch sharding + the one
all-reduce of loss / metric numerators.  Sharded global means must equal the serial ones,
including the reference's [B]*[B,1] weight-broadcast quirk (models/losses.py:38-39)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    import sfh_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    Lb, w = torch.rand(B, generator=g), torch.rand(B, generator=g) + 0.5
    Rb, sc = torch.rand(B, generator=g), torch.rand(B, generator=g)
    lo, hi = sfh_b200.shard_range(B, rank, world)
    out = sfh_b200.global_means(Lb[lo:hi], w[lo:hi], Rb[lo:hi], sc[lo:hi])
    if rank == 0:
        q.put({k: float(v) for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_means_equal_serial_means():
    B, world, port = 37, 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    Lb, w = torch.rand(B, generator=g), torch.rand(B, generator=g) + 0.5
    Rb, sc = torch.rand(B, generator=g), torch.rand(B, generator=g)
    assert got["n"] == B
    assert abs(got["rec_mean"] - Lb.double().mean().item()) < 1e-12
    assert abs(got["rec_weighted"] - torch.mean(Lb.double() * w.double()).item()) < 1e-12
    # per_sample_weighted_criterion with w of shape [B,1]: mean over the [B,B] outer product
    quirk = torch.mean(Lb.double() * w.double().reshape(-1, 1)).item()
    assert abs(got["rec_quirk"] - quirk) < 1e-12
    assert abs(got["reproj_mean"] - Rb.double().mean().item()) < 1e-12
    assert abs(got["consist_mean"] - sc.double().mean().item()) < 1e-12
