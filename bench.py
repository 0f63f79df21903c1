#!/usr/bin/env python
"""bench.py — throughput of the STN warp stage (BASELINE.json metric: warped frames/s, fwd+bwd, 1280x720).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload hd]

A "step" is one pass of the hot path over one batch of synthetic input (random homographies injected at the
warp boundary, real court template).  The default workload `hd` is the configuration the metric
is quoted on: the training warp forward+backward of BASELINE.json configs[1] (warp_mask + MSE loss vs int64 gt
+ POI reprojection RMSE + weighted batch-mean loss and dL/dtheta, ONE kernel launch per step) at the metric's
1280x720, batch 64 per GPU.  configs[1] at its own 640x360 (`c2`) is printed as a full peer entry
(`peers.c2`: own value, e2e, roofline, cpu_baseline, stock-torch arm, parity), the other BASELINE configs as
kernel-level entries (`other_workloads`), C5 (65,536 frames, batch-sharded) as `c5`.
Work is batch-sharded: every rank processes its own 64 frames (weak scaling); the only exchange is one NCCL
all-reduce of the loss numerators per step, issued on a side stream and never on the critical path.

value         whole-job frames/s with inputs resident in HBM (steps replayed from CUDA graphs so the Python
              launch cost does not gate a 50-150 us kernel); buffers rotate through > 4x L2
e2e           same metric through the public API with HOST (pinned) inputs: H2D of the step's inputs and D2H
              of loss + dtheta inside the timed region
roofline      algorithmic bytes per launch / kernel duration (CUDA events on the launching stream) vs
              MEASURED_PEAKS.json hbm_gbs
cpu_baseline  the restated kornia path (oracle/, torch CPU, all host threads) on a bounded sample
stock_torch_b200  the same restated path executed by stock PyTorch ON the B200 (~20 ATen launches, autograd):
              the reference's de-facto GPU implementation
parity        one batch slice of the TIMED stage (same mode, same buffers) against the C oracle, in this run
--impl reference   the CPU path as its own arm (the reference is pure Python + kornia; kornia is not installable
              offline, so the arm runs the oracle port of it — DESIGN.md)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "warped frames/s (fwd+bwd, 1280x720)"
_TRAIN = ("warp_mask fp32 + MSE vs int64 gt + POI RRMSE + weighted batch-mean loss + dL/dtheta "
          "(NCAA v4 nc4 template, theta family A)")
WORKLOADS = {
    # name: (W, H, B per GPU, template, kind, description, algorithmic bytes per frame)
    "hd": (1280, 720, 64, "ncaa_nc4", "train",
           "training warp fwd+bwd at the metric's 1280x720, batch 64/GPU (BASELINE configs[1] work): " + _TRAIN,
           1280 * 720 * 12),
    "c2": (640, 360, 64, "ncaa_nc4", "train",
           "C2 training warp fwd+bwd 640x360 batch 64/GPU: " + _TRAIN, 640 * 360 * 12),
    "c1": (640, 360, 16, "ncaa_nc4", "fwd",
           "C1 bilinear forward 640x360 batch 16, fp32 mask out", 640 * 360 * 4),
    "c3": (1280, 720, 15, "ncaa_nc4", "predict",
           "C3 predict tail 1280x720 batch 15: nearest warp -> int32 mask + CE consistency vs logits [4,360,640] + POI",
           1280 * 720 * 4 + 4 * 360 * 640 * 4),
    "c5": (1280, 720, 256, "ncaa_nc4", "predict",
           "C5 video-scale sweep: 65,536 frames of C3 work, batch-sharded (65,536 / n_gpus frames per rank) in "
           "micro-batches of 256 frames; output/logit buffers reused",
           1280 * 720 * 4 + 4 * 360 * 640 * 4),
    "c4": (1280, 720, 32, "pitch_v3_nc4", "fwd",
           "C4 pitch v3 HD template bilinear forward 1280x720 batch 32 + POI (one launch)", 1280 * 720 * 4),
    "consist": (640, 360, 64, "ncaa_nc4", "consist",
                "training consistency loss 640x360 batch 64 (SURVEY 8 f-2, train.py:219-223): CE(logits [4,360,640], "
                "trunc(warp_mask*4)) + dlogits in one launch", 640 * 360 * 4 + 2 * 4 * 360 * 640 * 4),
}
WORKLOADS["c2hd"] = WORKLOADS["hd"]          # round-1 name of the headline workload
PEERS = ("c2",)                               # printed with the full set of fields next to the headline
OTHERS = ("c1", "c3", "c4", "consist")        # kernel-level entries
L2_BYTES = 126 * 1024 * 1024
C5_FRAMES = 65536
L2_NOTE = ("inputs larger than L2: every step reads / writes its own buffer set, the sets rotate through > 4x the "
           "126 MB L2; the court template (<= 1 MB packed) stays cache-resident by design")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hd", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-extra", action="store_true", help="headline only: skip peers / other workloads / baselines")
    return ap.parse_args()


def config_of(name):
    """The `config` object of the JSON line: identical in both arms (the driver compares them)."""
    W, H, B, tname, kind, desc, bpf = WORKLOADS[name]
    return {"workload": desc, "frames_per_step_per_gpu": B, "size": [W, H], "template": tname, "l2": L2_NOTE}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.stop, self.thread, self.h = [], threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self.stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def finish(self, t0, t1):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop.set()
        self.thread.join()
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "sw_power_cap": 0x4, "hw_power_brake": 0x80, "sync_boost": 0x10}
        bits = 0
        for s in inside:
            bits |= s[2]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": statistics.median([s[1] for s in inside]) if inside else None,
                "sm_max_mhz": mx, "samples": len(inside),
                "reasons": [k for k, v in names.items() if bits & v]}


# ------------------------------------------------------- the reference path (oracle) on a torch device
def reference_step(W, H, B, kind, tmpl, poi, device="cpu", seed=0):
    """The reference's path for one batch: restated kornia ops on stock torch (CPU: the cpu_baseline / reference
    arm; cuda: the stock-torch-on-B200 arm), autograd backward.  Returns a callable running one step on B frames."""
    import torch
    from oracle import kornia_restated as kr
    from sfh_b200 import synth
    dev = torch.device(device)
    th0 = synth.theta_family_a(B, 1234 + seed)
    tm = tmpl.expand(B, -1, -1, -1).contiguous()
    pp = poi.expand(B, -1, -1).contiguous()
    with torch.no_grad():
        gt = (kr.warp(synth.perturb(th0), tm, H, W, "nearest") * 4).to(torch.int64).to(dev)
        gt_poi = kr.transform_poi(synth.perturb(th0, seed=5), pp).to(dev)
    th0, tm, pp = th0.to(dev), tm.to(dev), pp.to(dev)
    nz = torch.ones(B, pp.shape[1], device=dev)
    num = nz.sum(1)
    w = torch.ones(B, dtype=torch.float64, device=dev)
    logits = torch.randn(B, 4, 360, 640, generator=torch.Generator().manual_seed(3)).to(dev) \
        if kind in ("predict", "consist") else None
    wm = kr.warp(th0, tm, H, W, "bilinear").detach() if kind == "consist" else None

    def step():
        if kind == "consist":                   # train.py:219-223 + the backward to the logits
            lg = logits.clone().requires_grad_(True)
            loss = kr.consistency_loss(lg, wm, 4)
            loss.backward()
            return float(loss.detach())
        if kind == "train":
            th = th0.clone().requires_grad_(True)
            warp = kr.warp(th, tm, H, W, "bilinear")
            p = kr.transform_poi(th, pp)
            loss = kr.per_sample_weighted_criterion(torch.nn.MSELoss(reduction="none"), warp,
                                                    gt.to(torch.float32) / 4.0, w) \
                + 8.0 * kr.reprojection_loss(p, gt_poi, nz, num)
            loss.backward()
            return float(loss.detach()) + float(th.grad.sum())     # result read back, like the e2e leg
        with torch.no_grad():
            if kind == "fwd":
                return float(kr.warp(th0, tm, H, W, "bilinear").sum()) + float(kr.transform_poi(th0, pp).sum())
            r = kr.predict_tail(th0, tm, logits, pp, 4, H, W, "nearest")
            return float(r["consist_score"].sum())
    return step


def run_cpu(workload, steps, warmup, sample_frames=16, budget_s=None):
    import torch
    import sfh_b200
    W, H, B, name, kind, desc, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tmpl, poi = sfh_b200.load_bundled(name, (W, H), 4, 1)
    Bs = min(B, sample_frames)
    if budget_s is not None:       # size the per-step sample so steps+warmup fit the time budget
        probe = reference_step(W, H, 4, kind, tmpl, poi)
        probe()
        t0 = time.perf_counter()
        probe()
        per_frame = (time.perf_counter() - t0) / 4
        Bs = int(max(1, min(B, budget_s / ((steps + warmup) * per_frame))))
    step = reference_step(W, H, Bs, kind, tmpl, poi)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": Bs / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps x {Bs} frames of workload {workload} ({W}x{H}, {kind}), "
                      f"oracle/kornia_restated.py on torch CPU with {cores} threads, {dt * 1e3:.1f} ms/step",
            "ms_per_step": dt * 1e3, "frames_per_step": Bs}


def run_stock_torch_gpu(workload, dev, steps=5, warmup=2):
    """The restated kornia path executed by stock PyTorch on the B200 (full batch, CUDA events)."""
    import torch
    import sfh_b200
    W, H, B, name, kind, desc, bpf = WORKLOADS[workload]
    tmpl, poi = sfh_b200.load_bundled(name, (W, H), 4, 1)
    try:
        step = reference_step(W, H, B, kind, tmpl, poi, device=str(dev))
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"value": B / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms, "frames_per_step": B,
               "what": "oracle/kornia_restated.py (create_meshgrid -> transform_points -> grid_sample + the reference's "
                       "losses, autograd backward) on cuda with stock ATen/cuBLAS kernels; includes the reference's own "
                       "host syncs (boolean-mask indexing) and the loss read-back",
               "achieved_GBps_algorithmic": B * bpf / (ms * 1e-3) / 1e9}
    except Exception as e:          # e.g. out of memory on a small device
        out = {"error": repr(e)}
    torch.cuda.empty_cache()
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    r = run_cpu(args.workload, steps, warm, budget_s=120.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args.workload),
            "device": "host CPU", "frames_per_step_sampled": r["frames_per_step"],
            "note": "reference = kornia path restated on torch CPU ops (kornia itself is not installable offline); "
                    "each step is a bounded sample of the batch sized so steps+warmup end within ~2 minutes",
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------- GPU arm
class Workload:
    """Device-resident rotating buffer sets + the public-API call of one step."""

    def __init__(self, name, dev, seed, nsets=None):
        import torch
        import sfh_b200
        from sfh_b200 import synth
        self.torch = torch
        self.name = name
        self.W, self.H, self.B, tname, self.kind, self.desc, self.bytes_per_frame = WORKLOADS[name]
        W, H, B = self.W, self.H, self.B
        tmpl, poi = sfh_b200.load_bundled(tname, (W, H), 4, 1)
        self.tmpl_cpu, self.poi_cpu = tmpl, poi
        self.stage = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4,
                                           warp_with_nearest=(self.kind == "predict"))
        stn = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4, warp_with_nearest=True)
        stb = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4)
        step_bytes = B * self.bytes_per_frame
        self.nsets = nsets or max(2, -(-4 * L2_BYTES // step_bytes))
        self.sets = []
        for i in range(self.nsets):
            th = synth.theta_family_a(B, 1234 + 17 * seed + i).to(dev)
            s = {"theta": th, "out": {}}
            if self.kind == "train":
                s["gt"] = stn.predict_tail(synth.perturb(th.cpu(), seed=i).to(dev), None, False, False)["warp_mask"].to(torch.int64)
                s["gt_poi"] = stb.transform_poi(synth.perturb(th.cpu(), seed=100 + i).to(dev)).detach()
                nz = (torch.rand(B, poi.shape[1], generator=torch.Generator().manual_seed(i)) < 0.8).float()
                nz[:, 0] = 1.0
                s["nz"], s["num"] = nz.to(dev), nz.sum(1).to(dev)
                s["w"] = torch.ones(B, dtype=torch.float64, device=dev)      # utils/dataset.py:220
            elif self.kind == "predict":
                s["logits"] = torch.randn(B, 4, 360, 640, device=dev)
            elif self.kind == "consist":
                s["logits"] = torch.randn(B, 4, H, W, device=dev)
                s["wm"] = stb.warp(th)
            self.sets.append(s)
        del stn, stb

    def step(self, i):
        s = self.sets[i % self.nsets]
        if self.kind == "train":
            return self.stage.train_step(s["theta"], s["gt"], s["w"], "MSE", s["gt_poi"], s["nz"], s["num"],
                                         1.0, 8.0, True, s["out"])
        if self.kind == "predict":
            return self.stage.predict_tail(s["theta"], s["logits"], True, True, s["out"])
        if self.kind == "consist":
            import sfh_b200
            return sfh_b200.consistency_step(s["logits"], s["wm"], 4, 1.0, True, s["out"])
        if self.W == 1280:                    # C4: warp + POI, the forward tail of Reconstructor.forward in one launch
            return self.stage.forward_tail(s["theta"])
        return {"warp_mask": self.stage.warp(s["theta"])}

    def launches_per_step(self):
        # one fused launch per step; SFH_TWO_LAUNCH=1 (development switch) adds the finalize launch
        two = bool(os.environ.get("SFH_TWO_LAUNCH")) and self.kind in ("train", "predict")
        return 2 if two else 1


def time_workload(wl, steps, warmup, use_graph, dist_ring=None, isolated=True):
    """Returns (ms_per_step over the whole timed region, mean kernel us from per-launch event pairs, mode, (t0,t1)).

    Multi-GPU: after every step the step's loss numerators (a 2-float slot the fused tail wrote itself) are
    all-reduced by NCCL — one collective PER STEP — on NCCL's own stream, forked after the step; nothing on the
    device waits for it until the end of the captured graph (>= 8 steps), so it runs beside the next steps."""
    torch = wl.torch
    import torch.distributed as dist
    multi = dist_ring is not None
    has_sum = wl.kind in ("train", "predict", "consist")
    spg = wl.nsets                                     # steps per graph
    if multi:
        spg = wl.nsets * max(1, min(24, max(steps, wl.nsets)) // wl.nsets)
    pending = []

    def one(i, j=None):
        """step i; j = index inside the graph being captured (selects the all-reduce slot)."""
        if multi and has_sum:
            slot = dist_ring[(j if j is not None else i) % dist_ring.shape[0]]
            if wl.kind == "train":                     # the fused tail writes its scalar loss straight into the slot
                wl.sets[i % wl.nsets]["out"]["loss"] = slot[0]
        r = wl.step(i)
        if multi and has_sum:
            if wl.kind == "predict":
                slot[0].copy_(r["consist_score"].sum())      # metric numerator of the inference sweep
            elif wl.kind == "consist":
                slot[0].copy_(r["loss"])
            pending.append(dist.all_reduce(slot, async_op=True))
        return r

    def drain():
        for w in pending:
            w.wait()
        pending.clear()

    graph = None
    mode = "eager"
    with torch.no_grad():
        for i in range(max(warmup, wl.nsets)):
            one(i)
            if len(pending) > 8:
                pending.pop(0).wait()
        drain()
        torch.cuda.synchronize()
        if use_graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    # one eager step on the capture stream first: the library keeps one reduction workspace per
                    # (device, stream); allocated inside the capture, its zero-fill would be replayed with every graph
                    one(0, 0)
                    drain()
                    side.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        for j in range(spg):
                            one(j, j)
                        drain()                        # the only join: end of the graph
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g.replay()
                torch.cuda.synchronize()
                graph = g
                mode = f"cuda_graph ({spg} steps per graph)"
            except Exception as e:      # capture unsupported in this configuration: time eagerly
                sys.stderr.write(f"[bench] graph capture failed ({e!r}); timing eager launches\n")
                graph = None
                pending.clear()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        if graph is not None:
            for _ in range(steps // spg):
                graph.replay()
            for i in range(steps % spg):        # remainder so that EXACTLY `steps` steps are timed
                one(i)
            drain()
        else:
            for i in range(steps):
                one(i)
                if len(pending) > 8:
                    pending.pop(0).wait()
            drain()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if multi:
            dist.barrier()
        ms = e0.elapsed_time(e1) / steps
        kern_us = None
        if isolated:
            # per-launch kernel duration: an event pair around every launch, same stream, same rotation
            n = min(max(steps, 20), 200)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            for i, (a, b) in enumerate(evs):
                torch.cuda._sleep(400000)   # GPU stays busy while the CPU queues a / launch / b, so the pair
                a.record()                  # brackets only the kernel, not Python's launch latency
                wl.step(i)
                b.record()
            torch.cuda.synchronize()
            durs = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
            kern_us = statistics.mean(durs[: max(1, int(0.9 * n))])     # drop the slowest 10 % (launch hiccups)
    return ms, kern_us, mode, (t0, t1), spg


def time_e2e(wl, steps, warmup, u8_masks=False):
    """Public API with HOST inputs: per step H2D of that step's inputs (pinned) and D2H of
    loss + dtheta (train) / score + poi (predict) / a checksum (fwd)."""
    torch = wl.torch
    dev = wl.sets[0]["theta"].device
    host = []
    for s in wl.sets[:2]:
        h = {k: v.cpu().pin_memory() for k, v in s.items() if isinstance(v, torch.Tensor)}
        if u8_masks and "gt" in h:
            h["gt"] = h["gt"].to(torch.uint8).pin_memory()
        host.append(h)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    d2h = 0
    steps = max(1, min(steps, 50))

    def one(i):
        nonlocal d2h
        h = host[i % len(host)]
        d = {k: v.to(dev, non_blocking=True) for k, v in h.items()}
        if wl.kind == "train":
            r = wl.stage.train_step(d["theta"], d["gt"], d["w"], "MSE", d["gt_poi"], d["nz"], d["num"], 1.0, 8.0, True)
            outs = [r["loss"], r["dtheta"]]
        elif wl.kind == "predict":
            r = wl.stage.predict_tail(d["theta"], d["logits"], True, True)
            outs = [r["consist_score"], r["poi"]]
        elif wl.kind == "consist":
            import sfh_b200
            r = sfh_b200.consistency_step(d["logits"], d["wm"], 4, 1.0, True)
            outs = [r["loss"]]
        else:
            r = wl.stage.warp(d["theta"])
            outs = [r.sum()]
        got = [o.cpu() for o in outs]           # device -> host read of the step's result (syncs)
        d2h = sum(o.numel() * o.element_size() for o in got)

    with torch.no_grad():
        for i in range(min(max(warmup, 1), 3)):
            one(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            one(i)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
    return dt, h2d, d2h


def parity_check(wl, frames=3):
    """One slice of buffer set 0 through the TIMED stage object (production mode: edge-free shortcut, device
    meshgrid factors, in-launch reductions) against the C oracle fed the same meshgrid factors.
    Tolerances: BASELINE.json north_star (1e-5 abs float masks, bit-exact int masks, 1e-4 px POI, 1e-4 rel dtheta)."""
    import numpy as np
    from oracle import c_oracle as co
    torch = wl.torch
    s = wl.sets[0]
    n = min(frames, wl.B)
    th = s["theta"][:n].contiguous()
    xs, ys = (t.cpu().numpy() for t in wl.stage.warper.grid_factors(th.device))
    tm = wl.tmpl_cpu.numpy()
    out = {"frames": n, "against": "oracle/warp_oracle.c (plain-C restatement, same meshgrid factors)"}
    ok = True
    with torch.no_grad():
        if wl.kind == "train":
            r = wl.stage.train_step(th, s["gt"][:n].contiguous(), s["w"][:n].contiguous(), "MSE", s["gt_poi"][:n].contiguous(),
                                    s["nz"][:n].contiguous(), s["num"][:n].contiguous(), 1.0, 8.0, True)
            warp_ref, Lb_ref, J_ref = co.warp_loss(th.cpu().numpy(), tm, s["gt"][:n].cpu().numpy(), 4, "MSE", xs, ys)
            out["warp_mask_max_abs"] = float(np.abs(r["warp_mask"].cpu().numpy() - warp_ref).max())
            out["rec_loss_max_rel"] = float(np.abs(r["rec_per_sample"].cpu().numpy() / Lb_ref - 1).max())
            p64 = co.poi_fwd(th.cpu().numpy(), wl.poi_cpu.expand(n, -1, -1).numpy())
            out["poi_max_px"] = float(np.abs(r["poi"].cpu().numpy() - p64).max() * wl.W)
            # dtheta of the rec term alone (the oracle's J): a second call without the reprojection term
            r2 = wl.stage.train_step(th, s["gt"][:n].contiguous(), s["w"][:n].contiguous(), "MSE", None, None, None, 1.0, 0.0, False)
            g = r2["dtheta"].cpu().numpy().reshape(n, -1) * n
            jr = J_ref.reshape(n, -1)
            out["dtheta_max_rel"] = float((np.linalg.norm(g - jr, axis=1) / np.linalg.norm(jr, axis=1)).max())
            ok = out["warp_mask_max_abs"] <= 1e-5 and out["rec_loss_max_rel"] <= 1e-5 and \
                out["poi_max_px"] <= 1e-4 and out["dtheta_max_rel"] <= 1e-4
        elif wl.kind == "predict":
            lg = s["logits"][:n].contiguous()
            r = wl.stage.predict_tail(th, lg, True, True)
            m_ref, s_ref = co.predict_tail(th.cpu().numpy(), tm, lg.cpu().numpy(), 4, wl.H, wl.W, "nearest", xs, ys)
            out["mask_mismatches"] = int((r["warp_mask"].cpu().numpy() != m_ref).sum())
            out["score_max_rel"] = float(np.abs(r["consist_score"].cpu().numpy() / s_ref - 1).max())
            ok = out["mask_mismatches"] == 0 and out["score_max_rel"] <= 1e-5
        elif wl.kind == "fwd":
            r = wl.step(0)["warp_mask"][:n]
            ref = co.warp_fwd(th.cpu().numpy(), tm, wl.H, wl.W, "bilinear", xs, ys)[:, 0]
            out["warp_mask_max_abs"] = float(np.abs(r.reshape(n, wl.H, wl.W).cpu().numpy() - ref).max())
            ok = out["warp_mask_max_abs"] <= 1e-5
        else:
            return None
    out["ok"] = bool(ok)
    return out


def roofline_of(wl, ms, kern_us, peak, peak_src, traffic):
    step_bytes = wl.B * wl.bytes_per_frame
    # duration of the step's launch: the timed region itself (K back-to-back steps between one CUDA event pair on
    # the launching stream => average duration per step, kernel + gaps, an upper bound on the kernel's own time).
    # The isolated per-launch event pairs (2 us timer granularity, eager launch gaps inside the pair) are reported too.
    k = ms * 1e3 if kern_us is None else min(ms * 1e3, kern_us)
    achieved = step_bytes / (k * 1e-6) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "kernel": "sfh::k_fused",
            "kernel_us": k, "kernel_us_isolated_event_pairs": kern_us,
            "duration_source": "min(timed region / steps, isolated per-launch event pairs): CUDA events on the launching stream",
            "algorithmic_bytes_per_launch": step_bytes, "bytes_per_frame": wl.bytes_per_frame}


def size_matched_stream(wl, kern_us):
    """What a math-free kernel moving the SAME bytes (int64 in, fp32 out) achieves at this size."""
    import sfh_b200
    torch = wl.torch
    dev = wl.sets[0]["theta"].device
    srcs = [s_["gt"] for s_ in wl.sets]
    dsts = [torch.empty(s_["gt"].shape, dtype=torch.float32, device=dev) for s_ in wl.sets]
    cst = torch.cuda.current_stream().cuda_stream
    nel = srcs[0].numel()

    def timed(fn):
        evs = []
        for i in range(40):
            torch.cuda._sleep(400000)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(i % wl.nsets); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return statistics.median(a.elapsed_time(b) * 1e3 for a, b in evs[5:])
    us_torch = timed(lambda j: dsts[j].copy_(srcs[j]))
    us = timed(lambda j: sfh_b200._lib.lib().sfh_debug_stream_cast(srcs[j].data_ptr(), dsts[j].data_ptr(), nel, 148 * 8, cst))
    step_bytes = wl.B * wl.bytes_per_frame
    return {"what": "math-free int64->fp32 stream of one step's gt (same algorithmic bytes): this library's "
                    "grid-stride 128-bit kernel (sfh_debug_stream_cast, 1184 CTAs) and the stock torch cast",
            "us": us, "GBps": step_bytes / (us * 1e-6) / 1e9, "torch_cast_us": us_torch,
            "kernel_time_vs_this": kern_us / us}


def traffic_of(name):
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(name)
        if tj:
            return tj["dram_read_bytes"] + tj["dram_write_bytes"]
    return None


def run_c5(dev, world, rank, use_graph, dist_ring):
    """C5 as BASELINE.json states it: 65,536 frames of C3 work, batch-sharded, total time."""
    import torch
    import torch.distributed as dist
    wl = Workload("c5", dev, seed=rank, nsets=2)
    steps = C5_FRAMES // (wl.B * world)
    ms, _, mode, _, spg = time_workload(wl, steps, 4, use_graph, dist_ring, isolated=False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    par = parity_check(wl, frames=2) if rank == 0 else None
    total_ms = ms * steps
    out = {"frames_total": steps * wl.B * world, "n_gpus": world, "micro_batch": wl.B, "steps_per_rank": steps,
           "total_ms": total_ms, "frames_per_s": steps * wl.B * world / (total_ms * 1e-3), "ms_per_step": ms,
           "frac_of_hbm_peak_per_gpu": None, "launch": mode, "workload": wl.desc, "parity": par}
    del wl
    torch.cuda.empty_cache()
    return out


def pin_to_gpu_numa_node(index):
    """Restrict this rank to the CPU cores NVML reports as local to its GPU, so that the pinned host buffers of the
    e2e leg are first-touched on the GPU's own NUMA node (8 ranks pulling 50 GB/s each through one socket's memory
    controllers is what flattened the round-1 e2e scaling).  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(local & allowed)
        if use and len(use) < len(allowed):
            os.sched_setaffinity(0, use)
            return f"rank pinned to {len(use)} GPU-local cores ({use[0]}-{use[-1]})"
        return f"GPU-local cores = all {len(allowed)} visible cores (single NUMA node visible)"
    except Exception as e:
        return f"not pinned ({e!r})"


def main_ours(args):
    import torch
    import torch.distributed as dist
    import sfh_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — sfh_b200 has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:          # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local)         # before any pinned host buffer is allocated (first touch decides its node)
    dist_ring = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        dist_ring = torch.zeros(32, 2, dtype=torch.float32, device=dev)      # one slot per step of a graph
        dist.all_reduce(dist_ring)

    sfh_b200._lib.lib()                        # fail loudly if the CUDA library is missing
    use_graph = not args.no_graph
    wl = Workload(args.workload, dev, seed=rank)
    sampler = ClockSampler(local)
    sampler.start()
    ms, kern_us, mode, (t0, t1), spg = time_workload(wl, args.steps, args.warmup, use_graph, dist_ring)
    clocks = sampler.finish(t0, t1)
    t = torch.tensor([ms, kern_us], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)       # max over ranks, device-timed
    ms, kern_us = float(t[0]), float(t[1])
    e2e_dt, h2d, d2h = time_e2e(wl, args.steps, args.warmup)
    e2e8_dt, h2d8, _ = time_e2e(wl, args.steps, args.warmup, u8_masks=True) if wl.kind == "train" else (e2e_dt, h2d, d2h)
    t = torch.tensor([e2e_dt, e2e8_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_dt, e2e8_dt = float(t[0]), float(t[1])
    c5 = None
    if not args.no_extra and args.workload != "c5":
        del wl.sets[2:]                        # free HBM for the 256-frame micro-batches
        c5 = run_c5(dev, world, rank, use_graph, dist_ring)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent on this box"
        frames = wl.B * world
        pcie = 55.0e9

        def e2e_obj(w_, dt, dt8, hb, hb8, db, nrank):
            o = {"value": w_.B * nrank / dt, "unit": "frames/s", "h2d_bytes_per_step": hb, "d2h_bytes_per_step": db,
                 "pcie_frac": hb / dt / pcie,
                 "note": "public API (STNWarpStage.train_step / predict_tail) with pinned host inputs in the reference's "
                         "dtypes, H2D + launch + D2H per step; pcie_frac = h2d bytes / step time / 55 GB/s (PCIe Gen5 x16 "
                         "practical): the int64 gt masks make this leg PCIe-bound, not kernel-bound"}
            if dt8 is not None:
                o["with_uint8_masks"] = {"value": w_.B * nrank / dt8, "h2d_bytes_per_step": hb8,
                                         "note": "same call with uint8 gt masks at the surface (SURVEY §8 f-1, opt-in)"}
            return o
        if c5 is not None:
            c5["frac_of_hbm_peak_per_gpu"] = (c5["micro_batch"] * WORKLOADS["c5"][6]) / (c5["ms_per_step"] * 1e-3) / 1e9 / peak
        line = {
            "metric": METRIC, "value": frames / (ms * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args.workload),
            "launch": mode,
            "parallelism": (f"batch-sharded x{world}: no data-path collective; one NCCL all-reduce of the step's loss numerators "
                            f"PER STEP on NCCL's side stream, joined only at the end of each {spg}-step graph") if world > 1 else "single GPU",
            "clocks": clocks,
            "host_affinity": numa,
            "e2e": e2e_obj(wl, e2e_dt, e2e8_dt if wl.kind == "train" else None, h2d, h2d8, d2h, world),
            "gpu_launches": args.steps * wl.launches_per_step(),
            "roofline": roofline_of(wl, ms, kern_us, peak, peak_src, traffic_of(args.workload)),
        }
        if c5 is not None:
            line["c5"] = c5
        if world == 1 and not args.no_extra:
            line["parity"] = parity_check(wl)
            line["parity_ok"] = bool(line["parity"] and line["parity"]["ok"])
            if wl.kind == "train":
                line["roofline"]["size_matched_stream"] = size_matched_stream(wl, line["roofline"]["kernel_us"])
            line["cpu_baseline"] = {k: v for k, v in run_cpu(args.workload, 3, 1, sample_frames=8).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
            del wl
            torch.cuda.empty_cache()
            line["stock_torch_b200"] = run_stock_torch_gpu(args.workload, dev)
            peers = {}
            for name in PEERS:
                if WORKLOADS[name] is WORKLOADS[args.workload]:
                    continue
                try:
                    w2 = Workload(name, dev, seed=7)
                    m2, k2, md2, _, _ = time_workload(w2, max(args.steps, 200), max(args.warmup, 10), use_graph)
                    d2, hb, db = time_e2e(w2, 20, 3)
                    d28, hb8, _ = time_e2e(w2, 20, 3, u8_masks=True) if w2.kind == "train" else (None, None, None)
                    pe = {"metric": "warped frames/s (fwd+bwd)", "value": w2.B / (m2 * 1e-3), "unit": "frames/s", "ms_per_step": m2,
                          "config": config_of(name), "launch": md2,
                          "roofline": roofline_of(w2, m2, k2, peak, peak_src, traffic_of(name)),
                          "e2e": e2e_obj(w2, d2, d28, hb, hb8, db, 1),
                          "parity": parity_check(w2)}
                    if w2.kind == "train":
                        pe["roofline"]["size_matched_stream"] = size_matched_stream(w2, pe["roofline"]["kernel_us"])
                    pe["cpu_baseline"] = {k: v for k, v in run_cpu(name, 3, 1).items()
                                          if k in ("value", "unit", "cores", "kind", "sample")}
                    del w2
                    torch.cuda.empty_cache()
                    pe["stock_torch_b200"] = run_stock_torch_gpu(name, dev)
                    pe["e2e_vs_cpu_baseline"] = pe["e2e"]["value"] / pe["cpu_baseline"]["value"]
                    peers[name] = pe
                except Exception as e:
                    peers[name] = {"error": repr(e)}
            line["peers"] = peers
            extra = {}
            for name in OTHERS:
                if WORKLOADS[name] is WORKLOADS[args.workload]:
                    continue
                try:
                    w2 = Workload(name, dev, seed=7)
                    m2, k2, md2, _, _ = time_workload(w2, 200, 10, use_graph)
                    sb = w2.B * w2.bytes_per_frame
                    k2i, k2 = k2, min(k2, m2 * 1e3)
                    extra[name] = {"frames_per_s": w2.B / (m2 * 1e-3), "ms_per_step": m2, "kernel_us": k2,
                                   "kernel_us_isolated_event_pairs": k2i,
                                   "achieved_GBps": sb / (k2 * 1e-6) / 1e9, "frac_of_hbm_peak": sb / (k2 * 1e-6) / 1e9 / peak,
                                   "whole_job_frac_of_hbm_peak": sb / (m2 * 1e-3) / 1e9 / peak,
                                   "launch": md2, "workload": w2.desc, "parity": parity_check(w2)}
                    del w2
                    torch.cuda.empty_cache()
                except Exception as e:
                    extra[name] = {"error": repr(e)}
            line["other_workloads"] = extra
            line["parity_ok"] = bool(line["parity_ok"] and all(
                (v.get("parity") or {"ok": True})["ok"] for v in list(peers.values()) + list(extra.values()) if "error" not in v)
                and (c5 is None or (c5.get("parity") or {"ok": True})["ok"]))
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
