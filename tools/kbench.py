#!/usr/bin/env python
"""Development micro-benchmark: per-launch CUDA-event time of each fused launch at the BASELINE
sizes (not the bench contract; bench.py is).  Usage: python tools/kbench.py [--only train] [--n 30]"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sfh_b200  # noqa: E402
from sfh_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--n", type=int, default=30)
ap.add_argument("--sets", type=int, default=4)
ap.add_argument("--hd", action="store_true")
ap.add_argument("--oob", action="store_true", help="push every sample fully out of bounds: all patches edge-free")
ap.add_argument("--exact", action="store_true", help="disable the edge-free patch shortcut")
a = ap.parse_args()
dev = torch.device("cuda:0")
PEAK = 6545.3


def run(tag, fn, nbytes):
    if a.only and a.only not in tag:
        return
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.n)]
    for i, (e0, e1) in enumerate(evs):
        torch.cuda._sleep(400000)      # keep the GPU busy while the CPU queues e0 / launch / e1, so the
        e0.record()                    # pair brackets only the kernel, not Python's launch latency
        fn(i)
        e1.record()
    torch.cuda.synchronize()
    us = sorted(e0.elapsed_time(e1) * 1e3 for e0, e1 in evs)
    med = statistics.median(us)
    print(f"{tag:34s} med {med:8.1f} us  min {us[0]:8.1f} us  {nbytes / med / 1e3:7.0f} GB/s  {nbytes / med / 1e3 / PEAK * 100:5.1f} % of HBM peak", flush=True)


cfgs = [(640, 360, 64)] + ([(1280, 720, 32)] if a.hd else [])
for (W, H, B) in cfgs:
    tmpl, poi = sfh_b200.load_bundled("ncaa_nc4", (W, H), 4, 1)
    stb = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4, exact=a.exact)
    stn = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4, warp_with_nearest=True)
    sets = []
    for i in range(a.sets):
        th = synth.theta_family_a(B, 50 + i).to(dev)
        if a.oob:
            th[:, 0, 0, 2] = 5.0
        gt = stn.predict_tail(synth.perturb(th.cpu(), seed=i).to(dev), None, False, False)["warp_mask"].to(torch.int64)
        gt_poi = stb.transform_poi(synth.perturb(th.cpu(), seed=9).to(dev)).detach()
        nz = torch.ones(B, poi.shape[1], device=dev)
        sets.append(dict(th=th, gt=gt, gt_poi=gt_poi, nz=nz, num=nz.sum(1), w=torch.ones(B, dtype=torch.float64, device=dev),
                         logits=torch.randn(B, 4, 360, 640, device=dev), out={}, outp={}))
    px = B * H * W
    S = lambda i: sets[i % len(sets)]
    with torch.no_grad():
        run(f"store_bilinear_{W}x{H}_B{B}", lambda i: stb.warp(S(i)["th"]), px * 4)
        run(f"store_nearest_{W}x{H}_B{B}", lambda i: stn.warp(S(i)["th"]), px * 4)
        run(f"train_{W}x{H}_B{B}", lambda i: stb.train_step(S(i)["th"], S(i)["gt"], S(i)["w"], "MSE", S(i)["gt_poi"], S(i)["nz"], S(i)["num"], 1.0, 8.0, True, S(i)["out"]), px * 12)
        for s_ in sets:
            s_["gt8"] = s_["gt"].to(torch.uint8)
        run(f"train_u8gt_{W}x{H}_B{B}", lambda i: stb.train_step(S(i)["th"], S(i)["gt8"], S(i)["w"], "MSE", S(i)["gt_poi"], S(i)["nz"], S(i)["num"], 1.0, 8.0, True, S(i)["out"]), px * 5)
        run(f"train_nomask_{W}x{H}_B{B}", lambda i: stb.train_step(S(i)["th"], S(i)["gt"], S(i)["w"], "MSE", S(i)["gt_poi"], S(i)["nz"], S(i)["num"], 1.0, 8.0, False, S(i)["out"]), px * 8)
        run(f"predict_{W}x{H}_B{B}", lambda i: stn.predict_tail(S(i)["th"], S(i)["logits"], True, True, S(i)["outp"]), px * 4 + B * 4 * 360 * 640 * 4)
        run(f"predict_u8mask_{W}x{H}_B{B}", lambda i: stn.predict_tail(S(i)["th"], S(i)["logits"], True, True, S(i).setdefault("outp8", {}), torch.uint8), px * 1 + B * 4 * 360 * 640 * 4)
        run(f"predict_noscore_{W}x{H}_B{B}", lambda i: stn.predict_tail(S(i)["th"], None, False, False, S(i)["outp"]), px * 4)
        if (W, H) == (640, 360):      # training configuration: logits at the warp size
            for s_ in sets:
                s_["wm"] = stb.warp(s_["th"])
                s_["cons"] = {}
            run(f"consist_ce+dlogits_{W}x{H}_B{B}", lambda i: sfh_b200.consistency_step(S(i)["logits"], S(i)["wm"], 4, 1.0, True, S(i)["cons"]), px * 4 + 2 * B * 4 * H * W * 4)
            run(f"consist_ce_fwd_only_{W}x{H}_B{B}", lambda i: sfh_b200.consistency_step(S(i)["logits"], S(i)["wm"], 4, 1.0, False, S(i)["cons"]), px * 4 + B * 4 * H * W * 4)
        for s_ in sets:
            s_["mi"] = stn.predict_tail(s_["th"], None, False, False)["warp_mask"]
            s_["post"] = {}
        def post(i, src, mt, osz):
            S(i)["post"][(mt, osz)] = sfh_b200.postprocess_masks(S(i)[src], mt, osz, 4, S(i)["post"].get((mt, osz)))
        run(f"post_argmax_gray_{W}x{H}_B{B}", lambda i: post(i, "logits", "gray", None), B * 360 * 640 * 17)
        run(f"post_argmax_rgb_x2_{W}x{H}_B{B}", lambda i: post(i, "logits", "rgb", (1280, 720)), B * 360 * 640 * 16 + B * 1280 * 720 * 3)
        run(f"post_mask_i32_rgb_{W}x{H}_B{B}", lambda i: post(i, "mi", "rgb", None), px * 7)
        go = torch.randn(B, 1, H, W, device=dev)
        thg = S(0)["th"]
        def bwd(i):
            t = S(i)["th"].clone().requires_grad_(True)
            with torch.enable_grad():
                o = stb.warper(stb.court_img, t)
            o.backward(go)
        run(f"fwd+generic_bwd_{W}x{H}_B{B}", bwd, px * 12)
