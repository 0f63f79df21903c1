#!/usr/bin/env python
"""Development aid: per-frame time of the C2 training tail as a function of the batch size (how much the
partial last wave of CTAs costs): back-to-back launches between one event pair."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200  # noqa: E402
from sfh_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
W, H = 640, 360
tmpl, poi = sfh_b200.load_bundled("ncaa_nc4", (W, H), 4, 1)
stb = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4)
stn = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4, warp_with_nearest=True)
for B in (44, 59, 64, 74, 89, 128):
    sets = []
    for i in range(4):
        th = synth.theta_family_a(B, 50 + i).to(dev)
        gt = stn.predict_tail(synth.perturb(th.cpu(), seed=i).to(dev), None, False, False)["warp_mask"].to(torch.int64)
        gp = stb.transform_poi(synth.perturb(th.cpu(), seed=9).to(dev)).detach()
        nz = torch.ones(B, poi.shape[1], device=dev)
        sets.append((th, gt, gp, nz, nz.sum(1), torch.ones(B, dtype=torch.float64, device=dev), {}))
    def step(i):
        th, gt, gp, nz, num, w, out = sets[i % 4]
        stb.train_step(th, gt, w, "MSE", gp, nz, num, 1.0, 8.0, True, out)
    with torch.no_grad():
        for i in range(8):
            step(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for i in range(4):
                    step(i)
        torch.cuda.current_stream().wait_stream(s)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(250):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 1000
    ctas = 5 * 6 * B
    print(f"B={B:4d} CTAs={ctas:5d} waves={ctas / 444:5.2f}  {us:7.2f} us/step  {us / B * 1e3:7.1f} ns/frame")
