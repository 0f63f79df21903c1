#!/usr/bin/env python
"""Development check: print checksums of one C2 train step so that builds / env variants
(SFH_PERSISTENT, SFH_LIB_PATH) can be compared across processes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sfh_b200  # noqa: E402
from sfh_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
W, H, B = 640, 360, 64
tmpl, poi = sfh_b200.load_bundled("ncaa_nc4", (W, H), 4, 1)
stb = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4)
stn = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4, warp_with_nearest=True)
th = synth.theta_family_a(B, 50).to(dev)
gt = stn.predict_tail(synth.perturb(th.cpu(), seed=0).to(dev), None, False, False)["warp_mask"].to(torch.int64)
gt_poi = stb.transform_poi(synth.perturb(th.cpu(), seed=9).to(dev)).detach()
nz = torch.ones(B, poi.shape[1], device=dev)
w = torch.ones(B, dtype=torch.float64, device=dev)
with torch.no_grad():
    for _ in range(2):
        r = stb.train_step(th, gt, w, "MSE", gt_poi, nz, nz.sum(1), 1.0, 8.0, True, {})
torch.cuda.synchronize()
print("loss %.17g" % float(r["loss"]), "dtheta %.17g" % float(r["dtheta"].double().abs().sum()),
      "mask %.17g" % float(r["warp_mask"].double().sum()), "rec %.17g" % float(r["rec_per_sample"].double().sum()))
