#!/usr/bin/env python
"""Development aid: one small, ragged invocation of every entry point (for compute-sanitizer runs:
`compute-sanitizer --tool memcheck python tools/sanity_small.py`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200  # noqa: E402
from sfh_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
for (W, H, B) in [(640, 360, 3), (200, 77, 5), (130, 50, 2), (1280, 720, 2)]:
    tmpl, poi = sfh_b200.load_bundled("ncaa_nc4", (1280, 720) if W > 640 else (640, 360), 4, 1)
    for exact in (True, False):
        st = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4, exact=exact)
        stn = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4, warp_with_nearest=True, exact=exact)
        th = synth.theta_family_b(B, 3).to(dev)
        gt = stn.predict_tail(synth.perturb(th.cpu(), seed=1).to(dev), None, False, False)["warp_mask"].to(torch.int64)
        gp = st.transform_poi(synth.perturb(th.cpu(), seed=2).to(dev)).detach()
        nz = torch.ones(B, poi.shape[1], device=dev)
        w = torch.rand(B, 1).to(dev)
        for kind in ("MSE", "SmoothL1"):
            r = st.train_step(th, gt, w, kind, gp, nz, nz.sum(1), 1.0, 8.0, True, {})
            r8 = st.train_step(th, gt.to(torch.uint8), w, kind, gp, nz, nz.sum(1), 1.0, 8.0, False, {})
        tg = th.clone().requires_grad_(True)
        st.warper(st.court_img, tg).sum().backward()
        for lg_shape in ((B, 4, H, W), (B, 4, (H + 1) // 2, (W + 1) // 2), (B, 4, 33, 47)):
            lg = torch.randn(*lg_shape, device=dev)
            stn.predict_tail(th, lg, True, True, {})
            stn.predict_tail(th, lg, True, True, {}, torch.uint8)
        wm = st.warp(th)
        lg = torch.randn(B, 4, H, W, device=dev)
        sfh_b200.consistency_step(lg, wm, 4, 0.5)
        sfh_b200.consistency_step(lg[:, :3].contiguous(), wm, 3, 0.5)
        for mt in ("gray", "bin", "rgb"):
            sfh_b200.postprocess_masks(lg, mt, (W + 37, H + 11), 4)
            sfh_b200.postprocess_masks(lg, mt, None, 4)
            sfh_b200.postprocess_masks(r["warp_mask"].mul(4).to(torch.int32)[:, 0] if r["warp_mask"].ndim == 4 else r["warp_mask"].mul(4).to(torch.int32), mt, (W // 2, H // 2), 4)
        stn.render_masks(th, "rgb", (W * 2, H * 2))
    torch.cuda.synchronize()
    print("ok", W, H, B, flush=True)
print("done")
