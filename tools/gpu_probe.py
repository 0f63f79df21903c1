#!/usr/bin/env python
"""GPU-box probe (development tool, not part of the product or the tests).

1. Which fp32 operation order does stock ATen use ON THE B200 for the restated kornia path
   (meshgrid division, cuBLAS bmm K=3 accumulation, grid_sample unnormalise)?  The sfh kernels
   replay that order; this prints which candidate matches bit for bit.
2. Kernel-vs-oracle mismatch statistics at the BASELINE sizes.
3. First timings (CUDA events) of the fused launches.
Writes gpurun_out/probe.json.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sfh_b200  # noqa: E402
from sfh_b200 import synth  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import kornia_restated as kr  # noqa: E402

dev = torch.device("cuda:0")
out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
f32 = np.float32


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


# ----------------------------------------------------------------------------- 1. op order
W, H, B = 640, 360, 4
res = {}
xs_g = kr.create_meshgrid(H, W, device=dev)[0, 0, :, 0].cpu().numpy()
ys_g = kr.create_meshgrid(H, W, device=dev)[0, :, 0, 1].cpu().numpy()
i = np.arange(W, dtype=f32)
res["mesh_div"] = bool(np.array_equal(xs_g, ((i / f32(W - 1)) - f32(.5)) * f32(2)))
res["mesh_mul_recip"] = bool(np.array_equal(xs_g, ((i * (f32(1) / f32(W - 1))) - f32(.5)) * f32(2)))
res["mesh_equals_cpu"] = bool(np.array_equal(xs_g, kr.create_meshgrid(H, W)[0, 0, :, 0].numpy()))

for fam, th in (("a", synth.theta_family_a(B, 3)), ("b", synth.theta_family_b(B, 3))):
    tmpl = torch.zeros(B, 1, 8, 8, device=dev)
    grid = kr.create_meshgrid(H, W, device=dev)
    # raw bmm output on the GPU
    pts = grid.expand(B, -1, -1, -1).reshape(-1, W, 2)
    T = th.to(dev).reshape(-1, 3, 3).repeat_interleave(H, dim=0)
    ph = torch.nn.functional.pad(pts, [0, 1], value=1.0)
    raw = torch.bmm(ph, T.permute(0, 2, 1)).reshape(B, H, W, 3).cpu().numpy()
    u = xs_g[None, None, :]
    v = ys_g[None, :, None]
    t = th.numpy()[:, 0]
    cands = {}
    for j in range(3):
        h0, h1, h2 = (t[:, j, k][:, None, None] for k in range(3))
        one = np.ones((1, 1, 1), f32)
        c = {
            "fma(v,h1,u*h0)+h2": fma(v, h1, u * h0) + h2,
            "((u*h0)+(v*h1))+h2": ((u * h0) + (v * h1)) + h2,
            "fma(u,h0,fma(v,h1,h2))": fma(u, h0, fma(v, h1, np.broadcast_to(h2, (B, H, 1)))),
            "fma(v,h1,fma(u,h0,h2))": fma(v, h1, fma(u, np.broadcast_to(h0, (B, 1, 1)), np.broadcast_to(h2, (B, 1, 1)))),
            "fma(u,h0,v*h1)+h2": fma(u, h0, v * h1) + h2,
            "(u*h0)+fma(v,h1,h2)": (u * h0) + fma(v, h1, np.broadcast_to(h2, (B, H, 1))),
        }
        for k, val in c.items():
            ok = np.array_equal(np.broadcast_to(val, (B, H, W)), raw[..., j])
            cands[k] = cands.get(k, True) and bool(ok)
    res[f"bmm_{fam}"] = cands
    flow_g = kr.HomographyWarper(H, W).flow(tmpl, th.to(dev)).cpu().numpy()
    flow_c = kr.HomographyWarper(H, W).flow(tmpl.cpu(), th).numpy()
    res[f"flow_gpu_eq_cpu_{fam}"] = float((flow_g != flow_c).mean())
    # flow from the raw GPU bmm with IEEE 1/z
    z = raw[..., 2:3]
    s = np.where(np.abs(z) > f32(1e-8), f32(1) / np.where(np.abs(z) > f32(1e-8), z, f32(1)), f32(1))
    res[f"scale_ieee_div_{fam}"] = bool(np.array_equal(s * raw[..., :2], flow_g))

# grid_sample: same flow tensor on CPU and GPU
torch.manual_seed(0)
tm = torch.rand(2, 1, 360, 640)
fl = torch.rand(2, 360, 640, 2) * 2.4 - 1.2
a = torch.nn.functional.grid_sample(tm, fl, mode="bilinear", padding_mode="zeros", align_corners=False)
b = torch.nn.functional.grid_sample(tm.to(dev), fl.to(dev), mode="bilinear", padding_mode="zeros", align_corners=False).cpu()
res["grid_sample_gpu_vs_cpu_maxdiff"] = float((a - b).abs().max())
res["grid_sample_gpu_vs_cpu_frac_diff"] = float((a != b).float().mean())
an = torch.nn.functional.grid_sample(tm, fl, mode="nearest", padding_mode="zeros", align_corners=False)
bn = torch.nn.functional.grid_sample(tm.to(dev), fl.to(dev), mode="nearest", padding_mode="zeros", align_corners=False).cpu()
res["grid_sample_nearest_gpu_vs_cpu_frac_diff"] = float((an != bn).float().mean())
out["aten_cuda_order"] = res
print(json.dumps(res, indent=1))

# ------------------------------------------------------------------- 2. kernel vs oracles
par = {}
for (W, H, B, fam) in [(640, 360, 8, "a"), (640, 360, 8, "b"), (1280, 720, 4, "a"), (1280, 720, 4, "b")]:
    tmpl, poi = sfh_b200.load_bundled("ncaa_nc4", (W, H), 4, 1)
    th = synth.theta_family_a(B, 11) if fam == "a" else synth.theta_family_b(B, 11)
    st = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4)
    o = st.warp(th.to(dev)).cpu().numpy()
    rc = co.warp_fwd(th.numpy(), tmpl.numpy(), H, W)[:, 0]
    rg = kr.warp(th.to(dev), tmpl.to(dev).expand(B, -1, -1, -1), H, W).cpu().numpy()
    rcpu = kr.warp(th, tmpl.expand(B, -1, -1, -1), H, W).numpy()
    r64 = kr.warp(th.double(), tmpl.double().expand(B, -1, -1, -1), H, W).numpy()
    d = {}
    for nm, ref in (("c_oracle", rc), ("aten_gpu", rg), ("aten_cpu", rcpu), ("fp64", r64)):
        e = np.abs(o - ref)
        d[nm] = {"max": float(e.max()), "frac_gt_1e-5": float((e > 1e-5).mean()), "frac_ne": float((e > 0).mean())}
    e = np.abs(rg - rcpu)
    d["aten_gpu_vs_aten_cpu"] = {"max": float(e.max()), "frac_gt_1e-5": float((e > 1e-5).mean())}
    e = np.abs(rg - r64)
    d["aten_gpu_vs_fp64"] = {"max": float(e.max()), "frac_gt_1e-5": float((e > 1e-5).mean())}
    par[f"{W}x{H}_{fam}"] = d
out["parity"] = par
print(json.dumps(par, indent=1))


# ------------------------------------------------------------------------------ 3. timings
def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


tim = {}
for (W, H, B, name) in [(640, 360, 64, "ncaa_nc4"), (1280, 720, 32, "ncaa_nc4")]:
    tmpl, poi = sfh_b200.load_bundled(name, (W, H), 4, 1)
    th = synth.theta_family_a(B, 5).to(dev)
    stb = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4)
    stn = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4, warp_with_nearest=True)
    gt = stn.predict_tail(th, None, False, False)["warp_mask"].to(torch.int64)
    logits = torch.randn(B, 4, 360, 640, device=dev)
    gt_poi = stb.transform_poi(th).detach()
    nz = torch.ones(B, poi.shape[1], device=dev)
    num = nz.sum(1)
    wts = torch.ones(B, dtype=torch.float64, device=dev)
    px = B * H * W
    with torch.no_grad():
        t = timeit(lambda: stb.warp(th))
        tim[f"warp_bilinear_{W}x{H}_B{B}"] = {"us": t, "GBps": px * 4 / t / 1e3}
        t = timeit(lambda: stn.warp(th))
        tim[f"warp_nearest_{W}x{H}_B{B}"] = {"us": t, "GBps": px * 4 / t / 1e3}
        t = timeit(lambda: stb.train_tail(th, gt, "MSE", gt_poi, nz, num, True, wts, 1.0, 8.0))
        tim[f"train_tail_{W}x{H}_B{B}"] = {"us": t, "GBps": px * 12 / t / 1e3}
        t = timeit(lambda: stb.train_tail(th, gt, "MSE", gt_poi, nz, num, False, wts, 1.0, 8.0))
        tim[f"train_tail_nomask_{W}x{H}_B{B}"] = {"us": t, "GBps": px * 8 / t / 1e3}
        t = timeit(lambda: stn.predict_tail(th, logits, True, True))
        tim[f"predict_tail_{W}x{H}_B{B}"] = {"us": t, "GBps": (px * 4 + B * 4 * 360 * 640 * 4) / t / 1e3}
        # the reference's de-facto GPU path: stock ATen
        tm_b = tmpl.to(dev).expand(B, -1, -1, -1)
        t = timeit(lambda: kr.warp(th, tm_b, H, W), n=5, warm=2)
        tim[f"aten_warp_bilinear_{W}x{H}_B{B}"] = {"us": t, "GBps": px * 4 / t / 1e3}
out["timing"] = tim
print(json.dumps(tim, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
