#!/usr/bin/env python
"""Development aid: per-CTA phase timeline of k_fused (needs a -DSFH_TIMELINE build of the library,
loaded through SFH_LIB_PATH).  Prints the median duration of each phase and the concurrency."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200
from sfh_b200 import synth

dev = torch.device("cuda:0")
lib = ctypes.CDLL(sfh_b200._lib.LIB_PATH)
W, H, B = 640, 360, 64
tmpl, poi = sfh_b200.load_bundled("ncaa_nc4", (W, H), 4, 1)
stb = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4)
stn = sfh_b200.STNWarpStage(tmpl.to(dev), poi.to(dev), (W, H), 4, warp_with_nearest=True)
th = synth.theta_family_a(B, 50).to(dev)
if len(sys.argv) > 1 and sys.argv[1] == "oob":
    th[:, 0, 0, 2] = 5.0
gt = stn.predict_tail(synth.perturb(th.cpu()).to(dev), None, False, False)["warp_mask"].to(torch.int64)
w = torch.ones(B, dtype=torch.float64, device=dev)
out = {}
for _ in range(3):
    stb.train_step(th, gt, w, "MSE", out=out)
torch.cuda.synchronize()
n = 4096
buf = torch.zeros(n * 8, dtype=torch.int64, device=dev)
lib.sfh_debug_set_timeline.argtypes = [ctypes.c_void_p]
assert lib.sfh_debug_set_timeline(buf.data_ptr()) == 0
stb.train_step(th, gt, w, "MSE", out=out)
torch.cuda.synchronize()
lib.sfh_debug_set_timeline(None)
t = buf.cpu().numpy().reshape(n, 8)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
print("ctas", len(t), "kernel span us", (t[:, 6].max() - t0) / 1e3)
names = ["start->tables/sync", "->classified+list", "mbar wait (gt tile)", "patch loop", "reduce+partials", "ticket(membar)"]
for i, nm in enumerate(names):
    a, b = (i, i + 1) if i < 2 else (i + 1 if i >= 2 else i, i + 2 if i >= 2 else i + 1)
for (a, b, nm) in [(0, 1, "start -> first sync (tables, theta, corner grid)"), (1, 2, "classification + list (2 syncs)"),
                   (2, 3, "wait for the TMA gt tile"), (3, 4, "patch loop"), (4, 5, "warp/CTA reduction"), (5, 6, "ticket (membar)")]:
    d = (t[:, b] - t[:, a]) / 1e3
    print(f"{nm:52s} median {np.median(d):6.2f} us  p90 {np.percentile(d, 90):6.2f}  max {d.max():6.2f}")
d = (t[:, 6] - t[:, 0]) / 1e3
print(f"{'CTA lifetime':52s} median {np.median(d):6.2f} us  p90 {np.percentile(d, 90):6.2f}  max {d.max():6.2f}")
starts = np.sort(t[:, 0] - t0) / 1e3
print("CTA start times us: 10%%=%.1f 50%%=%.1f 90%%=%.1f last=%.1f" % tuple(np.percentile(starts, [10, 50, 90, 100])))
