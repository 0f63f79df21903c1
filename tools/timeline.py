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

# per-SM view: how many CTAs each SM ran, and how long a freed slot stays empty before the next CTA starts
sm = t[:, 7].astype(int)
gaps, per_sm, busy = [], [], []
span = (t[:, 6].max() - t0)
for m in np.unique(sm):
    c = t[sm == m]
    per_sm.append(len(c))
    ends = sorted(c[:, 6].tolist())
    for s_ in sorted(c[:, 0].tolist()):
        prev = [e for e in ends if e <= s_]
        if prev:
            e = max(prev)
            ends.remove(e)
            gaps.append((s_ - e) / 1e3)
    busy.append((c[:, 6] - c[:, 0]).sum() / span)
print("SMs %d  CTAs/SM min %d median %d max %d" % (len(per_sm), min(per_sm), int(np.median(per_sm)), max(per_sm)))
print("slot refill gap us: median %.2f p90 %.2f max %.2f (n=%d)" % (np.median(gaps), np.percentile(gaps, 90), max(gaps), len(gaps)))
print("average resident CTAs per SM over the kernel span: median %.2f min %.2f" % (np.median(busy), min(busy)))
ends_all = np.sort(t[:, 6] - t0) / 1e3
print("CTA end times us: 50%%=%.1f 90%%=%.1f 99%%=%.1f last=%.1f" % tuple(np.percentile(ends_all, [50, 90, 99, 100])))
