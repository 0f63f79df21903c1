#!/usr/bin/env python
"""Development aid: per-tile phase timeline of k_train_stream (needs a -DSFH_TIMELINE build loaded
through SFH_LIB_PATH): stamps of the first 8 tiles of every CTA."""
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
sets = []
for k in range(3):
    th = synth.theta_family_a(B, 50 + k).to(dev)
    if len(sys.argv) > 1 and sys.argv[1] == "oob":
        th[:, 0, 0, 2] = 5.0
    gt = stn.predict_tail(synth.perturb(th.cpu(), seed=k).to(dev), None, False, False)["warp_mask"].to(torch.int64)
    sets.append((th, gt, {}))
w = torch.ones(B, dtype=torch.float64, device=dev)
for k in range(6):
    th, gt, out = sets[k % 3]
    stb.train_step(th, gt, w, "MSE", out=out)
torch.cuda.synchronize()
n = 4096
buf = torch.zeros(n * 8, dtype=torch.int64, device=dev)
lib.sfh_debug_set_timeline.argtypes = [ctypes.c_void_p]
assert lib.sfh_debug_set_timeline(buf.data_ptr()) == 0
th, gt, out = sets[0]
stb.train_step(th, gt, w, "MSE", out=out)
torch.cuda.synchronize()
lib.sfh_debug_set_timeline(None)
t = buf.cpu().numpy().reshape(n, 8).astype(np.float64)
t = t[:444 * 8].reshape(444, 8, 8)            # [cta][tile][phase]
t0 = t[t > 0].min()
print("first stamp -> last stamp us", (t.max() - t0) / 1e3)
for i in range(8):
    x = t[:, i, :]
    ok = x[:, 0] > 0
    x = x[ok]
    d = lambda a, b: np.median(x[:, b] - x[:, a]) / 1e3
    print(f"tile {i}: start at {np.median(x[:, 0] - t0) / 1e3:6.2f} us | warp1: list wait {d(0, 1):5.2f}  full wait {d(1, 2):5.2f}  patches {d(2, 3):5.2f}  flush {d(3, 4):5.2f}"
          f" | warp0: start->full {d(5, 6):5.2f}  rest {d(6, 7):5.2f}   (n={ok.sum()})")
