#!/usr/bin/env python
"""Generate the committed fixtures from the UNMODIFIED reference (run in the build container only).

  python tools/make_golden.py            # needs /root/reference (read-only) — never used at test time

Writes
  sports-field-homography_b200/data/court_templates.npz   class-index templates (2 bit/px) + POI
  tests/golden/golden_small.npz                           known-answer vectors, 64x36
  tests/golden/golden_real_theta.npz                      the two real thetas of utils/mapping_example.py

What is executed from /root/reference (imported by path, not copied):
  utils/dataset.py   open_court_template, open_court_poi          (:47-96)
  models/reconstructor.py  Reconstructor.warp / transform_poi / predict   (:109-130, :196-247)
  models/losses.py   reprojection_loss, per_sample_weighted_criterion      (:6-41)
  dataset_utils/preparation.py:219-221  colour -> class map (values restated below)
kornia is absent from the image, so ``oracle/kornia_stub.py`` supplies HomographyWarper /
transform_points (restated, parity UNPINNED at that boundary); everything around them is the
reference's own code.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import kornia_stub  # noqa: E402

kornia_stub.install()
sys.path.insert(0, REF)
from models.reconstructor import Reconstructor  # noqa: E402
from models import losses as ref_losses  # noqa: E402
from utils.dataset import open_court_poi, open_court_template  # noqa: E402

NCAA = f"{REF}/assets/mask_ncaa_v4_nc4_m_onehot.png"
NCAA_POI = f"{REF}/assets/template_ncaa_v4_points.json"
PITCH = f"{REF}/assets/pitch_mask_v3_nc4_hd.png"
PITCH_POI = f"{REF}/assets/template_pitch_points.json"


def pack2(cls: np.ndarray) -> np.ndarray:
    c = cls.astype(np.uint8).reshape(-1)
    bits = np.stack([(c >> 1) & 1, c & 1], axis=1).reshape(-1)
    return np.packbits(bits)


def pitch_classes(size):
    """pitch_mask_v3_nc4_hd.png (RGBA colours) -> class ids with generate_onehot's nc=4 map
    (dataset_utils/preparation.py:219-221: BGR (0,255,0)->1, (255,0,0)->2, (0,0,255)->3)."""
    import cv2
    bgr = cv2.imread(PITCH, 1)
    if (bgr.shape[1], bgr.shape[0]) != tuple(size):
        bgr = cv2.resize(bgr, tuple(size), interpolation=cv2.INTER_NEAREST)
    cls = np.zeros(bgr.shape[:2], np.uint8)
    for k, col in {1: (0, 255, 0), 2: (255, 0, 0), 3: (0, 0, 255)}.items():
        cls[np.all(bgr == np.array(col, np.uint8), axis=2)] = k
    return cls


def make_templates():
    out = {}
    for size in [(640, 360), (1280, 720)]:
        t = open_court_template(NCAA, 4, size, 1)           # reference loader
        cls = (t[0, 0].numpy() * 4).round().astype(np.uint8)
        assert np.array_equal(cls.astype(np.float64) / 4.0, t[0, 0].numpy().astype(np.float64))
        out[f"ncaa_nc4_{size[0]}x{size[1]}_bits"] = pack2(cls)
        out[f"pitch_v3_nc4_{size[0]}x{size[1]}_bits"] = pack2(pitch_classes(size))
    out["ncaa_poi"] = open_court_poi(NCAA_POI, 1)[0].numpy()
    out["pitch_poi"] = open_court_poi(PITCH_POI, 1)[0].numpy()
    path = os.path.join(ROOT, "sports-field-homography_b200", "data", "court_templates.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def bare_reconstructor(tmpl, poi, size, nearest):
    return Reconstructor(tmpl, poi, target_size=size, warp_size=size, unet_size=size, mask_classes=4,
                         use_unet=False, use_resnet=False, warp_with_nearest=nearest)


def predict_tail_via_reference(net, theta, logits, consistency=True, project_poi=True):
    """Run Reconstructor.predict's own code (models/reconstructor.py:196-247) with the trunk
    replaced by fixed logits/theta: use_unet/use_resnet stay True, forward_unet/resnet_reg are
    stubbed on the instance."""
    net.use_unet, net.use_resnet = True, True
    from models.reconstructor import Input
    net.resnet_input = Input.MASK
    net.forward_unet = lambda x: (logits, None, None)
    net.resnet_reg = lambda y: theta
    with torch.no_grad():
        return net.predict(torch.zeros(1), consistency=consistency, project_poi=project_poi)


def make_small():
    torch.manual_seed(20261018)
    W, H, B = 64, 36, 6
    tmpl = open_court_template(NCAA, 4, (W, H), B)
    poi = open_court_poi(NCAA_POI, B)
    theta = torch.eye(3)[None, None].repeat(B, 1, 1, 1) + (torch.rand(B, 1, 3, 3) - 0.5) * 0.3
    theta[1] *= 13.25                      # un-normalised scale like utils/mapping_example.py
    theta[2, 0, 2, 0] = 0.9                # strong perspective: Z crosses small values
    theta[3] = torch.eye(3)[None]          # exact identity (not an identity resample)
    theta[4, 0, 0, 2] = 1.7                # mostly out of bounds
    net_b = bare_reconstructor(tmpl, poi, (W, H), nearest=False)
    net_n = bare_reconstructor(tmpl, poi, (W, H), nearest=True)

    th = theta.clone().requires_grad_(True)
    warp_b = net_b.warp(th, tmpl)                       # reference method
    poi_out = net_b.transform_poi(th, poi)              # reference method
    gt = (net_n.warp(theta + torch.randn_like(theta) * 0.01, tmpl) * 4).to(torch.int64)
    gt_f = gt.to(torch.float32) / 4.0
    w1 = torch.rand(B, dtype=torch.float64) + 0.5       # [B] fp64 (utils/dataset.py:220)
    w2 = torch.rand(B, 1) + 0.5                         # [B,1] fp32 (utils/dataset.py:205-207)
    gt_poi = net_b.transform_poi(theta + torch.randn_like(theta) * 0.01, poi).detach()
    nonzeros = (torch.rand(B, poi.shape[1]) < 0.8).float()
    nonzeros[:, 0] = 1.0
    num_nonzero = nonzeros.sum(1)

    out = dict(theta=theta.numpy(), gt=gt.numpy(), w1=w1.numpy(), w2=w2.numpy(), gt_poi=gt_poi.numpy(),
               nonzeros=nonzeros.numpy(), num_nonzero=num_nonzero.numpy(),
               template=tmpl[0, 0].numpy(), court_poi=poi[0].numpy(),
               warp_bilinear=warp_b.detach().numpy(), warp_nearest=net_n.warp(theta, tmpl).numpy(),
               poi=poi_out.detach().numpy())
    mse = torch.nn.MSELoss(reduction="none")
    sl1 = torch.nn.SmoothL1Loss(reduction="none")
    for name, crit in (("mse", mse), ("sl1", sl1)):
        for wn, w in (("w1", w1), ("w2", w2)):
            loss = ref_losses.per_sample_weighted_criterion(crit, warp_b, gt_f, w)
            (g,) = torch.autograd.grad(loss, th, retain_graph=True)
            out[f"rec_{name}_{wn}"] = loss.detach().numpy()
            out[f"rec_{name}_{wn}_dtheta"] = g.numpy()
    out["rec_mse_per_sample"] = torch.mean(mse(warp_b, gt_f), dim=(1, 2)).detach().numpy()
    out["rec_sl1_per_sample"] = torch.mean(sl1(warp_b, gt_f), dim=(1, 2)).detach().numpy()
    out["rec_fmse_eval"] = ref_losses.per_sample_weighted_criterion(
        torch.nn.functional.mse_loss, warp_b, gt_f, w1).detach().numpy()          # eval.py:186-188
    for red in ("mean", "sum"):
        loss = ref_losses.reprojection_loss(poi_out, gt_poi, nonzeros, num_nonzero, red)
        (g,) = torch.autograd.grad(loss, th, retain_graph=True)
        out[f"reproj_{red}"] = loss.detach().numpy()
        out[f"reproj_{red}_dtheta"] = g.numpy()
    (g,) = torch.autograd.grad(ref_losses.reprojection_loss(poi_out, gt_poi, nonzeros, num_nonzero),
                               poi_out, retain_graph=True)
    out["reproj_mean_dpoi"] = g.numpy()
    # an arbitrary upstream gradient through warp and through poi (drop-in autograd)
    go = torch.randn_like(warp_b)
    (g,) = torch.autograd.grad(warp_b, th, go, retain_graph=True)
    out["warp_grad_out"], out["warp_dtheta"] = go.numpy(), g.numpy()
    gp = torch.randn_like(poi_out)
    (g,) = torch.autograd.grad(poi_out, th, gp, retain_graph=True)
    out["poi_grad_out"], out["poi_dtheta"] = gp.numpy(), g.numpy()

    # predict tail through the reference's own predict(): logits at half size and at full size
    for tag, (lh, lw) in (("half", (H // 2, W // 2)), ("full", (H, W)), ("odd", (25, 40))):
        logits = torch.randn(B, 4, lh, lw)
        for mode, net in (("nearest", net_n), ("bilinear", net_b)):
            r = predict_tail_via_reference(bare_reconstructor(tmpl, poi, (W, H), mode == "nearest"),
                                           theta, logits)
            out[f"pred_{tag}_{mode}_mask"] = r["warp_mask"].numpy()
            out[f"pred_{tag}_{mode}_score"] = r["consist_score"].numpy()
            out[f"pred_{tag}_{mode}_poi"] = r["poi"].numpy()
        out[f"pred_{tag}_logits"] = logits.numpy()
    path = os.path.join(ROOT, "tests", "golden", "golden_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def make_real_theta():
    """The two real predicted frame->court homographies of utils/mapping_example.py:12-22,48-58."""
    theta = torch.tensor([
        [[8.030766487121582, -0.22687992453575134, 9.891857147216797],
         [3.553352117538452, 25.72734260559082, -0.09768841415643692],
         [0.1463453769683838, 5.179210662841797, 16.56546974182129]],
        [[5.78266048, -0.43701401, 8.0031395],
         [3.63819695, 15.77359295, -0.46604609],
         [0.14406031, 3.68673325, 13.25017166]]], dtype=torch.float32)[:, None]
    out = {"theta": theta.numpy()}
    for (W, H) in [(640, 360), (1280, 720)]:
        tmpl = open_court_template(NCAA, 4, (W, H), 2)
        poi = open_court_poi(NCAA_POI, 2)
        net_n = bare_reconstructor(tmpl, poi, (W, H), True)
        net_b = bare_reconstructor(tmpl, poi, (W, H), False)
        m = (net_n.warp(theta, tmpl) * 4).to(torch.int32).numpy()
        out[f"nearest_{W}x{H}_bits"] = pack2(m)
        wb = net_b.warp(theta, tmpl).numpy()
        # bilinear: keep a strided sample + per-sample double sums to stay small
        out[f"bilinear_{W}x{H}_sample"] = wb[:, ::7, ::5].copy()
        out[f"bilinear_{W}x{H}_sum"] = wb.astype(np.float64).sum(axis=(1, 2))
        out[f"poi_{W}x{H}"] = net_b.transform_poi(theta, poi).numpy()
    path = os.path.join(ROOT, "tests", "golden", "golden_real_theta.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    torch.set_num_threads(1)
    make_templates()
    make_small()
    make_real_theta()
