"""CPU tests (no GPU): pin the oracles.

The reference ships no tests or golden vectors, so the pins are
 (1) tests/golden/*.npz — outputs of the reference's OWN code (Reconstructor.warp /
     transform_poi / predict, models/losses.py, utils/dataset.py loaders) run in the build
     container with oracle/kornia_stub.py standing in for kornia (tools/make_golden.py);
 (2) closed-form known answers for the kornia + grid_sample semantics that stay unpinned.
Both oracles (torch restatement, plain-C op-order restatement) must reproduce them.
"""
import numpy as np
import pytest
import torch

from conftest import unpack2
from oracle import c_oracle as co
from oracle import kornia_restated as kr

torch.set_num_threads(2)


def T(a):
    return torch.from_numpy(np.asarray(a))


# ------------------------------------------------------------------------------- golden vectors
def test_torch_oracle_matches_reference_golden(golden_small):
    g = golden_small
    B = g["theta"].shape[0]
    H, W = g["template"].shape
    tmpl = T(g["template"])[None, None].repeat(B, 1, 1, 1)
    th = T(g["theta"])
    poi = T(g["court_poi"])[None].repeat(B, 1, 1)
    assert np.array_equal(kr.warp(th, tmpl, H, W, "bilinear").numpy(), g["warp_bilinear"])
    assert np.array_equal(kr.warp(th, tmpl, H, W, "nearest").numpy(), g["warp_nearest"])
    np.testing.assert_allclose(kr.transform_poi(th, poi).numpy(), g["poi"], rtol=0, atol=1e-6)
    for tag in ("half", "full", "odd"):
        for mode in ("nearest", "bilinear"):
            r = kr.predict_tail(th, tmpl, T(g[f"pred_{tag}_logits"]), poi, 4, H, W, mode)
            assert np.array_equal(r["warp_mask"].numpy(), g[f"pred_{tag}_{mode}_mask"])
            assert r["warp_mask"].dtype == torch.int32
            np.testing.assert_allclose(r["consist_score"].numpy(), g[f"pred_{tag}_{mode}_score"], rtol=1e-6)


def test_c_oracle_matches_reference_golden(golden_small):
    g = golden_small
    H, W = g["template"].shape
    tmpl = g["template"][None, None]
    assert np.array_equal(co.warp_fwd(g["theta"], tmpl, H, W, "bilinear")[:, 0], g["warp_bilinear"])
    assert np.array_equal(co.warp_fwd(g["theta"], tmpl, H, W, "nearest")[:, 0], g["warp_nearest"])
    B = g["theta"].shape[0]
    poi = co.poi_fwd(g["theta"], np.repeat(g["court_poi"][None], B, 0))
    # fp64 adjugate vs the reference's fp32 LU inverse: tolerance = the reference's own noise
    np.testing.assert_allclose(poi, g["poi"], rtol=0, atol=2e-5)
    for tag in ("half", "full", "odd"):
        for mode in ("nearest", "bilinear"):
            m, s = co.predict_tail(g["theta"], tmpl, g[f"pred_{tag}_logits"], 4, H, W, mode)
            assert np.array_equal(m, g[f"pred_{tag}_{mode}_mask"]), (tag, mode)
            np.testing.assert_allclose(s, g[f"pred_{tag}_{mode}_score"], rtol=2e-6)


def test_c_oracle_losses_and_gradients_match_reference_golden(golden_small):
    g = golden_small
    H, W = g["template"].shape
    B = g["theta"].shape[0]
    tmpl = g["template"][None, None]
    for kind, name in (("MSE", "mse"), ("SmoothL1", "sl1")):
        warp, Lb, J = co.warp_loss(g["theta"], tmpl, g["gt"], 4, kind)
        assert np.array_equal(warp, g["warp_bilinear"])
        np.testing.assert_allclose(Lb, g[f"rec_{name}_per_sample"], rtol=2e-6)
        # models/losses.py:38-39 with w [B] (elementwise) and w [B,1] (the [B,B] broadcast quirk)
        w1, w2 = g["w1"], g["w2"]
        np.testing.assert_allclose(np.mean(Lb * w1), g[f"rec_{name}_w1"], rtol=2e-6)
        np.testing.assert_allclose(np.mean(Lb[None, :] * w2), g[f"rec_{name}_w2"], rtol=2e-6)
        d1 = J * (w1 / B)[:, None, None]
        d2 = J * (w2.sum() / (B * B))
        for d, ref in ((d1, g[f"rec_{name}_w1_dtheta"]), (d2, g[f"rec_{name}_w2_dtheta"])):
            ref = ref.reshape(B, 3, 3)
            err = np.linalg.norm((d - ref).reshape(B, -1), axis=1) / (np.linalg.norm(ref.reshape(B, -1), axis=1) + 1e-30)
            assert err.max() < 1e-4, err
    dth = co.warp_bwd(g["theta"], tmpl, g["warp_grad_out"][:, None])
    ref = g["warp_dtheta"].reshape(B, 3, 3)
    err = np.linalg.norm((dth - ref).reshape(B, -1), axis=1) / np.linalg.norm(ref.reshape(B, -1), axis=1)
    assert err.max() < 1e-5, err
    Rb = co.reproj_per_sample(g["poi"], g["gt_poi"], g["nonzeros"], g["num_nonzero"])
    np.testing.assert_allclose(Rb.mean(), g["reproj_mean"], rtol=2e-6)
    np.testing.assert_allclose(Rb.sum(), g["reproj_sum"], rtol=2e-6)


def test_oracles_match_real_theta_golden(golden_real):
    """The two real predicted homographies of utils/mapping_example.py:12-22,48-58."""
    import sfh_b200
    g = golden_real
    for (W, H) in [(640, 360), (1280, 720)]:
        tmpl, poi = sfh_b200.load_bundled("ncaa_nc4", (W, H), 4, 2)
        m = co.warp_fwd(g["theta"], tmpl[:1].numpy(), H, W, "nearest")[:, 0] * 4
        assert np.array_equal(m.astype(np.int32), unpack2(g[f"nearest_{W}x{H}_bits"], (2, H, W)))
        wb = co.warp_fwd(g["theta"], tmpl[:1].numpy(), H, W, "bilinear")[:, 0]
        assert np.array_equal(wb[:, ::7, ::5], g[f"bilinear_{W}x{H}_sample"])
        np.testing.assert_allclose(wb.astype(np.float64).sum((1, 2)), g[f"bilinear_{W}x{H}_sum"], rtol=1e-12)
        p = co.poi_fwd(g["theta"], poi.numpy())
        assert np.abs(p - g[f"poi_{W}x{H}"]).max() * W < 5e-4       # px; reference fp32 LU noise
        if W == 640:
            wt = kr.warp(T(g["theta"]), tmpl, H, W, "bilinear").numpy()
            assert np.array_equal(wt, wb)


def test_bundled_templates_are_the_reference_loader_output(golden_small):
    import sfh_b200
    tmpl, poi = sfh_b200.load_bundled("ncaa_nc4", (640, 360), 4, 3)
    assert tmpl.shape == (3, 1, 360, 640) and tmpl.dtype == torch.float32
    assert set(np.unique(tmpl.numpy()).tolist()) == {0.0, 0.25, 0.5, 0.75}
    assert poi.shape == (3, 52, 2) and float(poi.abs().max()) <= 1.0
    np.testing.assert_array_equal(poi[0].numpy(), golden_small["court_poi"])
    hist = np.bincount((tmpl[0, 0].numpy() * 4).astype(int).ravel(), minlength=4) / (640 * 360)
    np.testing.assert_allclose(hist, [0.248, 0.453, 0.227, 0.072], atol=0.01)
    t2, p2 = sfh_b200.load_bundled("pitch_v3_nc4", (1280, 720), 4, 1)
    assert t2.shape == (1, 1, 720, 1280) and p2.shape == (1, 33, 2)
    hist = np.bincount((t2[0, 0].numpy() * 4).astype(int).ravel(), minlength=4) / (1280 * 720)
    np.testing.assert_allclose(hist, [0.090, 0.734, 0.152, 0.025], atol=0.01)


# ------------------------------------------------------------------- C oracle == torch oracle
# NOTE: torch CPU bmm switches to a non-MKL kernel with another rounding order when
# W*3*3 < 400 (W < 45); the reference never runs that small, so exact checks use W >= 48.
@pytest.mark.parametrize("W,H,B,scale", [(64, 36, 5, 1.0), (640, 360, 3, 1.0), (640, 360, 2, 15.0), (200, 77, 3, 1.0)])
def test_c_oracle_is_bit_identical_to_aten_cpu(W, H, B, scale):
    import sfh_b200
    from sfh_b200.synth import theta_family_a
    tmpl, _ = sfh_b200.load_bundled("ncaa_nc4", (640, 360), 4, 1)
    th = theta_family_a(B, seed=7) * scale
    ref = kr.HomographyWarper(H, W)(tmpl.expand(B, -1, -1, -1), th).numpy()
    assert np.array_equal(co.warp_fwd(th.numpy(), tmpl.numpy(), H, W), ref)
    refn = kr.HomographyWarper(H, W, mode="nearest")(tmpl.expand(B, -1, -1, -1), th).numpy()
    assert np.array_equal(co.warp_fwd(th.numpy(), tmpl.numpy(), H, W, "nearest"), refn)
    g = kr.create_meshgrid(H, W).numpy()
    assert np.array_equal(g[0, 0, :, 0], co.meshgrid(W)) and np.array_equal(g[0, :, 0, 1], co.meshgrid(H))


def test_c_oracle_multichannel_float_template_and_per_sample_templates():
    rng = np.random.default_rng(3)
    B, C, Hc, Wc, H, W = 3, 3, 23, 31, 17, 61
    tmpl = rng.random((B, C, Hc, Wc), dtype=np.float32)
    from sfh_b200.synth import theta_family_a
    th = theta_family_a(B, seed=5, amp=0.3)
    ref = kr.HomographyWarper(H, W)(T(tmpl), th).numpy()
    assert np.array_equal(co.warp_fwd(th.numpy(), tmpl, H, W), ref)
    thr = th.clone().requires_grad_(True)
    out = kr.HomographyWarper(H, W)(T(tmpl), thr)
    go = torch.randn_like(out)
    out.backward(go)
    d = co.warp_bwd(th.numpy(), tmpl, go.numpy())
    np.testing.assert_allclose(d, thr.grad.reshape(B, 3, 3).numpy(), rtol=2e-4, atol=1e-5)


# ------------------------------------------------------------------------- closed-form answers
def test_identity_theta_is_not_an_identity_resample():
    """meshgrid normalised by (W-1) but sampled with align_corners=False:
    a row [0..7] warps to ix = (8*i/7) - 0.5 -> [0, .32, .89, ...]."""
    row = torch.arange(8, dtype=torch.float32).reshape(1, 1, 1, 8)
    th = torch.eye(3)[None]
    # height 1 would divide by zero in the meshgrid; use 2 identical rows
    out = kr.HomographyWarper(2, 8)(row.repeat(1, 1, 2, 1), th)[0, 0, 0].numpy()
    i = np.arange(8, dtype=np.float64)
    ix = ((i / 7 - 0.5) * 2 + 1) * 8 / 2 - 0.5
    x0 = np.floor(ix)
    lo = np.where((x0 >= 0) & (x0 < 8), np.clip(x0, 0, 7), 0) * ((x0 >= 0) & (x0 < 8))
    hi = np.where(x0 + 1 < 8, x0 + 1, 0) * (x0 + 1 < 8)
    # rows: v = -1 -> iy = -0.5: the out-of-image row above gets weight 0.5, row 0 the other 0.5
    expect = 0.5 * (lo * (x0 + 1 - ix) + hi * (ix - x0))
    np.testing.assert_allclose(out, expect, atol=1e-5)
    np.testing.assert_allclose(out[:3], [0.0, 0.3214286, 0.8928571], atol=2e-6)
    c = co.warp_fwd(th.numpy(), row.repeat(1, 1, 2, 1).numpy(), 2, 8)[0, 0, 0]
    assert np.array_equal(c, out)


def test_pure_translation_and_out_of_bounds_are_zero():
    tmpl = torch.ones(1, 1, 10, 10)
    th = torch.eye(3)[None].clone()
    th[0, 0, 2] = 5.0                      # x' = u + 5: everything right of the template
    assert float(kr.HomographyWarper(6, 6)(tmpl, th).abs().max()) == 0.0
    assert float(np.abs(co.warp_fwd(th.numpy(), tmpl.numpy(), 6, 6)).max()) == 0.0
    th[0, 0, 2] = 0.0
    inner = kr.HomographyWarper(6, 6)(tmpl, th)[0, 0, 1:-1, 1:-1]
    np.testing.assert_allclose(inner.numpy(), 1.0, atol=1e-6)   # weights sum to one inside


def test_nearest_ties_round_half_to_even():
    """ix = k + 0.5 must pick the even neighbour (nearbyint), SURVEY §8a-1."""
    Wc = 8
    tmpl = torch.arange(Wc, dtype=torch.float32).reshape(1, 1, 1, Wc).repeat(1, 1, 2, 1) + 1
    # choose x so that ix = ((x+1)*Wc-1)/2 is exactly 2.5 and 3.5: x = (2*ix+1)/Wc - 1
    for ix, expect in ((2.5, 3.0), (3.5, 5.0), (0.5, 1.0), (1.5, 3.0)):
        x = (2 * ix + 1) / Wc - 1
        th = torch.tensor([[[0.0, 0.0, x], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]])
        v = kr.HomographyWarper(2, 2, mode="nearest")(tmpl, th)[0, 0, 0, 0].item()
        c = co.warp_fwd(th.numpy(), tmpl.numpy(), 2, 2, "nearest")[0, 0, 0, 0]
        assert v == expect and c == expect, (ix, v, c)


def test_tiny_z_uses_scale_one_and_nonfinite_maps_to_zero():
    tmpl = torch.ones(1, 1, 4, 4)
    th = torch.tensor([[[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1e-9]]])      # |z| <= eps -> scale 1
    a = kr.HomographyWarper(4, 4)(tmpl, th).numpy()
    ident = kr.HomographyWarper(4, 4)(tmpl, torch.eye(3)[None]).numpy()
    assert np.array_equal(a, ident)
    assert np.array_equal(co.warp_fwd(th.numpy(), tmpl.numpy(), 4, 4), ident)
    # non-finite coordinates: ATen's CUDA kernel maps them to -100 => 0 (GridSampler.cuh:140-147,
    # SURVEY §5) while ATen's CPU kernel lets the NaN weights through.  The reference runs on the
    # GPU, so the C oracle (and the sfh kernels) follow the CUDA rule.
    th = torch.tensor([[[float("nan"), 0, 0], [0, 1.0, 0], [0, 0, 1.0]]])
    assert bool(torch.isnan(kr.HomographyWarper(4, 4)(tmpl, th)).all())          # ATen CPU
    assert float(np.abs(co.warp_fwd(th.numpy(), tmpl.numpy(), 4, 4)).max()) == 0.0  # ATen CUDA rule
    th = torch.tensor([[[1e30, 0, 0], [0, 1e30, 0], [0, 0, 1.0]]])
    assert np.array_equal(kr.HomographyWarper(4, 4)(tmpl, th).numpy(), co.warp_fwd(th.numpy(), tmpl.numpy(), 4, 4))


def test_transform_poi_against_cv2_perspective_transform():
    """The author's own (commented-out) cross-check, eval.py:122-138, in fp64."""
    import cv2
    import sfh_b200
    from sfh_b200.synth import theta_family_b
    _, poi = sfh_b200.load_bundled("ncaa_nc4", (640, 360), 4, 4)
    th = theta_family_b(4, seed=11)
    ours = co.poi_fwd(th.numpy(), poi.numpy(), normalize=False).astype(np.float64)
    for b in range(4):
        Hinv = np.linalg.inv(th[b, 0].numpy().astype(np.float64))
        ref = cv2.perspectiveTransform(poi[b:b + 1].numpy().astype(np.float64), Hinv)[0]
        np.testing.assert_allclose(ours[b], ref, atol=2e-6)
    t32 = kr.transform_poi(th, poi, normalize=False).numpy()
    assert np.abs(t32 - ours).max() < 1e-4


def test_reference_shape_conventions():
    """theta [B,1,3,3] and [B,3,3] are both accepted; outputs keep the documented shapes/dtypes."""
    tmpl = torch.rand(3, 1, 9, 16)
    th = torch.eye(3)[None].repeat(3, 1, 1)
    a = kr.HomographyWarper(9, 16)(tmpl, th)
    b = kr.HomographyWarper(9, 16)(tmpl, th[:, None])
    assert a.shape == (3, 1, 9, 16) and torch.equal(a, b)
    pts = torch.rand(3, 5, 2)
    assert torch.equal(kr.transform_points(th, pts), kr.transform_points(th[:, None], pts))
    with pytest.raises(ValueError):
        kr.transform_points(th[:2], pts)


def test_consistency_loss_restatement_closed_form():
    """train.py:219-223: uniform logits give log(nc) whatever the mask; a confident correct prediction
    gives ~0; truncation maps values between two class levels to the lower class."""
    import math
    from oracle import kornia_restated as kr
    B, nc, H, W = 2, 4, 5, 7
    warp = torch.randint(0, nc, (B, 1, H, W)).float() / nc
    assert abs(float(kr.consistency_loss(torch.zeros(B, nc, H, W), warp, nc)) - math.log(nc)) < 1e-6
    cls = (warp * nc).long()[:, 0]
    conf = torch.nn.functional.one_hot(cls, nc).permute(0, 3, 1, 2).float() * 50.0
    assert float(kr.consistency_loss(conf, warp, nc, 3.0)) < 1e-6
    assert float(kr.consistency_loss(conf, warp + 0.2 / nc, nc)) < 1e-6          # still the same class after trunc
    assert float(kr.consistency_loss(conf, warp[:, 0], nc)) < 1e-6                # [B,H,W] accepted


def test_postprocess_restatement_matches_cv2_and_golden_tables():
    """the restated cv2.INTER_NEAREST index rule equals the tables recovered from the real
    cv2 (tests/golden/cv2_nearest_tables.npz, tools/make_golden_post.py) and, when cv2 is importable,
    cv2.resize itself; mask_type conversions follow predict.py:288-299 / utils/postprocess.py:21-58."""
    import os
    from oracle import postprocess_restated as pr
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_nearest_tables.npz"))
    n = 0
    for k in g.files:
        if k == "cv2_version":
            continue
        s, d = map(int, k.split("_"))
        assert np.array_equal(pr.nearest_table(s, d), g[k]), k
        n += 1
    assert n >= 18
    rng = np.random.default_rng(0)
    m = rng.integers(0, 4, size=(2, 36, 64)).astype(np.uint8)
    rgb = pr.onehot_to_image(m, 4)
    assert rgb.shape == (2, 36, 64, 3) and tuple(rgb[m == 1][0]) == (0, 255, 0) and tuple(rgb[m == 3][0]) == (0, 0, 255)
    assert not rgb[m == 0].any()
    logits = torch.randn(2, 4, 36, 64)
    assert np.array_equal(pr.preds_to_masks(logits, 4), logits.argmax(1).numpy().astype(np.uint8))
    b = pr.postprocess(torch.from_numpy(m.astype(np.int32)), "mask", "bin", (100, 50), 4)
    assert b.shape == (2, 50, 100) and set(np.unique(b)) <= {0, 255}
    try:
        import cv2
    except ImportError:
        return
    for (ow, oh) in [(128, 72), (100, 50), (64, 36), (31, 17), (200, 111)]:
        for arr in (m[0], rgb[0]):
            assert np.array_equal(pr.resize_nearest(arr, (ow, oh)), cv2.resize(arr, (ow, oh), interpolation=cv2.INTER_NEAREST))


def test_restated_focal_loss_closed_form():
    """kornia 0.5.x focal_loss conventions pinned on a hand-computed case: two equal logits -> p = 0.5,
    softmax + 1e-8, one-hot target + 1e-6 on BOTH classes (kornia.utils.one_hot adds its eps everywhere)."""
    import math
    from oracle import kornia_restated as kr
    lg = torch.zeros(1, 2, 1, 1, dtype=torch.float64)
    t = torch.zeros(1, 1, 1, dtype=torch.long)
    p = 0.5 + 1e-8
    term = -(1 - p) ** 2 * math.log(p)
    assert abs(float(kr.focal_loss(lg, t, 1.0, 2.0, "mean")) - ((1 + 1e-6) * term + 1e-6 * term)) < 1e-15
    # gamma = 0, alpha = 1 degenerates to (eps-perturbed) cross entropy
    lg = torch.randn(2, 4, 3, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    t = torch.randint(0, 4, (2, 3, 5), generator=torch.Generator().manual_seed(1))
    ce = torch.nn.functional.cross_entropy(lg, t)
    assert abs(float(kr.focal_loss(lg, t, 1.0, 0.0, "mean")) - float(ce)) < 1e-4
