"""CPU tests of the host side: the C-ABI library builds/loads and exports exactly what
include/sfh_b200.h declares (no compute calls without a GPU), the Python mirror keeps the
reference's error conventions, and there is no silent CPU fallback."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sfh_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sfh_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import sfh_b200
    lib_path = sfh_b200._lib.build()
    assert os.path.exists(lib_path)
    syms = _declared_symbols()
    assert len(syms) >= 14
    raw = ctypes.CDLL(lib_path)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/sfh_b200.h but not exported"
    assert sorted(sfh_b200._lib.SIGNATURES) == syms, "ctypes SIGNATURES table out of sync with the header"
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(syms) <= exported


def test_abi_metadata_and_argument_errors_without_gpu():
    import sfh_b200
    l = sfh_b200._lib.lib()
    assert l.sfh_abi_version() == 1
    assert b"sm_100a" in l.sfh_build_info()
    assert l.sfh_error_string(0) == b"success"
    assert b"workspace" in l.sfh_error_string(-4)
    # workspace size is pure host arithmetic: a fixed 262400-byte ticket / epoch area, B doubles (256-aligned),
    # 12 tagged 8-byte words per partial slot (one slot per 128x8 tile: the upper bound over tile heights),
    # 10 tagged words per sample for the POI block (256-aligned)
    assert l.sfh_workspace_bytes(64, 360, 640) == 262400 + 512 + 64 * (5 * 45) * 12 * 8 + 64 * 10 * 8
    assert l.sfh_workspace_bytes(0, 360, 640) == 0
    # argument validation happens before any CUDA call
    assert l.sfh_warp_fwd(None, None, None, None, 1, 1, 1, 0, None, None) < 0
    assert l.sfh_forward_tail(None, None, None, None, 1, 1, 1, 0, None, None, 0, 0, None, None) < 0
    assert l.sfh_warp_loss_fwd_bwd(None, None, None) < 0 and l.sfh_predict_tail(None, None, None) < 0
    assert l.sfh_poi_fwd(None, None, 0, 1, 1, 1, None, None) == -1
    assert l.sfh_transform_points_fwd(None, 1, None, 2, 3, None, None) == -1
    with pytest.raises(ValueError):
        sfh_b200._lib.check(-1, "x")
    with pytest.raises(RuntimeError):
        sfh_b200._lib.check(700, "x")


def test_sass_is_sm100a_only():
    import sfh_b200
    out = subprocess.run(["cuobjdump", "-lelf", sfh_b200._lib.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    """CPU tensors are refused loudly; nothing in the package imports the oracle."""
    import sfh_b200
    tmpl = torch.zeros(1, 1, 36, 64)
    with pytest.raises(TypeError):
        sfh_b200.CourtTemplate(tmpl)
    with pytest.raises(TypeError):
        sfh_b200.STNWarpStage(tmpl, None, (64, 36), 4)
    with pytest.raises(TypeError):
        sfh_b200.HomographyWarper(36, 64)(tmpl, torch.eye(3)[None])
    with pytest.raises(TypeError):
        sfh_b200.transform_points(torch.eye(3)[None], torch.zeros(1, 5, 2))
    with pytest.raises(TypeError):
        sfh_b200.reprojection_loss(torch.zeros(1, 5, 2), torch.zeros(1, 5, 2), torch.ones(1, 5), torch.ones(1))
    with pytest.raises(NotImplementedError):
        sfh_b200.HomographyWarper(36, 64, padding_mode="border")
    pkg = os.path.join(ROOT, "sports-field-homography_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_missing_library_fails_loudly(tmp_path):
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import sfh_b200\n"
        "sfh_b200._lib.LIB_PATH = %r\n"
        "try:\n"
        "    sfh_b200._lib.lib()\n"
        "except RuntimeError as e:\n"
        "    print('LOUD', e)\n" % (ROOT, str(tmp_path / "nope.so")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "LOUD" in out.stdout and "no CPU or eager fallback" in out.stdout


def test_loaders_mirror_reference_signatures(tmp_path):
    import json
    from PIL import Image
    import sfh_b200
    img = (np.arange(12 * 20).reshape(12, 20) % 4).astype(np.uint8)
    Image.fromarray(img, mode="L").save(tmp_path / "t.png")
    t = sfh_b200.open_court_template(str(tmp_path / "t.png"), 4, (10, 6), 3)
    assert t.shape == (3, 1, 6, 10) and t.dtype == torch.float32
    assert set(np.unique(t.numpy())) <= {0.0, 0.25, 0.5, 0.75}
    pts = {"ranges": [1.0, 1.0], "points": [{"coords": [0.5, 0.25]}, {"coords": [1.0, 0.0]}]}
    (tmp_path / "p.json").write_text(json.dumps(pts))
    p = sfh_b200.open_court_poi(str(tmp_path / "p.json"), 2)
    assert p.shape == (2, 2, 2)
    np.testing.assert_allclose(p[0].numpy(), [[0.0, -0.5], [1.0, -1.0]])
    (tmp_path / "bad.json").write_text("{}")
    with pytest.raises(ValueError):
        sfh_b200.open_court_poi(str(tmp_path / "bad.json"))


def test_synthetic_theta_families_are_seeded_and_sane():
    from sfh_b200 import synth
    a = synth.theta_family_a(8, 1)
    assert a.shape == (8, 1, 3, 3) and torch.equal(a, synth.theta_family_a(8, 1))
    b = synth.theta_family_b(8, 1)
    assert b.shape == (8, 1, 3, 3) and torch.equal(b, synth.theta_family_b(8, 1))
    u = torch.tensor([[-1.0, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]])
    z = torch.einsum("bij,kj->bki", b[:, 0], u)[..., 2]
    assert float((z / b[:, 0, 2, 2][:, None]).min()) > 0.25          # no horizon inside the frame


def test_shard_range_partitions_exactly():
    import sfh_b200
    for n in (0, 1, 7, 64, 65536):
        for world in (1, 2, 3, 8):
            spans = [sfh_b200.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
