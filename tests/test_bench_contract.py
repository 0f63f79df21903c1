"""CPU test of bench.py's reference arm: it must print ONE JSON line with the contract keys and
time the CPU restatement (oracle/) on a bounded sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_json():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0", "--workload", "c1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "frames/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_is_silent_on_nonzero_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_both_arms_print_the_same_config_and_metric():
    """The driver compares the two arms' `config` objects: they come from one function of the workload name only, and
    the headline workload is the metric's own configuration (fwd+bwd at 1280x720)."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    a = bench.parse.__globals__["config_of"]("hd")
    assert a == bench.config_of("hd") and a["size"] == [1280, 720] and a["frames_per_step_per_gpu"] == 64
    assert set(a) == {"workload", "frames_per_step_per_gpu", "size", "template", "l2"}
    assert "1280x720" in bench.METRIC and "fwd+bwd" in bench.METRIC
    assert bench.WORKLOADS["c2hd"] is bench.WORKLOADS["hd"]
    # default workload of both arms
    sys.argv, saved = ["bench.py"], sys.argv
    try:
        assert bench.parse().workload == "hd"
    finally:
        sys.argv = saved


def test_rescale_theta_matches_the_reference_formula():
    """dataset_utils/preparation.py:129-137: diag(dst_w, dst_h, 1) @ theta @ diag(1/src_w, 1/src_h, 1) (host arithmetic)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import sfh_b200
    rng = np.random.default_rng(0)
    th = rng.normal(size=(5, 3, 3))
    got = sfh_b200.rescale_theta((1280, 720), (640, 360), th)
    for k in range(5):
        ref = np.matmul(np.matmul(np.diag([640.0, 360.0, 1.0]), th[k]), np.diag([1 / 1280.0, 1 / 720.0, 1.0]))
        assert np.allclose(got[k], ref, rtol=1e-15, atol=0)
    assert sfh_b200.rescale_theta((1280, 720), (640, 360), th[0]).shape == (3, 3)
