"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm's JSON line, the loud
failure of the product arm when there is no device, and the helpers that build the roofline block."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", BENCH)
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "volumes/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1.0       # one volume per step
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_is_rank_zero_only():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
             {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29571"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_fails_loudly_without_a_device():
    r = _run(["--steps", "1", "--warmup", "1", "--no-cpu-baseline"])
    assert r.returncode != 0
    assert not any(l.lstrip().startswith("{") for l in r.stdout.splitlines())     # no JSON line from a CPU fallback


def test_memory_bound_launch_view_follows_launch_order():
    b = _bench_module()
    order = ["conv0_1to16_L0", "conv3_16to16_L0", "pool9_L0", "conv10_16to32_L1", "upconv59_32to8x16_L1", "conv65_16to16_L0"]
    acc = {n: 0.2 for n in order}
    per_launch = [{"dram_gb": g} for g in (0.5, 1.0, 0.1, 0.6, 1.6)]                 # one entry per tcgen05 launch
    out = b.hbm_side(acc, order, per_launch, {"hbm": 6500.0}, b.BATCH)
    assert set(out["launches"]) == {"conv0_1to16_L0", "conv3_16to16_L0", "conv65_16to16_L0"}
    assert out["launches"]["conv65_16to16_L0"]["gbs"] == pytest.approx(1.6 / 0.2e-3, rel=1e-3)
    assert out["launches"]["conv3_16to16_L0"]["frac_of_hbm_peak"] == pytest.approx(5000.0 / 6500.0, abs=1e-3)
    # SURVEY's fused-minimum bytes of the layer (134.2 MB per volume for a 16 -> 16 conv at 128^3) over the same time
    assert out["launches"]["conv3_16to16_L0"]["frac_of_hbm_peak_algorithmic"] == pytest.approx(
        134.2e6 * b.BATCH / 0.2e-3 / 1e9 / 6500.0, abs=1e-3)
    # a capture of another shape (other launch count) or batch is not used
    assert b.hbm_side(acc, order, per_launch[:-1], {"hbm": 6500.0}, b.BATCH) is None
    assert b.hbm_side(acc, order, per_launch, {"hbm": 6500.0}, 2) is None
