"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the CPU
port of the reference path timed on the host cores) prints one JSON line with the keys the driver reads, only rank 0
prints it, and the byte models of the pyramidal workloads are SURVEY 8(d)'s figures."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _run_reference_arm(extra_env, *flags):
    env = dict(os.environ, OF_BENCH_REF_BUDGET_S="2", **extra_env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *flags],
                          env=env, capture_output=True, text=True, timeout=300, cwd=str(ROOT))


def test_reference_arm_prints_the_contract_line():
    res = _run_reference_arm({})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpixel/s" and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["config"]["name"] == "single_1080p" and "workload" in d["config"] and d["config"]["window"] == 5
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 0 < d["pixels_per_step"] <= d["workload_pixels_per_step"]
    # the reference's literal per-pixel loop on one 320 x 240 pair
    assert cb["reference_literal_loop"]["us_per_pixel"] > 0.5


def test_reference_arm_is_silent_on_the_other_ranks():
    res = _run_reference_arm({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert res.returncode == 0, res.stderr[-2000:]
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]


def test_pyramidal_byte_models_and_workload_table():
    import bench

    # 3 levels x 3 iterations and 5 x 10: algorithmic bytes per finest-level pixel
    assert bench.pyramidal_bytes_per_pixel(3, 3) == pytest.approx(109.5, abs=0.1)
    assert bench.pyramidal_bytes_per_pixel(5, 10) == pytest.approx(335.6, abs=0.1)
    # early exits: only executed iterations count
    assert bench.pyramidal_bytes_per_pixel(5, 10, executed=[2.5, 5.5, 5.5, 6, 10]) < bench.pyramidal_bytes_per_pixel(5, 10)
    for name in bench.DEFAULT_EXTRA + ["single_1080p", "pyramidal_8k_exact"]:
        wl = bench.WORKLOADS[name]
        cfg = bench.workload_config(name, wl)
        assert cfg["name"] == name and cfg["window"] == wl.get("window", 5) and f"{cfg['window']}x{cfg['window']} window" in cfg["workload"]
        assert "model" not in cfg


def test_committed_ncu_captures_give_the_traffic_of_their_workloads():
    """roofline.traffic comes from `ncu --set full` captures committed under profiles/: DRAM bytes of the dominant
    kernel per launch.  For the single-launch workloads it must sit at the algorithmic bytes (no wasted re-reads)."""
    import bench

    for name, bpp in (("single_1080p", 16.0), ("single_1080p_u8", 10.0), ("fixed_1080p", 6.0), ("single_1080p_exact", 16.0)):
        wl = bench.WORKLOADS[name]
        traffic, src = bench.ncu_traffic_bytes(name, wl["batch"])
        assert traffic is not None and "profiles/" in src, name
        algorithmic = bpp * wl["batch"] * wl["H"] * wl["W"]
        assert 0.97 < traffic / algorithmic < 1.03, (name, traffic / algorithmic)
    assert bench.ncu_traffic_bytes("single_1080p", 128) == (None, None)  # a capture of another batch does not count
