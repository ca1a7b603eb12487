"""The reference arm of bench.py (`--impl reference`) runs on host cores only, so its JSON contract can be
checked without a GPU: one line, the base keys, `impl`, a `cpu_baseline` describing the run and an `e2e` block
repeating the value with zero transfer bytes.  Non-zero ranks print nothing and exit 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CMD = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
       "--size", "128", "--batch", "2", "--inputs", "netlike"]


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(CMD, capture_output=True, text=True, timeout=600, env=env, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run(CMD + ["--gpus", "2"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_gpu_arm_prints_one_contract_line():
    """bench.py's own arm on a small batch: one JSON line with the contract's keys, the roofline / e2e / parity
    blocks, a launch count that matches the kernels per step, and inputs that passed the oracle check."""
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3", "--batch", "4", "--streams", "2",
           "--no-cpu-baseline"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "parity", "stage_ms"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] >= 3 and d["dtype"] == "f32" and d["unit"] == "images/s"
    # 6 kernels per step here (aggregation, top-k with 8 warps per row, grouping, adjust/prepare, refine scan, refine
    # apply + records); 7 when the batch is large enough for the two-launch top-k
    assert d["value"] > 0 and d["gpu_launches"] in (6 * 4, 7 * 4) and "model" not in d["config"] and d["config"]["workload"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["parity"]["ok"] is True and d["parity"]["pipelined_records_identical"] is True and d["parity"]["checked"] >= 1
