"""bench.py contract: one JSON line with the keys the driver reads.  The reference arm (CPU, oracle port) runs anywhere; the
GPU arm is checked on the GPU box with a handful of steps."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    j = _run("--impl", "reference", "--steps", "4", "--warmup", "1")
    assert BASE_KEYS <= set(j) and j["impl"] == "reference" and j["metric"] == "env-steps/sec" and j["unit"] == "env-steps/s"
    assert j["value"] > 0 and j["higher_is_better"] is True and j["vs_baseline"] is None and j["dtype"] == "f64"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    j = _run("--steps", "6", "--warmup", "3")
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "timing"} <= set(j)
    assert j["n_gpus"] == 1 and j["steps"] == 6 and j["warmup"] == 3 and j["dtype"] == "f32" and j["scaling"] == "weak"
    assert j["gpu_launches"] == 6               # one step kernel per step (4096 envs fit one wave: no warp re-grouping launches)
    assert j["value"] > 1e6 and abs(j["ms_per_step"] * 1e-3 * j["value"] - j["config"]["envs_per_gpu"]) < 1e-3 * j["config"]["envs_per_gpu"]
    t = j["timing"]                             # >= 100 ms of kernel time whatever --steps says, median over blocks
    assert t["block_steps"] == 6 and t["blocks"] >= 5 and len(t["block_ms"]) == t["blocks"] and t["window_ms"] >= 100.0
    assert len(t["per_rank_ms_per_step"]) == 1 and abs(sorted(t["block_ms"])[t["blocks"] // 2] / 6 - j["ms_per_step"]) < 0.2 * j["ms_per_step"]
    rf = j["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert j["e2e"]["value"] > 1e6 and j["e2e"]["h2d_bytes_per_step"] == 4096 * 2 * 4 and j["e2e"]["d2h_bytes_per_step"] > 0
    assert j["e2e"]["value"] < j["value"]
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] > 1e3
    c = j["config"]
    assert c["error_flags_or"] == 0 and 0.0 <= c["resets_per_step"] < 0.02
    x = c["extras"]
    for name in ("walk_ik_demo_clock_x4", "north_star_size", "north_star_size_demo_clock_x4", "C3_gallop_ol_rand_gains",
                 "C4_turn_ik_heightfield", "C5_standup_arm_18dof"):
        assert "error" not in x[name], (name, x[name])
        assert x[name]["value"] > 1e6 and (x[name]["error_flags_or"] & 1) == 0, (name, x[name])
    assert x["north_star_size"]["envs_per_gpu"] == 65536 and x["north_star_size"]["value"] > 1e7
    assert x["walk_ik_demo_clock_x4"]["resets_per_step"] < c["resets_per_step"] + 1e-9      # walking robots do not fall
    assert "error" not in x["rollout_with_policy"]
