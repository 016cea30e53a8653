"""The bench lines committed under profiles/ (copied from real B200 runs of bench.py) carry every key of the bench
contract; bench.py's argument defaults stay within it.  Guards the JSON contract on the CPU tier."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("r1_bench_n1.json", 1), ("r1_bench_n2.json", 2)])
def test_committed_bench_lines_follow_the_contract(name, n):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == n and d["unit"] == "sims/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"] and "8192" in d["metric"]
    assert d["value"] == pytest.approx(d["tree_stats"]["sims"] * n / (d["ms_per_step"] * d["steps"] / 1000.0), rel=0.02)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["traffic"] and r["traffic"] >= r["alg_bytes_per_launch"]
    rn = d["roofline_nn"]
    assert rn["bound"] == "tensor" and rn["unit"] == "TFLOP/s" and 0.2 < rn["frac"] < 1.0
    e = d["e2e"]
    assert e["unit"] == "sims/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    c = d["clocks"]
    assert c["sm_mhz"] > 0.8 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        b = d["cpu_baseline"]
        assert b["kind"] == "reference" and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
        assert d["e2e_coach"]["value"] > e["value"]
    else:
        assert d["cpu_baseline"] is None and d["example_gather"]["samples"] > 0


def test_reference_arm_line():
    d = _line("r1_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["unit"] == "sims/s" and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
