"""The one-JSON-line contract of bench.py, checked on the lines committed under profiles/ (the evidence the docs quote):
every key the driver reads is present, typed and self-consistent.  CPU only — nothing is executed on a device."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def _line(name):
    path = os.path.join(PROFILES, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("r02_bench_1gpu.json", 1), ("r02_bench_8gpu.json", 8)])
def test_committed_bench_lines_follow_the_contract(name, n):
    d = _line(name)
    assert d["metric"] == "train clips/s" and d["unit"] == "clips/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["dtype"] in ("f16", "bf16")
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    # value = clips of all ranks / max-over-ranks time
    clips = 8 * n * d["steps"]
    assert abs(d["value"] - clips / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == "clips/s" and e["h2d_bytes_per_step"] > 6e7 and e["d2h_bytes_per_step"] == 4 and 0 < e["value"] <= 1.02 * d["value"]
    assert d["gpu_launches"] > 500 * d["steps"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.0 < r["frac"] < 1.0 and (r["traffic"] is None or r["traffic"] > 0)
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    if n == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["unit"] == "clips/s" and b["sample"]
        g = d["gpu_reference"]
        assert g["fp32"]["kind"] == "reference" and g["fp32"]["value"] < d["value"] and g["bf16_autocast"]["value"] < d["value"]
    else:
        k = d["dp_check"]
        assert k["ranks_equal"] is True and k["rel"] < max(2.5 * k["run_to_run_rel"], 5e-3)


def test_committed_reference_arm_line():
    d = _line("r02_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "train clips/s" and d["unit"] == "clips/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
