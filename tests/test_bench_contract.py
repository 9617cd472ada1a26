"""bench.py's JSON contract, checked on CPU against the lines recorded on B200 (profiles/) and on the helper logic."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"]


def test_recorded_bench_lines_carry_the_contract():
    for name, n in (("bench_n8_r1_final.json", 8), ("bench_n4_r1.json", 4), ("bench_n2_r1b.json", 2)):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        assert all(k in d for k in REQUIRED), (name, [k for k in REQUIRED if k not in d])
        assert d["metric"] == "wake_h2d_gbs" and d["unit"] == "GB/s" and d["higher_is_better"] is True
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "u8"
        assert "workload" in d["config"] and "model" not in d["config"]
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["bit_exact"] is True
        e = d["e2e"]
        assert e["unit"] == "GB/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.001
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["launches"] > 0
        assert d["clocks"]["sm_max_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_recorded_reference_line():
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_ref_n8_r1.json")))
    assert d["impl"] == "reference" and d["metric"] == "wake_h2d_gbs" and d["n_gpus"] == 8
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_helpers_and_cli_defaults():
    import bench

    assert bench.workload_for(1, None) == "llama-3-8b" and bench.workload_for(8, None) == "llama-3-70b-tp8"
    assert bench.workload_for(4, "mistral-7b") == "mistral-7b"
    peak, src = bench.hbm_peak()
    assert peak > 1000 and ("measured" in src or "fallback" in src)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--impl" in out.stdout and "--gpus" in out.stdout
