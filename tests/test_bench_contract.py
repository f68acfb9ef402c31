"""bench.py contract (CPU only): the committed bench lines carry every key the driver and the tier framing ask for."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_own_arm_line_has_the_contract_keys():
    d = _line("bench_r01.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "gates_per_second" and d["unit"] == "gates/s" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["warmup"] >= 3 and d["n_gpus"] == 1 and d["vs_baseline"] is None and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "QV-33" in d["config"]["workload"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or 0.9 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.1  # no wasted re-reads
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0.8 < e["value"] / d["value"] <= 1.0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    h = d["hbm_bound_regime"]  # explanatory: where memory is the bound the kernel sits at the HBM roofline
    assert h["unit"] == "GB/s" and h["frac"] >= 0.70 and abs(h["frac"] - h["achieved"] / h["peak"]) < 1e-9
    # value is the whole-job aggregate: gates of one circuit / time of one step
    assert abs(d["value"] - 528 / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6


def test_reference_arm_line():
    d = _line("bench_ref_r01.json")
    assert d["impl"] == "reference" and d["metric"] == "gates_per_second" and d["unit"] == "gates/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1


def test_multi_gpu_lines_scale_weakly_without_collectives():
    d = _line("bench_2gpu_r01.json")
    # whole-job aggregate: every rank simulated its own circuit in the (max-over-ranks) step time
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and abs(d["value"] - 2 * 528 / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
    s = d["sharded"]
    assert s["max_abs_err_vs_closed_form"] < 1e-10 and s["qubit_swaps"] >= 1 and 0.5 < s["nvlink_frac_of_770"] <= 1.0


def test_bench_cli_parses(tmp_path):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
