"""bench.py's output contract, checked without a GPU: the reference arm runs here (the CPU restatement of the reference's
algorithm, oracle/crossvar_ref.c - the one place besides tests/ where bench.py may execute oracle/), and the committed
line of the final GPU run carries every key the driver reads, with numbers that are consistent with each other."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"]


def test_reference_arm_runs_on_the_cpu_and_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "mini",
                        "--steps", "2", "--warmup", "1", "--cpu-sample", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                  # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in BASE_KEYS:
        assert k in d, k
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0                           # nothing of the product runs in this arm


def test_committed_gpu_line_is_complete_and_consistent():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_final_1gpu.json")))
    for k in BASE_KEYS + ["clocks", "roofline", "cpu_baseline"]:
        assert k in d, k
    assert "impl" not in d and d["n_gpus"] == 1 and d["steps"] >= 5 and d["warmup"] >= 3
    assert d["metric"].startswith("cross variance-covariance predictions/sec") and d["unit"] == "crosses/s"
    assert "33000 SNPs" in d["config"]["workload"] and "50000 crosses" in d["config"]["workload"]        # BASELINE.json configs[3]
    n = d["config"]["crosses_total"]
    assert abs(d["value"] - n / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    # end to end through the C ABI from host buffers: copies declared, and slower than the device-resident figure
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] >= n * 15 * 2 * 8 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["kernel_ms"] < d["ms_per_step"]                # the dominant kernel fits inside the step ...
    assert abs(r["achieved"] - r["algorithmic_ops_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e12) < 1e-6 * r["achieved"]
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] <= c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    # parity was checked inside the same run, against the oracle, within the tolerance north_star states
    for mode, p in d["parity"].items():
        assert p["nsegsnps_exact"] and p["max_rel_err_all_guarded"] < p["tolerance"] == 1e-9, mode
        assert p["max_rel_err_variances_strict"] < 1e-9, mode
