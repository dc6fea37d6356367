"""bench.py's output contract, checked without a GPU: the last line a B200 run printed (committed under profiles/) carries every
key the driver reads, and bench.py refuses to run its GPU arm without a CUDA device (no CPU fallback)."""
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest_line():
    files = glob.glob(os.path.join(ROOT, "profiles", "r2_bench_v*_cfg4_1gpu.json"))
    assert files, "no committed bench line under profiles/"
    files.sort(key=lambda f: int(re.search(r"_v(\d+)", f).group(1)))
    with open(files[-1]) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_committed_bench_line_has_the_contract_keys():
    d = _latest_line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "halfway-opt Mpixel-iters/s" and d["unit"] == "Mpixel-iters/s" and d["higher_is_better"] is True
    assert d["scaling"] == "strong" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["workload"].startswith("cfg4: 1280x720 video pair x 120 frames")      # the configuration the metric is quoted on
    assert d["warmup"] >= 3 and d["value"] > 0 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] <= d["value"] * 1.001                    # host buffers inside the timed region: never faster than the resident figure
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is None or r["traffic"] > 0
    f = d["roofline_fp32"]
    assert f["unit"] == "TFLOP/s" and abs(f["frac"] - f["achieved"] / f["peak"]) < 1e-12 and f["attempted_updates_per_launch"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert c["parity_max_dv_px"] == 0.0 and c["parity_iteration_logs_equal"] is True            # GPU == oracle on the CPU sample
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and k["sm_max_mhz"] >= k["sm_mhz"] and isinstance(k["reasons"], list)
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return                                                    # only meaningful on the CPU-only box
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
