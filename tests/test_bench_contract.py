"""bench.py prints ONE JSON line with the keys the driver reads: the reference arm on the CPU (here), the B200 arm with its
`roofline`, `e2e`, `clocks` and `gpu_launches` on the GPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_on_the_host_cores():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "sample-steps/s"
    assert d["config"]["workload"].startswith("config2") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    from oracle import reference
    assert cb["kind"] == ("reference" if reference.available() else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


@pytest.mark.gpu
def test_b200_arm_line_carries_roofline_e2e_clocks_and_launch_count():
    d = run_bench("--steps", "3", "--warmup", "3", "--batch", "4", "--no-cpu-baseline", "--no-eager-baseline", "--train-batch", "2")
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert d["metric"] == "unet3d_denoise_sample_steps_per_sec_16x64x64" and d["higher_is_better"] is True
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] <= d["value"] * 1.05
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0
    assert abs(r["frac_burst"] - r["achieved"] / r["peak_burst"]) < 1e-6
    t = d["train"]   # BASELINE.json configs[3]: the data-parallel training step, the one place the path has a collective
    assert t["value"] > 0 and t["unit"] == "samples/s" and t["batch_per_gpu"] == 2 and t["gpu_launches"] > 0
    assert t["collective"] is None and t["allreduce_exposed_ms"] is None   # one GPU: nothing to reduce
