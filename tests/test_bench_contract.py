"""bench.py prints ONE JSON line with the keys the driver reads (metric / value / roofline /
cpu_baseline / e2e / gpu_launches / clocks); checked on CPU for the reference arm and on a B200 for
the GPU arm at a reduced env count."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def run_bench(*flags, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the oracle port (C restatement of the reference algorithm) on the host cores."""
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        baseline = json.load(f)
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "3", "--ref-budget-s", "4")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == baseline["metric"] and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 1e5 and d["steps"] == 2 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "experiment 6" in d["config"]["workload"] and d["gpu_launches"] == 0


@pytest.mark.gpu
def test_gpu_arm_line():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n = 1 << 20
    d = run_bench("--steps", "20", "--warmup", "30", "--envs-per-gpu", str(n), "--e2e-steps", "3", "--no-cpu-baseline")
    assert (BASE_KEYS - {"cpu_baseline"}) <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["dtype"] == "f32" and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["value"] > 1e9 and d["ms_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0 < rf["frac"] < 1.2
    assert rf["achieved"] == pytest.approx(165 * n / (rf["kernel_ms"] * 1e-3) / 1e9, rel=1e-6)
    assert d["e2e"]["h2d_bytes_per_step"] == 4 * n and d["e2e"]["d2h_bytes_per_step"] == (44 + 4 + 1) * n
    assert 0 < d["e2e"]["value"] < d["value"]
    assert d["gpu_launches"] >= 2 * 20          # one policy launch + one step launch per step
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
