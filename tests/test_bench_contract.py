"""bench.py contract on the CPU: the reference arm (`--impl reference`, the oracle port on the host cores — the one arm that
needs no GPU) must print exactly one JSON line with the keys the driver reads, with a bounded sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"]


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, ADB_BENCH_CPU_BATCH="1")            # one waveform per evaluation keeps this to a few seconds
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "diffwave_sc09_edm_heun18_samples_per_sec" and d["unit"] == "samples/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    # "reference" = the reference's own classes (importable here and, through baseline/_ref, on the GPU box); "port" = the oracle
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


import pytest


@pytest.mark.gpu
def test_gpu_arm_prints_one_contract_line():
    """The product arm at a small batch: one JSON line with the roofline / e2e / clocks / launch-count keys, measured on the
    device (kernel launches > 0, e2e with host buffers, no CPU fallback)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--batch", "8", "--steps", "1", "--warmup", "3",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in [k for k in REQUIRED if k not in ("impl", "cpu_baseline")] + ["roofline", "clocks"]:
        assert k in d, k
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["dtype"] == "bf16" and d["scaling"] == "weak"
    assert d["gpu_launches"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 8 * 16000 * 4 == d["e2e"]["d2h_bytes_per_step"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
