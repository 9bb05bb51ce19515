"""bench.py's reference arm runs on the CPU alone (the oracle on a bounded sample of configs[1]): check the
JSON line the driver parses.  The GPU arm of the same contract is exercised on the GPU box by the driver."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=str(ROOT), capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"] == "echelonize_time_to_rank" and line["unit"] == "s" and line["higher_is_better"] is False
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    # the CPU arm times a BOUNDED SAMPLE (fewer rows of the same generator): that is not the metric of the 200000-row
    # workload, so the top-level value is null and the comparable numbers are numeric fields of cpu_baseline
    assert line["value"] is None and "value_note" in line
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb and cb["sample_n"] >= 4000
    assert abs(line["ms_per_step"] - 1e3 * cb["value"]) < 1e-6
    assert "gpu_same_sample_s" in cb and "gpu_same_sample_e2e_s" in cb
    e = line["e2e"]
    assert e["value"] is None and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["vs_baseline"] is None


def test_product_arm_fails_loudly_without_gpu():
    """no CPU fallback: on a box without a GPU the product arm must refuse, not compute"""
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, "bench.py", "--rows", "500", "--steps", "1", "--warmup", "0", "--no-cpu", "--no-secondary"], cwd=str(ROOT),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert not any(l.startswith('{"metric"') for l in r.stdout.splitlines())
