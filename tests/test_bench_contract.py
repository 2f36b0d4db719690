"""CPU: the reference arm of bench.py (the part of the driver's contract that runs without a GPU) prints one
JSON line with the agreed keys, for the same metric / unit / workload naming as the CUDA arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--cpu-svd-dim", "256", "--seq-len", "64", "--samples", "8"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "weight_matrices_compressed_per_sec"
    assert line["unit"] == "matrices/s" and line["higher_is_better"] is True and line["steps"] == 2
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the workload is BASELINE configs[1] whatever --steps is; ms_per_step is what was timed, value the job's metric
    assert "NUM_PRUNE_LAYERS=8" in line["config"]["workload"] and line["config"]["extrapolated"] is True
    assert line["ms_per_step"] < 120e3 and line["config"]["extrapolated_job_s"] > 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(os.environ, RANK="1"))
    assert res.returncode == 0 and res.stdout.strip() == ""
