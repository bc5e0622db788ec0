"""The reference arm of bench.py runs on the CPU box too (it only needs oracle/_ref or the C restatement): its
JSON line must carry the contract's keys, use every host thread even when OMP_NUM_THREADS=1 is exported (torchrun
does that), and time a bounded sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "mini_grid64_lya_lyb", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "spectra_per_s" and line["unit"] == "spectra/s"
    assert line["vs_baseline"] is None and line["higher_is_better"] is True and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == line["value"] and "sightlines" in cb["sample"]
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert cb["cores"] == ncpu, "the reference arm must use every host thread"
    assert line["e2e"] == {"value": line["value"], "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "mini_grid64_lya_lyb"
