"""Multi-GPU parity of the product's sharded path on real GPUs (needs >= 2 GPUs: skipped on a one-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).  The worker (tests/mgpu_worker.py) runs
Spectra(shard="sightlines") and Spectra(shard="particles") under torchrun/NCCL and compares with one GPU: bit-identical
rows for sightline sharding, <= 1e-12 for particle sharding."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_spectra_on_gpus(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29600 + world
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    rep = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert rep["ok"] and rep["tau_1215_bitwise"] and rep["tau_1025_bitwise"] and rep["colden_bitwise"]
    assert rep["particles_max_rel"] <= 1e-12
    print(json.dumps(rep))
