"""GPU, >= 2 devices: the z-slab (multi-GPU) path against the single-GPU path on the same inputs:
S*x to 1e-12, k to 1e-9, flux to 1e-7, parity and fast modes. Launches tests/slab_gpu_worker.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("world", [2, 4])
def test_slab_matches_single_gpu(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world), os.path.join(HERE, "slab_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0
    assert r.stdout.count("OK") >= world
