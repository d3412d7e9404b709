"""Real multi-GPU path: one process per GPU, CUDA-IPC peer mappings over NVLink (needs >= 2 GPUs on the box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_shared_model(pkg):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr",
           "127.0.0.1", "--master-port", "29612", os.path.join(ROOT, "tests", "dist_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "dist_worker ewma OK" in out.stdout and "dist_worker lstm OK" in out.stdout
    assert "dist_worker sync OK" in out.stdout and "dist_worker replica sync OK" in out.stdout
    assert "dist_worker sync oracle-equality OK" in out.stdout
