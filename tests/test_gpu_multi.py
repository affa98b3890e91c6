"""GPU (>= 2 devices): the sharded run over NCCL, launched with torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_gpu_time_block_sharding():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mgpu ok" in r.stdout
