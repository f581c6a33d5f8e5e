"""-m gpu: sharded solve over NCCL, one process per GPU.  Needs >= 2 visible GPUs (skipped on a single-GPU box; the
2-rank host logic is covered on CPU by tests/test_distributed_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_two_gpu_sharded_solve():
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "DIST_WORKER_OK 2" in res.stdout
