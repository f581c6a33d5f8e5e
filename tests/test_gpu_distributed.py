"""-m gpu: sharded solve, one process per rank.

* two ranks over NCCL on two GPUs (skipped on a single-GPU box);
* two ranks SHARING cuda:0 (runs everywhere): process group over gloo, the per-iteration exchange through CUDA-IPC windows
  and in-kernel arrival flags exactly as on several GPUs -- the two processes time-slice the device.
The 2-rank host logic is also covered on CPU by tests/test_distributed_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _launch(port, extra_env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "dist_worker.py")]
    env = dict(os.environ)
    env.update(extra_env)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "DIST_WORKER_OK 2" in res.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_two_gpu_sharded_solve():
    _launch(29533, {})


def test_two_rank_sharded_solve_sharing_one_gpu():
    _launch(29534, {"DUALIP_TEST_ONE_GPU": "1", "DUALIP_PEER_TIMEOUT_MS": "8000"})
