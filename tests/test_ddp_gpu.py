"""1-GPU vs N-GPU gradient equality after the bucketed all-reduce (test plan item 6;
VERDICT round 1 "missing" #7): torchrun over NCCL, one rank per GPU; skipped on a box with fewer than 2 GPUs
(`gpurun --gpus 2` runs it).  The checks live in tests/ddp_worker.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("deep_supervision", [False, True])
def test_ddp_gradients_equal_single_gpu_mean(deep_supervision):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    env = dict(os.environ, MMR_DDP_DS="1" if deep_supervision else "0", NCCL_DEBUG="WARN")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "ddp_worker.py")]
    res = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "ddp ok: world %d" % world in res.stdout
