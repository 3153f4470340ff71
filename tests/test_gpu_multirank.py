"""Launches tests/multi_gpu_parity.py (NCCL + CUDA kernels on 2 real ranks vs the oracle) when the box has >= 2 GPUs; the
one-GPU box of the round-end run skips it.  A log of a 2-GPU run is kept under profiles/."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_parity():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run: gpurun --gpus 2 -- 'python -m torch.distributed.run ... tests/multi_gpu_parity.py')")
    port = 29600 + (os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_parity.py")],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
