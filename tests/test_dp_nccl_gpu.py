"""GPU (>= 2 devices): real-NCCL data-parallel steps.  DP gradients == sum of the per-rank oracle gradients, the
parameters stay bit-identical across ranks, for both the CUDA-graph path (two graphs with the body all-reduce
overlapped with the subsampler backward) and the eager bucketed path."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_dp_two_gpus_nccl(cuda, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, TASR_TEST_CKPT=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "dp_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "DP_NCCL_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
