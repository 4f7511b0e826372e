"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): packet exchange == dense NCCL all-reduce, replicas bitwise equal."""
import os
import subprocess
import sys

import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


def test_packet_exchange_equals_dense_allreduce_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           str(29600 + os.getpid() % 300), os.path.join(H.ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multigpu ok" in r.stdout
