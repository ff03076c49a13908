"""Sharded cloud on real GPUs (needs >= 2): in-kernel NVLink all-reduce of the update sums, and the
request/response resample exchange over NCCL.  Skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_sharded_updater_matches_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + "\n" + r.stderr[-4000:]
