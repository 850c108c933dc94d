"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): the data-parallel train step on real devices over NCCL.
tools/dp_check.py asserts that the bucketed all-reduce overlapped with backward yields the gradient of the
single-device step on the concatenated batch, and that a few DP steps with the optimizer run to completion."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_step_matches_single_device_gradient():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "tools", "dp_check.py"), "8"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=170, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("dp_check ok") == 2
