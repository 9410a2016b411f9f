"""Slab domain decomposition (dfr_slab_configure) against the single-context path.  Needs two GPUs: one process per GPU,
NCCL for the particle / ghost exchanges (tests/slab_check.py); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("manager", [0, 1])
def test_two_slabs_equal_one_domain(manager):
    if _gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29517 + manager), os.path.join(ROOT, "tests", "slab_check.py"), "30000", "6", str(manager)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "SLAB_CHECK_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
