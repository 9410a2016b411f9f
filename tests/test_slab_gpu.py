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


@pytest.mark.parametrize("manager,transport", [(0, "p2p"), (1, "p2p"), (0, "nccl")])
def test_two_slabs_equal_one_domain(manager, transport):
    """transport: ghost rows written by the producing kernels into the neighbour's memory (peer stores, default) or
    shipped by NCCL send/recv after every producing kernel (DFR_SLAB_TRANSPORT=nccl) - the fused k_rho passes carry a
    second ghost array in both cases."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29517 + manager + (4 if transport == "nccl" else 0)), os.path.join(ROOT, "tests", "slab_check.py"), "30000",
           str(6 + manager),  # an odd step count makes slab_check load a state through the per-rank form
           str(manager)]
    env = dict(os.environ)
    if transport == "nccl":
        env["DFR_SLAB_TRANSPORT"] = "nccl"
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "SLAB_CHECK_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
