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


def _run(n_particles, steps, manager, port, env_extra=None, burst=0, ranks=2, row_capacity=0):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "slab_check.py"), str(n_particles), str(steps), str(manager), str(burst),
           "2", str(row_capacity)]
    env = dict(os.environ)
    env.update(env_extra or {})
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "SLAB_CHECK_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize("manager,transport", [(0, "p2p"), (1, "p2p"), (0, "nccl")])
def test_two_slabs_equal_one_domain(manager, transport):
    """transport: ghost rows written by the producing kernels into the neighbour's memory (peer stores, default) or
    shipped by NCCL send/recv after every producing kernel (DFR_SLAB_TRANSPORT=nccl) - the fused k_rho passes carry a
    second ghost array in both cases.  Six or seven steps with a reset and a state load in between: all of them on the
    stream path (the four steps after finalize / reset / load are)."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    # an odd step count makes slab_check load a state through the per-rank form
    _run(30000, 6 + manager, manager, 29517 + manager + (4 if transport == "nccl" else 0),
         {"DFR_SLAB_TRANSPORT": "nccl"} if transport == "nccl" else None)


@pytest.mark.parametrize("exchange", ["device", "host", "device-small-rows"])
def test_replayed_slab_steps_equal_one_domain(exchange):
    """16 single steps (reset after 8, state load after 9: steps 5-8 and 14-16 are graph replays) and then two calls of
    dfr_step(5), i.e. replays back to back with nothing read back in between.  exchange: the particle exchange at the head
    of a replayed step over peer memory inside the graph (default) or by NCCL with two host read-backs
    (DFR_SLAB_HOST_EXCHANGE=1); "device-small-rows" starts with ELL rows of 24 fluid / 12 boundary neighbours, so that the
    lists have to grow both on the stream path (rebuild inside the step) and between replayed steps."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _run(30000, 16, 1, 29531 + ["device", "host", "device-small-rows"].index(exchange),
         {"DFR_SLAB_HOST_EXCHANGE": "1"} if exchange == "host" else None, burst=5, row_capacity=24 if exchange.endswith("rows") else 0)
