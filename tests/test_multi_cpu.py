"""The N > 1 host logic on CPU, two processes over gloo:
  * population sharding (difffr_b200/rollouts.py): two ranks evaluating one population return exactly what one process
    returns, in candidate order.  The rollouts themselves run on the CPU oracle here (test infrastructure standing in for
    the CUDA contexts, which need a GPU) - the code under test is the sharding / gathering, which is backend independent;
  * slab planning (dfr_slab_plan, the cut dfr_finalize uses): every particle is owned by exactly one slab, counts are
    balanced, and too many ranks for a thin scene are refused.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

CANDIDATES = [((0.2 * k, -0.1, 0.05 * k), (0.3, 0.5 * k, -0.2)) for k in range(5)]


def _oracle_context():
    import ctypes

    from difffr_b200 import scenes
    from difffr_b200.cabi import Context

    olib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    sc = scenes.dam_break_scene(600, n_boxes=1)
    return scenes.build_context(lambda **k: Context(lib=olib, prefix="orc_", **k), sc, target_time=0.004, max_error=0.05)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from difffr_b200 import cabi, rollouts, scenes

    res, steps = rollouts.run_population(_oracle_context, CANDIDATES, body=1, rank=rank, world_size=world, gather=rollouts.torch_gather(world))
    np.save(os.path.join(out_dir, f"pop_{rank}.npy"), np.concatenate([res, steps[:, None].astype(np.float64)], axis=1))
    # slab ownership: both ranks plan the same cut; the union of what they own is everything, once
    sc = scenes.dam_break_scene(20000, n_boxes=0, jitter=0.3, seed=rank * 0 + 5)
    z = sc["fluid"][:, 2]
    cell = 2 * sc["radius"] * (1 + 1e-7)
    z0 = z.min() - 4 * cell
    nz = int(np.floor((z.max() - z0) / cell)) + 5
    planes = cabi.slab_plan(z, z0, cell, nz, 2, world)
    zc = np.clip(np.floor((z - z0) * (1.0 / cell)).astype(np.int64), 0, nz - 1)
    mine = torch.from_numpy(((zc >= planes[rank]) & (zc < planes[rank + 1])).astype(np.int64))
    count = torch.tensor([int(mine.sum())])
    dist.all_reduce(mine)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    assert bool((mine == 1).all()), "every particle must be owned by exactly one slab"
    counts = np.array([int(c) for c in counts])
    assert counts.sum() == len(z) and counts.max() - counts.min() <= 0.12 * len(z), counts
    dist.destroy_process_group()


def test_population_sharding_and_slab_ownership_two_ranks(tmp_path, oracle_lib):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "pop_0.npy"), np.load(tmp_path / "pop_1.npy")
    assert np.array_equal(a, b)  # every rank ends with the full, identically ordered result table
    from difffr_b200 import rollouts

    single, steps = rollouts.run_population(_oracle_context, CANDIDATES, body=1)
    assert np.array_equal(a[:, :-1], single) and np.array_equal(a[:, -1].astype(np.int64), steps)
    r0 = rollouts.unpack_result(single[0])
    assert r0["grad_x_to_v0"].shape == (3, 3) and r0["grad_quaternion_to_omega0"].shape == (4, 3)
    assert not np.array_equal(single[0], single[4])  # different candidates, different trajectories


def test_shard_indices_cover_everything_once():
    from difffr_b200 import rollouts

    for n in (0, 1, 5, 64, 65):
        for world in (1, 2, 3, 8):
            got = sum((rollouts.shard_indices(n, r, world) for r in range(world)), [])
            assert got == list(range(n))
            sizes = [len(rollouts.shard_indices(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rollouts.shard_indices(4, 2, 2)


def test_slab_plan_balances_and_refuses_thin_slabs():
    from difffr_b200 import cabi

    rng = np.random.default_rng(0)
    z = rng.uniform(0.0, 1.0, 50000)
    cell = 0.01
    for world in (1, 2, 4, 8):
        planes = cabi.slab_plan(z, 0.0, cell, 100, 2, world)
        assert planes[0] == 0 and planes[-1] == 100 and np.all(np.diff(planes) > 0)
        zc = np.clip(np.floor(z / cell).astype(int), 0, 99)
        counts = np.array([np.sum((zc >= planes[r]) & (zc < planes[r + 1])) for r in range(world)])
        assert counts.sum() == z.size and counts.max() - counts.min() <= 2 * z.size / 100  # within two layers of particles
    with pytest.raises(cabi.DfrError):
        cabi.slab_plan(z, 0.0, cell, 100, 2, 32)  # 100 layers / 32 ranks < 5 layers per slab
    # strongly skewed distribution: planes follow the particles, not the geometry
    z2 = np.concatenate([rng.uniform(0.0, 0.2, 40000), rng.uniform(0.2, 1.0, 10000)])
    planes = cabi.slab_plan(z2, 0.0, cell, 100, 2, 2)
    assert planes[1] < 20


def test_population_with_concurrent_contexts_on_the_oracle(oracle_lib):
    """run_population(concurrency=k): k contexts driven by k host threads give the serial results in candidate order (the
    thread pool / work queue is backend independent; on the GPU the contexts' kernels overlap, tests/test_rollouts_gpu.py)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from difffr_b200 import rollouts

    serial, steps1 = rollouts.run_population(_oracle_context, CANDIDATES, body=1)
    stats = {}
    conc, steps2 = rollouts.run_population(_oracle_context, CANDIDATES, body=1, concurrency=3, stats=stats)
    assert np.array_equal(steps1, steps2)
    assert np.array_equal(serial, conc)
    assert "kernel_launches" in stats
    # contexts handed in by the caller; more contexts than candidates are not used
    ctxs = [_oracle_context() for _ in range(2)]
    again, _ = rollouts.run_population(None, CANDIDATES[:1], body=1, contexts=ctxs)
    assert np.array_equal(again, serial[:1])
