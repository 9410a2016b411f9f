"""BASELINE.json's full size (the bench workload: 2^20 fluid particles, 4 dynamic boxes): two steps against the CPU oracle
(about a second per step on the box's host cores), and size-independent properties where more steps are needed:
neighbour symmetry, run-to-run bit reproducibility, reset, finite state and density near rest."""
import numpy as np
import pytest

from difffr_b200 import scenes

pytestmark = pytest.mark.gpu
CFG = dict(surface_tension_method=2, surface_tension=0.2, target_time=1.0, max_error=0.05)


@pytest.fixture(scope="module")
def big_scene():
    return scenes.dam_break_scene(1 << 20, n_boxes=4)


def test_two_steps_against_the_oracle_1m(gpu_factory, oracle_factory, big_scene):
    """The bench workload itself, step by step against the oracle: iteration counts and time steps identical, fluid and
    rigid state, forces and all 16 sensitivity blocks of the four boxes within the north-star tolerances (observed ~1e-12)."""
    from conftest import rel_err

    gpu = scenes.build_context(gpu_factory, big_scene, **CFG)
    orc = scenes.build_context(oracle_factory, big_scene, **CFG)
    worst = 0.0
    for _ in range(2):
        gpu.step(1)
        orc.step(1)
        ig, io = gpu.step_info(), orc.step_info()
        assert (ig.iterations, ig.iterations_v) == (io.iterations, io.iterations_v)
        assert abs(ig.time_step_size - io.time_step_size) <= 1e-12 * io.time_step_size
        for f in ("position", "velocity", "density", "factor", "kappa", "kappa_v", "density_adv", "acceleration"):
            e = rel_err(gpu.fluid(f), orc.fluid(f))
            assert e <= 1e-6, (f, e)
            worst = max(worst, e)
        for body in range(1, 5):
            sg, so = gpu.body_state(body), orc.body_state(body)
            for k in sg:
                e = rel_err(sg[k], so[k])
                assert e <= 1e-6, (body, k, e)
                worst = max(worst, e)
            for w in range(16):
                e = rel_err(gpu.body_grad(body, w), orc.body_grad(body, w))
                assert e <= 1e-4, (body, w, e)
                worst = max(worst, e)
    assert worst <= 1e-8
    orc.close()


def test_neighbor_symmetry_and_counts_1m(gpu_factory, big_scene):
    gpu = scenes.build_context(gpu_factory, big_scene, **CFG)
    n = gpu.num_fluid
    assert n >= 1_000_000
    gpu.step(2)
    cnt, idx = gpu.neighbors(-1, -1)
    assert cnt.sum() == idx.size
    # symmetry: the multiset of (i, j) equals the multiset of (j, i)
    src = np.repeat(np.arange(n, dtype=np.int64), cnt)
    fwd = src * n + idx
    bwd = idx.astype(np.int64) * n + src
    fwd.sort()
    bwd.sort()
    assert np.array_equal(fwd, bwd)
    assert 20 < cnt.mean() < 60
    # rows come back sorted and without self / duplicates
    starts = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    d = np.diff(idx.astype(np.int64))
    row_break = np.zeros(idx.size - 1, dtype=bool)
    row_break[starts[1:][cnt[1:] > 0] - 1] = True
    assert np.all((d > 0) | row_break)
    assert not np.any(idx == src)
    # dynamic body -> fluid lists mirror fluid -> body lists
    cb, ib = gpu.neighbors(1, -1)
    cf, jf = gpu.neighbors(-1, 1)
    assert cb.sum() == cf.sum()


def test_bit_reproducible_and_reset_1m(gpu_factory, big_scene):
    a = scenes.build_context(gpu_factory, big_scene, **CFG)
    b = scenes.build_context(gpu_factory, big_scene, **CFG)
    a.step(3)
    b.step(3)
    for f in ("position", "velocity", "kappa", "density"):
        assert np.array_equal(a.fluid(f), b.fluid(f)), f
    for body in range(1, 5):
        sa, sb = a.body_state(body), b.body_state(body)
        for k in sa:
            assert np.array_equal(sa[k], sb[k])
        for w in range(16):
            assert np.array_equal(a.body_grad(body, w), b.body_grad(body, w))
    va = a.fluid("velocity")
    a.reset()
    a.step(3)
    assert np.array_equal(va, a.fluid("velocity"))
    rho = a.fluid("density")
    assert np.all(np.isfinite(rho)) and 500.0 < rho.mean() < 1100.0
    info = a.step_info()
    assert info.total_particle_steps == 3 * a.num_fluid
    assert info.iterations >= 2 and info.iterations_v >= 1
