"""CUDA-graph stepping (dfr_api.cu: StepGraph) against the stream path of the same library.

A replayed step graph holds the Jacobi loops as conditional WHILE nodes (the stopping rules of
TimeStepDiffDFSPH.cpp:711-743 and :828-861 run in k_residual_finish) and never returns to the host; the stream path
(DFR_NO_GRAPH=1) enqueues speculated batches of iterations and reads the convergence flag back.  Where both run the same
kernels they must give the same bits; with the fused non-pressure pass enabled they may differ in which launch carries
it, hence 1e-10 there.  The oracle comparison of the graph path itself is what every other GPU test does (graphs are
the default).
"""
import numpy as np
import pytest

from difffr_b200 import scenes

pytestmark = pytest.mark.gpu

CFG = dict(surface_tension_method=2, surface_tension=0.2, target_time=0.05, max_error=0.05)
FIELDS = ["position", "velocity", "density", "factor", "kappa", "kappa_v", "density_adv", "acceleration", "sum_grad_p_k"]


def snapshot(ctx, scene):
    out = {f: ctx.fluid(f).copy() for f in FIELDS}
    info = ctx.step_info()
    out["info"] = (info.step_count, info.iterations, info.iterations_v, info.time, info.time_step_size, info.trajectory_finished,
                   info.total_pressure_iterations, info.total_divergence_iterations)
    for b, body in enumerate(scene["bodies"]):
        if not body["dynamic"]:
            continue
        for k, v in ctx.body_state(b).items():
            out[f"b{b}.{k}"] = np.array(v, copy=True)
        for w in range(16):
            out[f"b{b}.g{w}"] = np.array(ctx.body_grad(b, w), copy=True)
    return out


def assert_same(a, b, exact=True):
    """exact: same bits.  Otherwise 1e-10 relative: with the fused non-pressure pass enabled the two paths may pick
    different launches to carry it (the stream path speculates in batches, the graph decides per iteration), and the
    fused and the stand-alone kernel sum the same terms in differently contracted arithmetic."""
    assert a.keys() == b.keys()
    for k in a:
        if k == "info":
            assert a[k] == b[k], (a[k], b[k])
        elif exact:
            assert np.array_equal(a[k], b[k], equal_nan=True), k
        else:
            scale = max(float(np.nanmax(np.abs(b[k]))), 1e-300)
            assert float(np.nanmax(np.abs(a[k] - b[k]))) <= 1e-10 * scale, (k, float(np.nanmax(np.abs(a[k] - b[k]))) / scale)


@pytest.mark.parametrize("cfg", [dict(), dict(max_error=0.001, max_error_v=0.01), dict(enable_divergence_solver=0),
                                 dict(use_divergence_warmstart=0), dict(surface_tension_method=0, viscosity_method=0)])
def test_replayed_steps_equal_the_stream_path(gpu_factory, monkeypatch, cfg):
    sc = scenes.dam_break_scene(5000, n_boxes=2)
    kw = dict(CFG)
    kw.update(cfg)
    monkeypatch.setenv("DFR_NO_GRAPH", "1")
    a = scenes.build_context(gpu_factory, sc, **kw)
    a.step(3)
    a.step(9)
    sa = snapshot(a, sc)
    la = a.device_time_ms()[1]
    monkeypatch.delenv("DFR_NO_GRAPH")
    b = scenes.build_context(gpu_factory, sc, **kw)
    b.step(3)   # the first steps after finalize watch the list capacities on the stream path
    b.step(9)   # replayed
    sb = snapshot(b, sc)
    fused = kw.get("enable_divergence_solver", 1) and kw.get("use_divergence_warmstart", 1)
    assert_same(sa, sb, exact=not fused)
    # launch accounting of replayed steps: the same order of magnitude as the stream path (which also counts idle
    # speculated launches), and never zero
    lb = b.device_time_ms()[1]
    assert 0.5 * la <= lb <= 1.5 * la, (la, lb)
    # reset + replay again: graphs stay valid across dfr_reset
    b.reset()
    a.reset()
    monkeypatch.setenv("DFR_NO_GRAPH", "1")
    a.step(7)
    monkeypatch.delenv("DFR_NO_GRAPH")
    b.step(7)
    assert_same(snapshot(a, sc), snapshot(b, sc), exact=not fused)


def test_run_trajectory_batches_stop_at_the_end_of_the_trajectory(gpu_factory, monkeypatch):
    """dfr_run_trajectory enqueues replayed steps in batches; steps past `finished` are skipped on the device, so the step
    count, the final state and a continued dfr_step afterwards equal the step-by-step stream path."""
    sc = scenes.dam_break_scene(3000, n_boxes=1)
    kw = dict(CFG, target_time=0.021, uniform_acc_rb_time=0.0, cfl_method=0, time_step_size=0.001)
    monkeypatch.setenv("DFR_NO_GRAPH", "1")
    a = scenes.build_context(gpu_factory, sc, **kw)
    na = a.run_trajectory(1000)
    sa = snapshot(a, sc)
    a.step(2)
    sa2 = snapshot(a, sc)
    monkeypatch.delenv("DFR_NO_GRAPH")
    for batch in ("", "5", "1"):
        if batch:
            monkeypatch.setenv("DFR_TRAJECTORY_BATCH", batch)
        b = scenes.build_context(gpu_factory, sc, **kw)
        nb = b.run_trajectory(1000)
        assert nb == na and 15 < nb < 30
        assert_same(sa, snapshot(b, sc), exact=False)
        b.step(2)   # buffer parity after skipped steps
        assert_same(sa2, snapshot(b, sc), exact=False)
        # a second trajectory after reset, and max_steps smaller than the trajectory
        b.reset()
        assert b.run_trajectory(7) == 7
        assert b.step_info().step_count == 7


def test_gradient_mode_change_rerecords_the_graphs(gpu_factory, oracle_factory):
    sc = scenes.dam_break_scene(3000, n_boxes=1)
    gpu = scenes.build_context(gpu_factory, sc, **CFG)
    orc = scenes.build_context(oracle_factory, sc, **CFG)
    gpu.step(6)
    orc.step(6)
    gpu.set_gradient_mode(0)
    orc.set_gradient_mode(0)
    gpu.reset()
    orc.reset()
    gpu.step(6)
    orc.step(6)
    for w in range(16):
        g, o = gpu.body_grad(1, w), orc.body_grad(1, w)
        assert np.max(np.abs(g - o)) <= 1e-4 * max(np.max(np.abs(o)), 1e-300), w
