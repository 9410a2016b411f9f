"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same seeded inputs.

Tolerances are the north-star's: neighbour sets identical (compared sorted), per-step particle and rigid
state within 1e-6 relative in FP64, sensitivities d(state)/d(v0, omega0) within 1e-4 relative.
Relative error of a field = max|a-b| / max|b|.
"""
import numpy as np
import pytest

from conftest import rel_err
from difffr_b200 import scenes
from difffr_b200.cabi import GRAD_NAMES, DfrError

pytestmark = pytest.mark.gpu

STATE_TOL = 1e-6
GRAD_TOL = 1e-4
FLUID_FIELDS = ["position", "velocity", "density", "factor", "kappa", "kappa_v", "density_adv", "acceleration", "sum_grad_p_k"]
BASE_CFG = dict(surface_tension_method=2, surface_tension=0.2, target_time=0.05, max_error=0.05)


def build_pair(gpu_factory, oracle_factory, scene, **cfg):
    kw = dict(BASE_CFG)
    kw.update(cfg)
    return scenes.build_context(gpu_factory, scene, **kw), scenes.build_context(oracle_factory, scene, **kw)


def dynamic_bodies(ctx, scene):
    return [i for i, b in enumerate(scene["bodies"]) if b["dynamic"]]


def compare_step(gpu, orc, scene, state_tol=STATE_TOL, grad_tol=GRAD_TOL, grads=True):
    ig, io = gpu.step_info(), orc.step_info()
    assert ig.step_count == io.step_count
    assert ig.iterations == io.iterations and ig.iterations_v == io.iterations_v, (ig.iterations, io.iterations, ig.iterations_v, io.iterations_v)
    assert abs(ig.time_step_size - io.time_step_size) <= 1e-12 * io.time_step_size
    assert abs(ig.time - io.time) <= 1e-12 * max(io.time, 1e-30)
    assert ig.num_fluid_particles == io.num_fluid_particles
    worst = 0.0
    for f in FLUID_FIELDS:
        e = rel_err(gpu.fluid(f), orc.fluid(f))
        assert e <= state_tol, (f, e)
        worst = max(worst, e)
    for b in dynamic_bodies(gpu, scene):
        sg, so = gpu.body_state(b), orc.body_state(b)
        for k in sg:
            e = rel_err(sg[k], so[k])
            assert e <= state_tol, (b, k, e)
            worst = max(worst, e)
        pg, po = gpu.body_properties(b), orc.body_properties(b)
        fscale = max(np.max(np.abs(po["force"])), np.max(np.abs(po["torque"])), 1e-300)
        assert np.max(np.abs(pg["force"] - po["force"])) <= state_tol * fscale
        assert np.max(np.abs(pg["torque"] - po["torque"])) <= state_tol * fscale
        if grads:
            for w in range(16):
                a, o = gpu.body_grad(b, w), orc.body_grad(b, w)
                e = rel_err(a, o)
                assert e <= grad_tol, (b, GRAD_NAMES[w], e)
    return worst


def compare_neighbors(gpu, orc, scene):
    pairs = [(-1, -1)] + [(-1, b) for b in range(len(scene["bodies"]))] + [(b, -1) for b in dynamic_bodies(gpu, scene)]
    for a, b in pairs:
        cg, ig = gpu.neighbors(a, b)
        co, io = orc.neighbors(a, b)
        assert np.array_equal(cg, co), (a, b)
        assert np.array_equal(ig, io), (a, b)


def test_boundary_volumes_and_neighbor_sets(gpu_factory, oracle_factory):
    sc = scenes.dam_break_scene(6000, n_boxes=2)
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc)
    for b in range(gpu.num_bodies):
        assert rel_err(gpu.body_particles(b, "volume"), orc.body_particles(b, "volume")) <= 1e-12
        assert rel_err(gpu.body_particles(b, "position"), orc.body_particles(b, "position")) <= 1e-14
    compare_neighbors(gpu, orc, sc)  # lattice state: many pairs at distance == support radius up to rounding
    gpu.step(3)
    orc.step(3)
    compare_neighbors(gpu, orc, sc)


@pytest.mark.parametrize("gradient_mode", [1, 0, 2])
def test_per_step_state_and_sensitivities(gpu_factory, oracle_factory, gradient_mode):
    sc = scenes.dam_break_scene(4000, n_boxes=1)
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc, gradient_mode=gradient_mode)
    worst = 0.0
    for _ in range(8):
        gpu.step(1)
        orc.step(1)
        worst = max(worst, compare_step(gpu, orc, sc))
    assert worst <= 1e-9  # observed ~1e-13: only summation order differs


@pytest.mark.parametrize("mode", ["1", "2", "3"])
def test_pass_fusion_modes_agree(gpu_factory, oracle_factory, mode, monkeypatch):
    """density/factor, normals and the non-pressure forces ride on k_rho passes of the divergence solve by default
    (DESIGN.md §4).  DFR_NO_FUSION=1 runs every pass as its own kernel, =2 only the non-pressure pass, =3 runs the
    fused non-pressure pass and then discards it (the path taken when the solve needs more iterations than
    speculated): all of them must match the oracle and the default path."""
    sc = scenes.dam_break_scene(4000, n_boxes=1)
    ref = scenes.build_context(gpu_factory, sc, **BASE_CFG)
    ref.step(6)
    monkeypatch.setenv("DFR_NO_FUSION", mode)  # read by dfr_step at every step
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc)
    for _ in range(6):
        gpu.step(1)
        orc.step(1)
        compare_step(gpu, orc, sc)
    for f in FLUID_FIELDS:
        assert rel_err(gpu.fluid(f), ref.fluid(f)) <= 1e-11, f
    sg, sr = gpu.body_state(1), ref.body_state(1)
    for k in sg:
        assert rel_err(sg[k], sr[k]) <= 1e-11, k


@pytest.mark.parametrize("cfg", [
    dict(rigid_body_mode=1),
    dict(optimize_rotation=0),
    dict(surface_tension_method=0, viscosity_method=0),
    dict(use_pressure_warmstart=0, use_divergence_warmstart=0),
    dict(enable_divergence_solver=0),
    dict(cfl_method=2),
    dict(cfl_method=0, time_step_size=5e-4),
])
def test_solver_and_rigid_options(gpu_factory, oracle_factory, cfg):
    sc = scenes.dam_break_scene(3000, n_boxes=1)
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc, **cfg)
    for _ in range(5):
        gpu.step(1)
        orc.step(1)
        compare_step(gpu, orc, sc)


def test_gradient_manager_two_bodies(gpu_factory, oracle_factory):
    sc = scenes.dam_break_scene(5000, n_boxes=2)
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc, use_rigid_gradient_manager=1)
    for _ in range(6):
        gpu.step(1)
        orc.step(1)
        compare_step(gpu, orc, sc, grads=False)
    dyn = dynamic_bodies(gpu, sc)
    for R in dyn:
        for RR in dyn:
            for w in range(16):
                e = rel_err(gpu.manager_grad(R, RR, w), orc.manager_grad(R, RR, w))
                assert e <= GRAD_TOL, (R, RR, GRAD_NAMES[w], e)


def test_init_velocity_ramp_and_trajectory_end(gpu_factory, oracle_factory):
    """uniformAccelerateRBTime protocol (TimeStepDiffDFSPH.cpp:379-429) and the trajectory-finished flag (:448)."""
    sc = scenes.dam_break_scene(3000, n_boxes=1)
    sc["bodies"][1]["init_v"] = (0.8, -0.5, 0.1)
    sc["bodies"][1]["init_omega"] = (1.0, 2.0, -0.5)
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc, uniform_acc_rb_time=0.004, target_time=0.008)
    ng = gpu.run_trajectory(200)
    no = orc.run_trajectory(200)
    assert ng == no and ng < 200
    assert gpu.step_info().trajectory_finished == 1
    compare_step(gpu, orc, sc)


def test_jittered_positions_and_loaded_state(gpu_factory, oracle_factory):
    sc = scenes.dam_break_scene(5000, n_boxes=1, jitter=0.3, seed=7)
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc)
    rng = np.random.default_rng(3)
    n = gpu.num_fluid
    x = sc["fluid"]
    v = rng.normal(scale=0.2, size=(n, 3))
    kap = -1e-6 * rng.random(n)
    kapv = -1e-3 * rng.random(n)
    gpu.load_fluid_state(x, v, kap, kapv)  # --load-fluid-pos-and-vel + kappa fields of the state file
    orc.load_fluid_state(x, v, kap, kapv)
    compare_neighbors(gpu, orc, sc)
    for _ in range(5):
        gpu.step(1)
        orc.step(1)
        compare_step(gpu, orc, sc)


def test_reset_restores_the_initial_state_bit_exactly(gpu_factory):
    sc = scenes.dam_break_scene(4000, n_boxes=1)
    gpu = scenes.build_context(gpu_factory, sc, **BASE_CFG)
    gpu.step(4)
    a = {f: gpu.fluid(f) for f in ("position", "velocity", "kappa")}
    sa = gpu.body_state(1)
    ga = [gpu.body_grad(1, w) for w in range(8)]
    gpu.reset()
    assert gpu.step_info().step_count == 0
    assert np.array_equal(gpu.fluid("position"), sc["fluid"])
    gpu.step(4)
    for f in a:
        assert np.array_equal(a[f], gpu.fluid(f)), f  # run-to-run bit reproducible (fixed-order reductions)
    sb = gpu.body_state(1)
    for k in sa:
        assert np.array_equal(sa[k], sb[k])
    for w in range(8):
        assert np.array_equal(ga[w], gpu.body_grad(1, w))


def test_neighbour_capacities_grow_on_demand(gpu_factory, oracle_factory):
    """The reference's neighbour lists have no capacity (vector<vector<unsigned>>).  The ELL rows here do; a row that
    does not fit makes the context grow its capacities and rebuild the lists before anything else of the step has run."""
    sc = scenes.dam_break_scene(4000, n_boxes=1)
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc, neighbor_capacity_fluid=8, neighbor_capacity_boundary=4, body_neighbor_capacity=4)
    compare_neighbors(gpu, orc, sc)
    for _ in range(4):
        gpu.step(1)
        orc.step(1)
        compare_step(gpu, orc, sc)
    compare_neighbors(gpu, orc, sc)


@pytest.mark.parametrize("case", ["fluid_only", "static_only", "sparse", "escaping"])
def test_edge_cases(gpu_factory, oracle_factory, case):
    r = 0.025
    if case == "fluid_only":  # no boundary at all
        sc = dict(radius=r, fluid=scenes.fluid_block((0, 0, 0), (0.5, 0.5, 0.5), r), bodies=[])
    elif case == "static_only":  # tank, no dynamic body: nothing to differentiate
        sc = scenes.dam_break_scene(2000, n_boxes=0)
    elif case == "sparse":  # fewer than 20 neighbours everywhere: the particle-deficiency cut (:2027-2040)
        sc = dict(radius=r, fluid=scenes.fluid_block((0, 0, 0), (0.2, 0.15, 0.15), r), bodies=[])
    else:  # particles far outside the initial bounding box: clamped cells stay exhaustive
        f = scenes.fluid_block((0, 0, 0), (0.4, 0.4, 0.4), r)
        v = np.zeros_like(f)
        v[:, 0] = 40.0 * (f[:, 0] > 0.2)
        sc = dict(radius=r, fluid=f, fluid_velocity=v, bodies=[])
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc)
    compare_neighbors(gpu, orc, sc)
    for _ in range(12 if case == "escaping" else 4):
        gpu.step(1)
        orc.step(1)
        compare_step(gpu, orc, sc)
    compare_neighbors(gpu, orc, sc)


def test_api_errors(gpu_factory):
    ctx = gpu_factory()
    with pytest.raises(DfrError):
        ctx.step(1)  # not finalized
    with pytest.raises(DfrError):
        ctx.finalize()  # empty scene
    sc = scenes.dam_break_scene(1000, n_boxes=1)
    gpu = scenes.build_context(gpu_factory, sc, **BASE_CFG)
    with pytest.raises(DfrError):
        gpu.body_state(7)
    with pytest.raises(DfrError):
        gpu.body_grad(1, 99)
    with pytest.raises(DfrError):
        gpu.set_fluid(sc["fluid"])  # after finalize


def test_medium_scene_parity_50k(gpu_factory, oracle_factory):
    sc = scenes.dam_break_scene(50000, n_boxes=4)
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc)
    compare_neighbors(gpu, orc, sc)
    for _ in range(3):
        gpu.step(1)
        orc.step(1)
        compare_step(gpu, orc, sc)


def test_emitter_grows_the_fluid(gpu_factory, oracle_factory):
    """Box emitter (Emitter.cpp:89-227, config 4 of BASELINE.json): emitted particles, AnimatedByEmitter state and the
    particle count agree step by step."""
    sc = scenes.dam_break_scene(2000, n_boxes=1)
    he = sc["tank_half_extent"]
    rot = np.array([0, 1, 0, -1, 0, 0, 0, 0, 1], dtype=np.float64)  # emit direction (first column) = -y
    sc["emitters"] = [dict(width=4, height=3, position=(0.3 * he[0], 1.2 * he[1], 0.0), rotation=rot, velocity=2.0, emit_start=0.0, emit_end=0.03),
                      dict(width=2, height=2, position=(0.6 * he[0], 1.1 * he[1], 0.1), rotation=rot, velocity=3.0, emit_start=0.004, emit_end=0.02)]
    # Surface tension is off here on purpose: emitted sheets form an exact 2r lattice whose in-sheet / sheet-to-sheet
    # pairs sit at distance == support radius.  Akinci-2013's normal term -k (n_i - n_j) is not kernel-weighted
    # (SurfaceTension_Akinci2013.cpp:104-110), so whether such a pair is in the list changes the force by O(1): a
    # discontinuity of the reference algorithm itself once released particles differ in the last bit.
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc, max_emitted_particles=200, cfl_max_time_step=0.002, target_time=1.0,
                          surface_tension_method=0)
    n0 = gpu.num_fluid
    for s in range(30):
        gpu.step(1)
        orc.step(1)
        assert gpu.num_fluid == orc.num_fluid, s
        compare_step(gpu, orc, sc, state_tol=1e-6)
    assert gpu.num_fluid > n0
    gpu.reset()
    orc.reset()
    assert gpu.num_fluid == n0 == orc.num_fluid
    gpu.step(12)
    orc.step(12)
    compare_step(gpu, orc, sc)


def test_emitter_with_surface_tension(gpu_factory, oracle_factory):
    """The same emitters with Akinci-2013 surface tension ON, as in diff-high-diving-duck.json (BASELINE.json configs[3]).
    Emitted positions are computed with the reference's rounding (k_emit_spawn), so the sheets' pairs at distance ==
    support radius fall on the same side of the strict `<` on both sides and the O(1) normal term agrees."""
    sc = scenes.dam_break_scene(2000, n_boxes=1)
    he = sc["tank_half_extent"]
    rot = np.array([0, 1, 0, -1, 0, 0, 0, 0, 1], dtype=np.float64)
    sc["emitters"] = [dict(width=4, height=3, position=(0.3 * he[0], 1.2 * he[1], 0.0), rotation=rot, velocity=2.0, emit_start=0.0, emit_end=0.03),
                      dict(width=2, height=2, position=(0.6 * he[0], 1.1 * he[1], 0.1), rotation=rot, velocity=3.0, emit_start=0.004, emit_end=0.02)]
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc, max_emitted_particles=200, cfl_max_time_step=0.002, target_time=1.0,
                          surface_tension_method=2, surface_tension=0.5)
    n0 = gpu.num_fluid
    for s in range(20):
        gpu.step(1)
        orc.step(1)
        assert gpu.num_fluid == orc.num_fluid, s
        compare_step(gpu, orc, sc, state_tol=1e-6)
    assert gpu.num_fluid > n0
    compare_neighbors(gpu, orc, sc)


def test_emitter_capacity_is_respected(gpu_factory, oracle_factory):
    sc = scenes.dam_break_scene(1500, n_boxes=0)
    he = sc["tank_half_extent"]
    rot = np.array([0, 1, 0, -1, 0, 0, 0, 0, 1], dtype=np.float64)
    sc["emitters"] = [dict(width=3, height=3, position=(0.3 * he[0], 1.2 * he[1], 0.0), rotation=rot, velocity=4.0, emit_start=0.0, emit_end=1.0)]
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc, max_emitted_particles=20, cfl_max_time_step=0.002, target_time=1.0,
                          surface_tension_method=0)
    n0 = gpu.num_fluid
    for s in range(40):
        gpu.step(1)
        orc.step(1)
        assert gpu.num_fluid == orc.num_fluid <= n0 + 20
    assert gpu.num_fluid == n0 + 20
    compare_step(gpu, orc, sc)


CONTACT_CFG = dict(use_rigid_contact_solver=1, rigid_contact_beta=2000.0, rigid_contact_friction=0.4)


def compare_manager(gpu, orc, scene, tol=GRAD_TOL):
    dyn = dynamic_bodies(gpu, scene)
    for R in dyn:
        for RR in dyn:
            for w in range(16):
                e = rel_err(gpu.manager_grad(R, RR, w), orc.manager_grad(R, RR, w))
                assert e <= tol, (R, RR, GRAD_NAMES[w], e)


@pytest.mark.parametrize("manager,gradient_mode", [(1, 1), (1, 0), (0, 1)])
def test_penalty_contact_solver(gpu_factory, oracle_factory, manager, gradient_mode):
    """Penalty rigid-rigid contact + Coulomb friction (RigidContactSolver.cpp:307-345, 419-555) with two dynamic boxes near the
    floor and each other; a reset in the middle replays the reference's history-dependent contact order."""
    sc = scenes.contact_scene(1800)
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc, use_rigid_gradient_manager=manager, gradient_mode=gradient_mode, **CONTACT_CFG)
    v0 = gpu.body_state(1)["v"].copy()
    worst = 0.0
    for s in range(10):
        if s == 6:
            gpu.reset()
            orc.reset()
        gpu.step(1)
        orc.step(1)
        worst = max(worst, compare_step(gpu, orc, sc, grads=not manager))
        if manager:
            compare_manager(gpu, orc, sc)
    assert worst <= 1e-9
    # the contact forces did act: without them the boxes would only feel gravity and the fluid
    assert np.abs(gpu.body_state(1)["omega"] - np.array(sc["bodies"][1]["init_omega"])).max() > 1e-3 and v0 is not None


def test_penalty_contact_order_resort_after_500_steps(gpu_factory, oracle_factory):
    """The contact impulses are applied in the reference's storage order, re-sorted every 500 steps
    (TimeStepDiffDFSPH.cpp:2044-2056); a short fixed time step keeps 505 steps cheap for the CPU oracle."""
    sc = scenes.contact_scene(500)
    gpu, orc = build_pair(gpu_factory, oracle_factory, sc, use_rigid_gradient_manager=1, cfl_method=0, time_step_size=2e-4, target_time=10.0,
                          **CONTACT_CFG)
    gpu.step(499)
    orc.step(499)
    compare_step(gpu, orc, sc, grads=False, state_tol=1e-6)
    for _ in range(6):
        gpu.step(1)
        orc.step(1)
        compare_step(gpu, orc, sc, grads=False, state_tol=1e-6)
    compare_manager(gpu, orc, sc)
