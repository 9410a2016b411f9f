"""Replay a golden case (tests/golden/*.npz, produced by the reference's own code — see make_golden.py) through any
binding of the C ABI (CUDA product or CPU oracle) and compare every recorded quantity."""
import glob
import os

import numpy as np

from conftest import rel_err
from difffr_b200.cabi import GRAD_NAMES

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))
               if not os.path.basename(p).startswith("paper_"))
PAPER_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(HERE, "golden", "paper_*.npz")))
FLUID_FIELDS = ["position", "velocity", "density", "factor", "kappa", "kappa_v", "density_adv", "acceleration", "sum_grad_p_k"]
INT_KEYS = {"cfl_method", "min_iterations", "max_iterations", "max_iterations_v", "enable_divergence_solver", "use_pressure_warmstart",
            "use_divergence_warmstart", "viscosity_method", "surface_tension_method", "gradient_mode", "rigid_body_mode", "optimize_rotation",
            "use_rigid_gradient_manager", "use_rigid_contact_solver", "max_emitted_particles", "use_release_rigid_body_mode"}


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def build(factory, g):
    cfg = {}
    for k, v in zip(g["cfg_keys"], g["cfg_vals"]):
        k = str(k)
        cfg[k] = int(v) if k in INT_KEYS else float(v)
    ctx = factory(particle_radius=float(g["radius"]), **cfg)
    ctx.set_fluid(g["fluid"])
    nb = int(g["n_bodies"])
    for b in range(nb):
        ctx.add_body(g[f"body{b}_x_local"], bool(g[f"body{b}_dynamic"]), float(g[f"body{b}_density"]), g[f"body{b}_position"], g[f"body{b}_quat"])
    for b in range(nb):
        if np.any(g[f"body{b}_init_v"] != 0) or np.any(g[f"body{b}_init_omega"] != 0):
            ctx.set_init_v_omega(b, g[f"body{b}_init_v"], g[f"body{b}_init_omega"])
    ctx.finalize()
    return ctx, cfg


def replay_and_compare(factory, name, state_tol, grad_tol):
    """Returns the worst relative error seen.  Neighbour sets and iteration counts must match exactly."""
    g = load(name)
    ctx, cfg = build(factory, g)
    nb = int(g["n_bodies"])
    dyn = [b for b in range(nb) if int(g[f"body{b}_dynamic"])]
    worst = 0.0
    for b in range(nb):
        e = rel_err(ctx.body_particles(b, "volume"), g[f"body{b}_volume"])
        assert e <= state_tol, ("boundary volume", b, e)
        worst = max(worst, e)
    if "state_x" in g.files:
        ctx.load_fluid_state(g["state_x"], g["state_v"], g["state_kappa"], g["state_kappa_v"])
    for key, (a, b) in {"ff": (-1, -1), "fb1": (-1, 1), "b1f": (1, -1)}.items():
        cnt, idx = ctx.neighbors(a, b)
        assert np.array_equal(cnt, g[f"nbr_{key}_counts"]), ("neighbour counts", key)
        assert np.array_equal(idx, g[f"nbr_{key}_indices"]), ("neighbour sets", key)
    steps = int(g["steps"])
    reset_at = int(g["reset_at"]) if "reset_at" in g.files else -1
    for s in range(steps):
        if s == reset_at:
            ctx.reset()
        ctx.step(1)
        info = ctx.step_info()
        assert info.iterations == int(g["step_iters"][s]) and info.iterations_v == int(g["step_iters_v"][s]), (s, info.iterations, info.iterations_v)
        assert abs(info.time - g["step_time"][s]) <= 1e-12 * max(abs(g["step_time"][s]), 1e-30)
        assert abs(info.time_step_size - g["step_h"][s]) <= 1e-12 * g["step_h"][s]
        assert info.trajectory_finished == int(g["step_finished"][s])
        for b in dyn:
            st = ctx.body_state(b)
            got = np.concatenate([st["x"], st["q"], st["v"], st["omega"]])
            ref = g[f"body{b}_state"][s]
            for sl, nm in ((slice(0, 3), "x"), (slice(3, 7), "q"), (slice(7, 10), "v"), (slice(10, 13), "omega")):
                e = rel_err(got[sl], ref[sl])
                assert e <= state_tol, (s, b, nm, e)
                worst = max(worst, e)
            pr = ctx.body_properties(b)
            ft = g[f"body{b}_force_torque"][s]
            scale = max(np.max(np.abs(ft)), 1e-300)
            e = max(np.max(np.abs(pr["force"] - ft[:3])), np.max(np.abs(pr["torque"] - ft[3:]))) / scale
            assert e <= state_tol, (s, b, "force/torque", e)
            for w in range(16):
                a = ctx.body_grad(b, w).ravel()
                e = rel_err(a, g[f"body{b}_grads"][s, w, : a.size])
                assert e <= grad_tol, (s, b, GRAD_NAMES[w], e)
                worst = max(worst, e)
        if "manager_grads" in g.files:
            for i, R in enumerate(dyn):
                for j, RR in enumerate(dyn):
                    for w in range(16):
                        a = ctx.manager_grad(R, RR, w).ravel()
                        e = rel_err(a, g["manager_grads"][s, i, j, w, : a.size])
                        assert e <= grad_tol, (s, R, RR, GRAD_NAMES[w], e)
                        worst = max(worst, e)
        if s in (0, steps - 1):
            for f in FLUID_FIELDS:
                e = rel_err(ctx.fluid(f), g[f"fluid_{f}_step{s + 1}"])
                assert e <= state_tol, (s, f, e)
                worst = max(worst, e)
    return worst


def replay_paper_and_compare(factory, name, state_tol, grad_tol):
    """Replay a paper-scene golden (tests/golden/make_paper_golden.py: complete dfr_config, arrays from the host scene
    loader, the shipped fluid state; one recorded segment per configuration variant) and compare what the reference
    recorded.  Returns the worst relative error.  Iteration counts must match exactly."""
    from difffr_b200.cabi import Config

    g = load(name)
    worst = 0.0
    for seg in [str(x) for x in g["segments"]]:
        P = seg + "_"
        cfg = Config.from_buffer_copy(g["config_bytes"].tobytes())
        for k, v in zip(g[P + "cfg_keys"], g[P + "cfg_vals"]):
            setattr(cfg, str(k), int(v) if str(k) in INT_KEYS else float(v))
        ctx = factory(config=cfg)
        ctx.set_fluid(g["fluid_x"], g["fluid_v"])
        nb = int(g["n_bodies"])
        for b in range(nb):
            ctx.add_body(g[f"body{b}_samples"], bool(g[f"body{b}_dynamic"]), float(g[f"body{b}_density"]), g[f"body{b}_translation"],
                         g[f"body{b}_rotation"])
        dyn = [b for b in range(nb) if int(g[f"body{b}_dynamic"])]
        for k in range(int(g["n_emitters"]) if "n_emitters" in g.files else 0):
            wh, vse = g[f"emitter{k}_wh"], g[f"emitter{k}_vse"]
            ctx.add_emitter(width=int(wh[0]), height=int(wh[1]), position=g[f"emitter{k}_position"], rotation=g[f"emitter{k}_rotation"],
                            velocity=float(vse[0]), emit_start=float(vse[1]), emit_end=float(vse[2]))
        for b in dyn:
            ctx.set_init_v_omega(b, g[f"body{b}_init_v"], g[f"body{b}_init_omega"])
        ctx.finalize()
        for b in range(nb):
            e = abs(float(np.sum(ctx.body_particles(b, "volume"))) - float(g[f"body{b}_volume_sum"])) / float(g[f"body{b}_volume_sum"])
            assert e <= state_tol, ("boundary volume sum", b, e)
        e = rel_err(ctx.body_particles(dyn[0], "volume"), g[f"body{dyn[0]}_volume"])
        assert e <= state_tol, ("boundary volume", e)
        if "state_x" in g.files:
            ctx.load_fluid_state(g["state_x"], g["state_v"], g["state_kappa"], g["state_kappa_v"])
        b = dyn[0]
        for s in range(int(g["steps"])):
            ctx.step(1)
            info = ctx.step_info()
            assert info.iterations == int(g[P + "step_iters"][s]) and info.iterations_v == int(g[P + "step_iters_v"][s]), (
                seg, s, info.iterations, info.iterations_v)
            assert abs(info.time - g[P + "step_time"][s]) <= 1e-12 * max(abs(g[P + "step_time"][s]), 1e-30)
            assert abs(info.time_step_size - g[P + "step_h"][s]) <= 1e-10 * g[P + "step_h"][s], (seg, s)
            assert info.trajectory_finished == int(g[P + "step_finished"][s])
            if P + "step_num_fluid" in g.files:  # emitter scenes: emitted particles
                assert info.num_fluid_particles == int(g[P + "step_num_fluid"][s]), (seg, s, info.num_fluid_particles)
            st = ctx.body_state(b)
            got = np.concatenate([st["x"], st["q"], st["v"], st["omega"]])
            ref = g[P + "body_state"][s]
            for sl, nm in ((slice(0, 3), "x"), (slice(3, 7), "q"), (slice(7, 10), "v"), (slice(10, 13), "omega")):
                e = rel_err(got[sl], ref[sl])
                assert e <= state_tol, (seg, s, nm, e)
                worst = max(worst, e)
            pr = ctx.body_properties(b)
            ft = g[P + "body_force_torque"][s]
            scale = max(np.max(np.abs(ft)), 1e-300)
            e = max(np.max(np.abs(pr["force"] - ft[:3])), np.max(np.abs(pr["torque"] - ft[3:]))) / scale
            assert e <= state_tol, (seg, s, "force/torque", e)
            worst = max(worst, e)
            if s + 1 == int(g["fluid_step"]):
                for f in ("position", "velocity", "kappa", "density_adv"):
                    got, ref = ctx.fluid(f)[:: int(g["fluid_stride"])], g[P + "fluid_" + f]
                    # the shipped state file holds 22 particles with NaN positions; the reference carries them along
                    assert np.array_equal(np.isnan(got), np.isnan(ref)), (seg, f, "NaN pattern")
                    ok = ~np.isnan(ref)
                    e = rel_err(got[ok], ref[ok])
                    assert e <= state_tol, (seg, f, e)
                    worst = max(worst, e)
            for w in range(16 if s < int(g["grad_steps"]) else 0):  # (make_paper_golden.py: where Jacobians are well conditioned)
                a = ctx.body_grad(b, w).ravel()
                e = rel_err(a, g[P + "body_grads"][s, w, : a.size])
                assert e <= grad_tol, (seg, s, GRAD_NAMES[w], e)
                worst = max(worst, e)
                if P + "manager_grads" in g.files:
                    a = ctx.manager_grad(b, b, w).ravel()
                    e = rel_err(a, g[P + "manager_grads"][s, w, : a.size])
                    assert e <= grad_tol, (seg, s, "manager", GRAD_NAMES[w], e)
                    worst = max(worst, e)
        ctx.close()
    return worst
