"""Generates tests/golden/*.npz by running the REFERENCE's own code (oracle/_ref/libref.so, built by
oracle/ref/Makefile from /root/reference; only possible in the build container) on small seeded scenes.

Each file stores the complete inputs (config fields, fluid particles, body samples and poses) and the reference's
outputs after every step, so the tests need neither /root/reference nor oracle/_ref at run time:
  python tests/golden/make_golden.py
One ref context can exist per process (the reference keeps its state in singletons), so cases run in subprocesses.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

FLUID_FIELDS = ["position", "velocity", "density", "factor", "kappa", "kappa_v", "density_adv", "acceleration", "sum_grad_p_k"]

CASES = {
    # name: (scene kwargs, config overrides, extras)
    "incomplete_1box": (dict(n_target=1500, n_boxes=1), dict(surface_tension_method=2, surface_tension=0.2, max_error=0.05, target_time=0.05), dict(steps=6)),
    "complete_manager_2box": (dict(n_target=2000, n_boxes=2), dict(surface_tension_method=2, surface_tension=0.2, max_error=0.05, target_time=0.05,
                                                                   gradient_mode=0, use_rigid_gradient_manager=1), dict(steps=6)),
    "ramp_jitter_state": (dict(n_target=1500, n_boxes=1, jitter=0.3, seed=11), dict(surface_tension_method=2, surface_tension=0.5, max_error=0.05,
                                                                                   uniform_acc_rb_time=0.004, target_time=0.012),
                          dict(steps=8, init_v=(0.8, -0.5, 0.1), init_omega=(1.0, 2.0, -0.5), load_state=True)),
    "cfl_iter_nowarm_nogyro": (dict(n_target=1200, n_boxes=1), dict(cfl_method=2, use_pressure_warmstart=0, use_divergence_warmstart=0, rigid_body_mode=1,
                                                                    gradient_mode=2, max_error=0.05, target_time=0.05), dict(steps=6)),
    # boundary viscosity (Viscosity_Standard.cpp:273-318): acceleration of the fluid, reaction force / torque on the body and
    # the dF/dv Jacobian contribution; no shipped scene sets it, the reference's code path is live all the same
    "boundary_viscosity_1box": (dict(n_target=1500, n_boxes=1, jitter=0.2, seed=5), dict(surface_tension_method=2, surface_tension=0.2, max_error=0.05,
                                                                                           target_time=0.05, viscosity=0.02, viscosity_boundary=0.3,
                                                                                           gradient_mode=0), dict(steps=6, init_v=(0.6, -0.4, 0.2),
                                                                                                                  init_omega=(1.0, -2.0, 0.5))),
    # useReleaseRigidBodyMode (billiards-on-water-2balls.json; TimeStepDiffDFSPH.cpp:381-407): body 1 is held until the ramp
    # time has passed and then starts every step with its initial velocities, body 2 moves freely from t = 0
    "release_mode_2box": (dict(n_target=1500, n_boxes=2, jitter=0.2, seed=9), dict(surface_tension_method=2, surface_tension=0.2, max_error=0.05,
                                                                                     uniform_acc_rb_time=0.004, target_time=0.02,
                                                                                     use_release_rigid_body_mode=1, use_rigid_gradient_manager=1),
                          dict(steps=10, init_v=(0.8, -0.5, 0.1), init_omega=(1.0, 2.0, -0.5))),
    # the billiards configuration itself: no ramp time, contact solver and gradient manager on
    "release_mode_contact": (dict(n_target=1500), dict(surface_tension_method=2, surface_tension=0.2, max_error=0.05, target_time=0.05,
                                                       uniform_acc_rb_time=0.0, use_release_rigid_body_mode=1, use_rigid_gradient_manager=1,
                                                       use_rigid_contact_solver=1, rigid_contact_beta=20000.0),
                             # (with init_v = (0.5, -0.3, 0.2) step 5 of this scene sits on a discontinuity of the contact Jacobian: the reference's own
                             # blocks then differ by 16 % between 3 and 16 OpenMP threads; these velocities are clear of it)
                             dict(steps=8, scene="contact", init_v=(0.3, -0.2, 0.1), init_omega=(0.4, 0.8, -0.3))),
    # penalty rigid-rigid contact + friction + manager (BASELINE.json configs[2]); reset() after step 4 exercises the
    # history-dependent contact order of the reference (oracle/oracle_contact.inc)
    "contact_manager_2box": (dict(n_target=1500), dict(surface_tension_method=2, surface_tension=0.3, max_error=0.05, target_time=0.05,
                                                       use_rigid_gradient_manager=1, use_rigid_contact_solver=1, rigid_contact_beta=2000.0,
                                                       rigid_contact_friction=0.4), dict(steps=8, scene="contact", reset_at=4)),
}


def run_case(name):
    from difffr_b200 import scenes
    from difffr_b200.cabi import Context

    rlib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    skw, cfg, extra = CASES[name]
    sc = scenes.contact_scene(**skw) if extra.get("scene") == "contact" else scenes.dam_break_scene(**skw)
    if "init_v" in extra:
        sc["bodies"][1]["init_v"] = extra["init_v"]
        sc["bodies"][1]["init_omega"] = extra["init_omega"]
    ctx = scenes.build_context(lambda **k: Context(lib=rlib, prefix="ref_", **k), sc, **cfg)
    out = {"radius": sc["radius"], "fluid": sc["fluid"], "n_bodies": len(sc["bodies"]), "steps": extra["steps"]}
    out["cfg_keys"] = np.array(list(cfg.keys()))
    out["cfg_vals"] = np.array([float(v) for v in cfg.values()])
    for b, bd in enumerate(sc["bodies"]):
        out[f"body{b}_x_local"] = bd["x_local"]
        out[f"body{b}_dynamic"] = int(bd["dynamic"])
        out[f"body{b}_density"] = bd["density"]
        out[f"body{b}_position"] = np.asarray(bd["position"], dtype=np.float64)
        out[f"body{b}_quat"] = np.asarray(bd["quat"], dtype=np.float64)
        out[f"body{b}_init_v"] = np.asarray(bd.get("init_v", (0, 0, 0)), dtype=np.float64)
        out[f"body{b}_init_omega"] = np.asarray(bd.get("init_omega", (0, 0, 0)), dtype=np.float64)
        out[f"body{b}_volume"] = ctx.body_particles(b, "volume")
    if extra.get("load_state"):
        rng = np.random.default_rng(5)
        n = ctx.num_fluid
        st = dict(x=sc["fluid"], v=rng.normal(scale=0.2, size=(n, 3)), kappa=-1e-6 * rng.random(n), kappa_v=-1e-3 * rng.random(n))
        ctx.load_fluid_state(st["x"], st["v"], st["kappa"], st["kappa_v"])
        for k, v in st.items():
            out["state_" + k] = v
    cnt, idx = ctx.neighbors(-1, -1)
    out["nbr_ff_counts"], out["nbr_ff_indices"] = cnt, idx
    cnt, idx = ctx.neighbors(-1, 1)
    out["nbr_fb1_counts"], out["nbr_fb1_indices"] = cnt, idx
    cnt, idx = ctx.neighbors(1, -1)
    out["nbr_b1f_counts"], out["nbr_b1f_indices"] = cnt, idx
    dyn = [b for b, bd in enumerate(sc["bodies"]) if bd["dynamic"]]
    per_step = {k: [] for k in ("time", "h", "iters", "iters_v", "finished")}
    body_state = {b: [] for b in dyn}
    body_grads = {b: [] for b in dyn}
    body_ft = {b: [] for b in dyn}
    mgr = []
    if "reset_at" in extra:
        out["reset_at"] = extra["reset_at"]
    for s in range(extra["steps"]):
        if s == extra.get("reset_at", -1):
            ctx.reset()
        ctx.step(1)
        info = ctx.step_info()
        per_step["time"].append(info.time)
        per_step["h"].append(info.time_step_size)
        per_step["iters"].append(info.iterations)
        per_step["iters_v"].append(info.iterations_v)
        per_step["finished"].append(info.trajectory_finished)
        for b in dyn:
            st = ctx.body_state(b)
            body_state[b].append(np.concatenate([st["x"], st["q"], st["v"], st["omega"]]))
            g = np.zeros((16, 12))
            for w in range(16):
                a = ctx.body_grad(b, w).ravel()
                g[w, : a.size] = a
            body_grads[b].append(g)
            pr = ctx.body_properties(b)
            body_ft[b].append(np.concatenate([pr["force"], pr["torque"]]))
        if cfg.get("use_rigid_gradient_manager"):
            m = np.zeros((len(dyn), len(dyn), 16, 12))
            for i, R in enumerate(dyn):
                for j, RR in enumerate(dyn):
                    for w in range(16):
                        a = ctx.manager_grad(R, RR, w).ravel()
                        m[i, j, w, : a.size] = a
            mgr.append(m)
        if s in (0, extra["steps"] - 1):
            for f in FLUID_FIELDS:
                out[f"fluid_{f}_step{s + 1}"] = ctx.fluid(f)
    for k, v in per_step.items():
        out["step_" + k] = np.array(v)
    for b in dyn:
        out[f"body{b}_state"] = np.array(body_state[b])
        out[f"body{b}_grads"] = np.array(body_grads[b])
        out[f"body{b}_force_torque"] = np.array(body_ft[b])
    if mgr:
        out["manager_grads"] = np.array(mgr)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", {k: np.asarray(v).shape for k, v in out.items() if k.startswith("step_")})


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for name in CASES:
            subprocess.run([sys.executable, os.path.abspath(__file__), name], check=True, stdout=None)
