"""Whole-trajectory golden vectors of a PAPER scene, recorded from the REFERENCE's own code (oracle/_ref).

BASELINE.json configs[0] "stone skipping": experiments/rigid_body_trajectory_optimization/scene/diff-stone-skipping.json
(237,699 fluid particles, r = 0.015, a 0.2 x 0.05 x 0.2 ellipsoid thrown at 25 m/s over a 0.05 s velocity ramp,
targetTime 0.18 s).  The scripts load a settled fluid state whose particle file is not in the reference repository
(state/stone_skipping/state_18.bin is there, state_18_particle_Fluid.bgeo is not), so - SURVEY.md §8c - the state is
regenerated: the reference build runs the scene forward from the lattice with the stone parked as a static body until
the column has collapsed to its rest height, and the positions are kept as float32 (the precision of the reference's own
.bgeo state files).  Both sides then start from that file with zero velocities (--load-fluid-pos semantics).

  python tests/golden/make_trajectory_golden.py settle stone_skipping     -> tests/golden/trajectory/stone_skipping_settled.npz
      (CPU, reference build; slow: the column sloshes for seconds of simulated time.  The committed file was made on the
      GPU instead: `dump stone_skipping gpurun_in/stone.npz`, then tools/settle_scene.py under gpurun)
  python tests/golden/make_trajectory_golden.py record stone_skipping     -> tests/golden/trajectory/traj_stone_skipping.npz
  python tests/golden/make_trajectory_golden.py record stone_skipping orc -> /tmp/traj_stone_skipping_orc.npz (the oracle
      port on the same inputs: how far two FP-different CPU implementations drift over the trajectory)
  python tests/golden/make_trajectory_golden.py drift stone_skipping      -> tests/golden/trajectory/stone_skipping_cpu_drift.npz

The record holds the complete inputs except the fluid positions (the settled file), per step the rigid state, time step
and iteration counts, and every GRAD_EVERY steps plus at the end of the trajectory the 16 Jacobian / sensitivity blocks.
Only possible in the build container (needs /root/reference and oracle/_ref).
"""
import ctypes
import os
import sys
import time

import numpy as np

GOLDEN = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.join(GOLDEN, "trajectory")  # kept apart from the per-step goldens, which the tests find by globbing *.npz
ROOT = os.path.dirname(os.path.dirname(GOLDEN))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF = "/root/reference/experiments/rigid_body_trajectory_optimization"
SCENES = {
    # name: (scene file, settle time [s], max steps of the record)
    "stone_skipping": ("diff-stone-skipping.json", 1.6, 4000),
    # BASELINE.json configs[1]; its fluid was settled on the GPU around the parked bunny (tools/run_reference_script.py)
    "water_rafting": ("diff-water-rafting-bunny.json", 1.6, 8000),
}
GRAD_EVERY = 25


def load(name):
    from pysph_util import import_sph
    from difffr_b200.cabi import Config

    sph = import_sph()
    sc = sph._load_scene_full(os.path.join(REF, "scene", SCENES[name][0]), "")
    return sc, Config.from_buffer_copy(sc["config"])


def make_ctx(sc, cfg, lib_kind, park_dynamic=False):
    from difffr_b200.cabi import Context

    if lib_kind == "orc":
        lib, prefix = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle_fast.so")), "orc_"
    else:
        lib, prefix = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "fast", "libref.so")), "ref_"
    ctx = Context(config=cfg, lib=lib, prefix=prefix)
    ctx.set_fluid(sc["fluid_x"], sc["fluid_v"])
    for b in sc["bodies"]:
        ctx.add_body(b["samples"], bool(b["dynamic"]) and not park_dynamic, float(b["density"]), b["translation"], b["rotation"])
    if not park_dynamic:
        for i, b in enumerate(sc["bodies"]):
            if b["dynamic"]:
                ctx.set_init_v_omega(i, b["init_v"], b["init_omega"])
    ctx.finalize()
    return ctx


def settle(name):
    sc, cfg = load(name)
    cfg.target_time = 1.0e9
    ctx = make_ctx(sc, cfg, "ref", park_dynamic=True)
    t_end = SCENES[name][1]
    t0 = time.time()
    while True:
        ctx.step(10)
        info = ctx.step_info()
        v = ctx.fluid("velocity")
        print(f"t {info.time:.4f} h {info.time_step_size:.2e} steps {info.step_count} max|v| {np.abs(v).max():.3f} "
              f"mean|v| {np.linalg.norm(v, axis=1).mean():.4f} wall {time.time() - t0:.0f}s", flush=True)
        if info.time >= t_end:
            break
    x = ctx.fluid("position").astype(np.float32)
    path = os.path.join(HERE, name + "_settled.npz")
    np.savez_compressed(path, x=x, time=info.time, steps=info.step_count,
                        note="float32 positions in particle-id order after the reference build ran the scene from the lattice "
                             "with the dynamic body parked; load with zero velocities")
    print("wrote", path, os.path.getsize(path), "bytes; y range", x[:, 1].min(), x[:, 1].max())


def record(name, lib_kind="ref"):
    sc, cfg = load(name)
    st = np.load(os.path.join(HERE, name + "_settled.npz"))
    x0 = st["x"].astype(np.float64)
    assert x0.shape == sc["fluid_x"].shape
    ctx = make_ctx(sc, cfg, lib_kind)
    ctx.load_fluid_state(x0, np.zeros_like(x0), None, None)
    dyn = [i for i, b in enumerate(sc["bodies"]) if b["dynamic"]]
    b = dyn[0]
    from pysph_util import import_sph
    summary = import_sph()._load_scene_summary(os.path.join(REF, "scene", SCENES[name][0]), "")
    out = {"config_bytes": np.frombuffer(sc["config"], dtype=np.uint8), "n_bodies": len(sc["bodies"]), "dyn_body": b, "grad_every": GRAD_EVERY,
           "n_fluid": x0.shape[0], "target_x": np.asarray(summary["bodies"][b]["target_x"], dtype=np.float64)}
    for i, bd in enumerate(sc["bodies"]):
        out[f"body{i}_samples"] = bd["samples"]
        out[f"body{i}_dynamic"] = int(bd["dynamic"])
        out[f"body{i}_density"] = float(bd["density"])
        out[f"body{i}_translation"] = bd["translation"]
        out[f"body{i}_rotation"] = bd["rotation"]
        out[f"body{i}_init_v"] = bd["init_v"]
        out[f"body{i}_init_omega"] = bd["init_omega"]
    rec = {k: [] for k in ("time", "h", "iters", "iters_v")}
    states, fts, grad_steps, grads = [], [], [], []
    t0 = time.time()
    for s in range(SCENES[name][2]):
        ctx.step(1)
        info = ctx.step_info()
        rec["time"].append(info.time)
        rec["h"].append(info.time_step_size)
        rec["iters"].append(info.iterations)
        rec["iters_v"].append(info.iterations_v)
        bs = ctx.body_state(b)
        states.append(np.concatenate([bs["x"], bs["q"], bs["v"], bs["omega"]]))
        pr = ctx.body_properties(b)
        fts.append(np.concatenate([pr["force"], pr["torque"]]))
        last = bool(info.trajectory_finished)
        if (s + 1) % GRAD_EVERY == 0 or last:
            g = np.zeros((16, 12))
            for w in range(16):
                a = ctx.body_grad(b, w).ravel()
                g[w, : a.size] = a
            grad_steps.append(s + 1)
            grads.append(g)
        if (s + 1) % 25 == 0 or last:
            print(f"{lib_kind} step {s + 1} t {info.time:.5f} h {info.time_step_size:.2e} it {info.iterations} {info.iterations_v} "
                  f"x {bs['x']} wall {time.time() - t0:.0f}s", flush=True)
        if last:
            break
    for k, v in rec.items():
        out["step_" + k] = np.array(v)
    out["body_state"] = np.array(states)
    out["body_force_torque"] = np.array(fts)
    out["grad_steps"] = np.array(grad_steps)
    out["body_grads"] = np.array(grads)
    path = os.path.join(HERE, f"traj_{name}.npz") if lib_kind == "ref" else f"/tmp/traj_{name}_{lib_kind}.npz"
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(states), "steps")


def drift(name):
    """tests/golden/trajectory/<name>_cpu_drift.npz: the oracle port's record against the reference's, step by step."""
    def rel(a, b):
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    r, o = np.load(os.path.join(HERE, f"traj_{name}.npz")), np.load(f"/tmp/traj_{name}_orc.npz")
    n = min(len(r["body_state"]), len(o["body_state"]))
    sl4 = (slice(0, 3), slice(3, 7), slice(7, 10), slice(10, 13))
    err = np.array([[rel(o["body_state"][s][sl], r["body_state"][s][sl]) for sl in sl4] for s in range(n)])
    dh = np.abs(o["step_h"][:n] - r["step_h"][:n]) / r["step_h"][:n]
    tgt = r["target_x"] if "target_x" in r.files else np.array([1.7, 1.6, 0.0])

    def lg(st, g):
        gx = st[:3] - tgt
        return np.concatenate([g[0, :9].reshape(3, 3).T @ gx, g[1, :9].reshape(3, 3).T @ gx])

    eo, er, go, gr = o["body_state"][-1], r["body_state"][-1], o["body_grads"][-1], r["body_grads"][-1]
    os_, rs_ = [int(v) for v in o["grad_steps"]], [int(v) for v in r["grad_steps"]]
    gs = [s for s in rs_ if s <= n and s in os_]
    gerr = [[rel(o["body_grads"][os_.index(s)][w], r["body_grads"][rs_.index(s)][w]) for w in range(16)] for s in gs]
    np.savez_compressed(os.path.join(HERE, f"{name}_cpu_drift.npz"), steps_oracle=len(o["body_state"]), steps_reference=len(r["body_state"]),
                        state_err=err, h_err=dh, grad_steps=np.array(gs), grad_err=np.array(gerr),
                        end_state_err=np.array([rel(eo[sl], er[sl]) for sl in sl4]), end_loss_gradient_err=rel(lg(eo, go), lg(er, gr)),
                        end_loss_gradient_oracle=lg(eo, go), end_loss_gradient_reference=lg(er, gr),
                        end_sensitivity_err=np.array([rel(go[w], gr[w]) for w in range(16)]),
                        iteration_mismatch_steps=int(np.sum((o["step_iters"][:n] != r["step_iters"][:n]) | (o["step_iters_v"][:n] != r["step_iters_v"][:n]))),
                        note="the CPU oracle port against the reference's own code on the same trajectory: how far two FP64 "
                             "implementations of the same algorithm drift apart")


def dump(name, path):
    """The scene as arrays (what tools/settle_scene.py needs on the GPU box, where /root/reference does not exist)."""
    sc, cfg = load(name)
    out = {"config_bytes": np.frombuffer(sc["config"], dtype=np.uint8), "fluid_x": sc["fluid_x"], "n_bodies": len(sc["bodies"])}
    for i, bd in enumerate(sc["bodies"]):
        out[f"body{i}_samples"] = bd["samples"]
        out[f"body{i}_density"] = float(bd["density"])
        out[f"body{i}_translation"] = bd["translation"]
        out[f"body{i}_rotation"] = bd["rotation"]
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    cmd, name = sys.argv[1], sys.argv[2]
    if cmd == "dump":
        dump(name, sys.argv[3])
    elif cmd == "drift":
        drift(name)
    elif cmd == "settle":
        settle(name)
    else:
        record(name, sys.argv[3] if len(sys.argv) > 3 else "ref")
