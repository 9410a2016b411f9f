"""Golden vectors of a PAPER scene, recorded from the REFERENCE's own code (oracle/_ref/libref.so).

BASELINE.json configs[2] "bottle flipping stage 2": experiments/rigid_body_trajectory_optimization/scene/
diff-bottle-model-collide.json + state/bottle_flip/state_54_particle_Fluid.bgeo (13,312 fluid particles inside a
dynamic bottle of 13,085 boundary samples, penalty rigid-rigid contact solver + gradient manager, 0.25 s velocity ramp).
The scene is parsed and its meshes sampled by this repository's host loader (difffr_b200/host/scene_host.hpp, the
same code the pysplishsplash mirror runs), the resulting arrays drive the reference build, and inputs + outputs go into
tests/golden/paper_bottle_stage2.npz so that the tests need neither /root/reference nor oracle/_ref:
  python tests/golden/make_paper_golden.py            (build container only)
Two segments of 24 steps from the same inputs: "paper" is the scene as shipped (the body is `animated` during the 0.25 s
velocity ramp and carries no Jacobians); "short_ramp" only lowers uniformAccelerateRBTime to 0.004 s so that the free
bottle, with fluid forces, Jacobians, manager blocks and sensitivities, falls inside the window (12 ramp + 12 free steps,
up to 46 pressure / 36 divergence iterations per step).  The window is short on purpose.  This scene amplifies rounding
differences between any two FP-different implementations (the reference with another thread count included): the
manager's omega-sensitivity recurrence grows them ~3x per step during the ramp (1e-18 -> 1e-3 in 40 steps), the free
bottle's state goes from 5e-11 to 1e-5 within 15 steps (net force = small difference of large pressure forces), and
after ~45 steps one borderline convergence decision (15 vs 14 pressure iterations) separates the trajectories for good.
Recorded per step: rigid state, force/torque, iteration counts, time step, all 16 sensitivity blocks and the manager
blocks; the fluid
fields (every 32nd particle) after step 8.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF = "/root/reference/experiments/rigid_body_trajectory_optimization"
# golden name -> (scene file, fluid state file or None, steps, {segment: config overrides}[, steps whose sensitivity
# blocks are compared])
SCENES = {
    # BASELINE.json configs[2]
    "paper_bottle_stage2": ("diff-bottle-model-collide.json", "bottle_flip/state_54_particle_Fluid.bgeo", 24,
                            {"paper": {}, "short_ramp": {"uniform_acc_rb_time": 0.004}}),
    # configs[1]: the bunny floats from the first step (no ramp, per-body chain rule, no manager); lattice start - the
    # settled state_130 the scripts load is not in the reference repository
    "paper_water_rafting": ("diff-water-rafting-bunny.json", None, 5, {"paper": {}}, 2),
    # configs[3]: the high-diving scene (118,389 fluid particles, the duck + four static meshes with 296 k samples, the box
    # emitter firing in the first step, Akinci-2013 surface tension on).  The body samples are rounded to float32 before
    # BOTH sides use them, which halves the fixture (3.4 MB of the file are those samples).
    "paper_high_diving": ("diff-high-diving-duck.json", None, 6, {"paper": {}}, 6),
    # the fifth paper scene (not in BASELINE.json's configs): billiards on water - two dynamic balls, penalty contact solver,
    # gradient manager, useReleaseRigidBodyMode (ball 1 is held for the first steps and then starts EVERY step with its
    # initial velocity, TimeStepDiffDFSPH.cpp:381-407), with the shipped fluid state state/billiards/state_17 (87,374
    # particles, float32 on disk and in the fixture; body samples rounded to float32 as for high diving)
    "paper_billiards": ("billiards-on-water-2balls.json", "billiards/state_17_particle_Fluid.bgeo", 12, {"paper": {}}, 12),
    # (five steps of state, two of sensitivities: the bunny starts inside a perfect lattice whose particles all have
    # rho* = 1 up to rounding, i.e. they sit ON the rho* > 1 gate of the Jacobians (TimeStepDiffDFSPH.cpp:1539).  Which side
    # a particle falls on is rounding noise, so from the third step on the net Jacobians of any two FP-different runs (the
    # reference with another thread count included) differ by 1e-4..1e-3 although their states agree to 1e-12.  The stone
    # of configs[0] only reaches the water after several hundred steps and its settled state is not in the reference
    # repository: no golden for it.)
}
F32_FIXTURE = {"paper_high_diving", "paper_billiards"}  # body samples (and a float32 state file's arrays) stored as float32
FLUID_FIELDS = ["position", "velocity", "kappa", "density_adv"]
FLUID_STRIDE = 32  # recorded for every 32nd particle (fixture size)
FLUID_STEP = 8  # fluid fields are recorded after this step (later the sloshing fluid has amplified rounding differences too far)


def run_segment(name, seg):
    scene_file, state_file, STEPS, SEGMENTS = SCENES[name][:4]
    GRAD_STEPS = SCENES[name][4] if len(SCENES[name]) > 4 else STEPS
    SCENE = os.path.join(REF, "scene", scene_file)
    from pysph_util import import_sph
    from difffr_b200.cabi import Config, Context

    sph = import_sph()
    sc = sph._load_scene_full(SCENE, "")
    if name in F32_FIXTURE:
        for bd in sc["bodies"]:
            bd["samples"] = bd["samples"].astype(np.float32).astype(np.float64)
    st = sph._read_bgeo(os.path.join(REF, "state", state_file)) if state_file else None
    cfg = Config.from_buffer_copy(sc["config"])
    for k, v in SEGMENTS[seg].items():
        setattr(cfg, k, v)
    assert st is None or st["n"] == sc["fluid_x"].shape[0]
    rlib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    ctx = Context(config=cfg, lib=rlib, prefix="ref_")
    ctx.set_fluid(sc["fluid_x"], sc["fluid_v"])
    for b in sc["bodies"]:
        ctx.add_body(b["samples"], bool(b["dynamic"]), float(b["density"]), b["translation"], b["rotation"])
    for e in sc["emitters"]:
        ctx.add_emitter(**e)
    for i, b in enumerate(sc["bodies"]):
        if b["dynamic"]:
            ctx.set_init_v_omega(i, b["init_v"], b["init_omega"])
    ctx.finalize()
    if st is not None:
        ctx.load_fluid_state(st["x"], st["v"], st["kappa"], st["kappa_v"])
    out = {}
    first = (seg == list(SEGMENTS.keys())[0])
    use_mgr = bool(cfg.use_rigid_gradient_manager)
    if first:  # the inputs, stored once
        out.update({"config_bytes": np.frombuffer(sc["config"], dtype=np.uint8), "fluid_x": sc["fluid_x"], "fluid_v": sc["fluid_v"],
                    "n_bodies": len(sc["bodies"]), "steps": STEPS, "grad_steps": GRAD_STEPS, "fluid_step": min(FLUID_STEP, STEPS), "fluid_stride": FLUID_STRIDE, "segments": np.array(list(SEGMENTS.keys()))})
        if st is not None:
            f32 = name in F32_FIXTURE
            for k in ("x", "v", "kappa", "kappa_v"):
                assert not f32 or np.array_equal(st[k].astype(np.float32).astype(np.float64), st[k]), "state file is not float32"
                out["state_" + k] = st[k].astype(np.float32) if f32 else st[k]
        out["n_emitters"] = len(sc["emitters"])
        for k, e in enumerate(sc["emitters"]):
            out[f"emitter{k}_wh"] = np.array([e["width"], e["height"]])
            out[f"emitter{k}_position"] = np.asarray(e["position"], dtype=np.float64)
            out[f"emitter{k}_rotation"] = np.asarray(e["rotation"], dtype=np.float64)
            out[f"emitter{k}_vse"] = np.array([e["velocity"], e["emit_start"], e["emit_end"]], dtype=np.float64)
        for i, b in enumerate(sc["bodies"]):
            out[f"body{i}_samples"] = b["samples"].astype(np.float32) if name in F32_FIXTURE else b["samples"]
            out[f"body{i}_dynamic"] = int(b["dynamic"])
            out[f"body{i}_density"] = float(b["density"])
            out[f"body{i}_translation"] = b["translation"]
            out[f"body{i}_rotation"] = b["rotation"]
            out[f"body{i}_init_v"] = b["init_v"]
            out[f"body{i}_init_omega"] = b["init_omega"]
            out[f"body{i}_volume_sum"] = float(np.sum(ctx.body_particles(i, "volume")))
    dyn = [i for i, b in enumerate(sc["bodies"]) if b["dynamic"]]
    b = dyn[0]
    if first:
        out[f"body{b}_volume"] = ctx.body_particles(b, "volume")
    P = seg + "_"
    out[P + "cfg_keys"] = np.array(list(SEGMENTS[seg].keys()))
    out[P + "cfg_vals"] = np.array([float(v) for v in SEGMENTS[seg].values()])
    rec = {k: [] for k in ("time", "h", "iters", "iters_v", "finished", "num_fluid")}
    states, fts, grads, mgrs = [], [], [], []
    for s in range(STEPS):
        ctx.step(1)
        info = ctx.step_info()
        rec["time"].append(info.time)
        rec["h"].append(info.time_step_size)
        rec["iters"].append(info.iterations)
        rec["iters_v"].append(info.iterations_v)
        rec["finished"].append(info.trajectory_finished)
        rec["num_fluid"].append(info.num_fluid_particles)
        bs = ctx.body_state(b)
        states.append(np.concatenate([bs["x"], bs["q"], bs["v"], bs["omega"]]))
        pr = ctx.body_properties(b)
        fts.append(np.concatenate([pr["force"], pr["torque"]]))
        g = np.zeros((16, 12))
        m = np.zeros((16, 12))
        for w in range(16):
            a = ctx.body_grad(b, w).ravel()
            g[w, : a.size] = a
            if use_mgr:
                a = ctx.manager_grad(b, b, w).ravel()
                m[w, : a.size] = a
        grads.append(g)
        mgrs.append(m)
        if s + 1 == min(FLUID_STEP, STEPS):
            for f in FLUID_FIELDS:
                out[P + "fluid_" + f] = ctx.fluid(f)[::FLUID_STRIDE]
        print(seg, "step", s + 1, "t", info.time, "h", info.time_step_size, "it", info.iterations, info.iterations_v, flush=True)
    for k, v in rec.items():
        out[P + "step_" + k] = np.array(v)
    out[P + "body_state"] = np.array(states)
    out[P + "body_force_torque"] = np.array(fts)
    out[P + "body_grads"] = np.array(grads)
    if use_mgr:
        out[P + "manager_grads"] = np.array(mgrs)
    np.savez_compressed(os.path.join(HERE, f"_paper_seg_{seg}.npz"), **out)


def main():
    import subprocess

    if len(sys.argv) > 2:  # one reference context per process (the reference keeps its state in singletons)
        run_segment(sys.argv[1], sys.argv[2])
        return
    for name in (sys.argv[1:] or list(SCENES)):
        merged = {}
        for seg in SCENES[name][3]:
            subprocess.run([sys.executable, os.path.abspath(__file__), name, seg], check=True)
            part = os.path.join(HERE, f"_paper_seg_{seg}.npz")
            with np.load(part) as z:
                merged.update({k: z[k] for k in z.files})
            os.remove(part)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **merged)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
