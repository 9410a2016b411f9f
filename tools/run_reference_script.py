"""Runs the REFERENCE's own optimisation script, unmodified, against this repository's `pysplishsplash` module.

  stage (build container, /root/reference present):
      python tools/run_reference_script.py stage
          copies experiments/rigid_body_trajectory_optimization/{python/gradient-based-optimize.py, python/utils.py,
          scene/diff-bottle-model-collide.json, models/{UnitBox,bottle}.obj, state/bottle_flip/*} into gpurun_in/refscript/
          - a git-ignored scratch directory that only exists to carry the files to the GPU box (the reference checkout
          does not exist there); delete it after the run.  Nothing of it is committed.
  optimise (GPU box):
      python tools/run_reference_script.py optimise stone|water|billiards [out_dir]
          the README's command for that task (README.md:62-78) with a bounded --maxIter: stone skipping (BASELINE.json
          configs[0]) for one gradient iteration on the regenerated settled state; billiards (README.md:84-87, the
          reference's diff-rigid-contact-multi.py on the shipped state_17) for four iterations; water rafting (configs[1]) with Adam,
          lr 0.1 as in the authors' log (raw_record_and_plot/water_rafting/2023-05-14-dambreak-bunny-ours) for 25 iterations
          after settling its fluid on the GPU.  Loss per iteration and wall time go to reference_script_<task>.json.
  run (GPU box):
      python tools/run_reference_script.py run [out_dir]
          1. BASELINE.json configs[2]: gradient-based-optimize.py --taskType bottle-flip on diff-bottle-model-collide.json +
             state_54 for ONE gradient iteration (--maxIter 0: the script quits after its first optimiser step), with
             PYTHONPATH = tests/standins (quaternion stand-in) : difffr_b200 (pysplishsplash) : the script's directory;
          2. the same trajectory driven directly through the C ABI (difffr_b200.cabi) with the scene parsed by the same
             host loader; loss and gradients are formed the way the script forms them (position_loss / rotation_loss /
             Simulator_layer_v / Simulator_layer_omega with the gradient manager) and compared with what the script
             logged.  Both use the same library, so they must agree to the printed digits.
"""
import json
import os
import re
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "gpurun_in", "refscript")
REF = "/root/reference/experiments/rigid_body_trajectory_optimization"
FILES = ["python/gradient-based-optimize.py", "python/utils.py", "scene/diff-bottle-model-collide.json", "scene/diff-stone-skipping.json",
         "scene/diff-water-rafting-bunny.json", "models/UnitBox.obj", "models/bottle.obj", "models/sphere.obj", "models/bunny-fix.obj",
         "state/bottle_flip/state_54.bin", "state/bottle_flip/state_54_particle_Fluid.bgeo", "state/stone_skipping/state_18.bin",
         "state/water_rafting/state_130.bin", "python/diff-rigid-contact-multi.py", "scene/billiards-on-water-2balls.json",
         "state/billiards/state_17.bin", "state/billiards/state_17_particle_Fluid.bgeo"]
# the README's commands (README.md:62-78) with a bounded number of optimiser iterations
TASKS = {
    "bottle": dict(scene="diff-bottle-model-collide.json", state="bottle_flip/state_54.bin", args=["--taskType", "bottle-flip", "--maxIter", "0"]),
    "stone": dict(scene="diff-stone-skipping.json", state="stone_skipping/state_18.bin",
                  args=["--load-fluid-pos", "--taskType", "stone-skipping", "--maxIter", "0"]),
    "water": dict(scene="diff-water-rafting-bunny.json", state="water_rafting/state_130.bin",
                  args=["--load-fluid-pos-and-vel", "--taskType", "water-rafting", "--optimizer", "adam", "--lr-v", "0.1", "--lr-omega", "0.1",
                        "--patience", "10", "--maxIter", "24"]),
    # README.md:84-87: on-water billiards with the reference's OTHER script (two dynamic balls, contact solver, gradient
    # manager blocks across bodies, useReleaseRigidBodyMode) on the shipped state_17, Adam as coded in the script
    "billiards": dict(script="diff-rigid-contact-multi.py", scene="billiards-on-water-2balls.json", state="billiards/state_17.bin",
                      args=["--load-fluid-pos", "--maxIter", "3"]),
}


def write_state_bgeo(task, x):
    """The particle file the scripts' --state argument implies (state_<n>_particle_Fluid.bgeo beside state_<n>.bin), from
    regenerated settled positions; velocities and warm-start stiffnesses zero.  Written with this repository's partio
    Bgeo writer (difffr_b200/host/state_io.hpp)."""
    sys.path.insert(0, os.path.join(ROOT, "difffr_b200"))
    import pysplishsplash as sph

    stem = os.path.join(STAGE, "state", TASKS[task]["state"])[: -len(".bin")]
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    sph._write_bgeo(stem + "_particle_Fluid.bgeo", x, np.zeros_like(x), np.zeros(n), np.zeros(n))
    print("wrote", stem + "_particle_Fluid.bgeo", n, "particles")


def stage():
    for f in FILES:
        dst = os.path.join(STAGE, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), dst)
    # stone skipping: the settled state of the trajectory golden
    write_state_bgeo("stone", np.load(os.path.join(ROOT, "tests", "golden", "trajectory", "stone_skipping_settled.npz"))["x"])
    # water rafting: its fluid is settled on the GPU box first (tools/settle_scene.py); the scene as arrays travels with the stage
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_trajectory_golden.py"), "dump", "water_rafting",
                    os.path.join(ROOT, "gpurun_in", "water_rafting_scene.npz")], check=True)
    print("staged", len(FILES), "files under", STAGE, "(scratch; remove after the GPU run)")


def numbers(text):
    return [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", text)]


def run_script(task, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    T = TASKS[task]
    script = os.path.join(STAGE, "python", T.get("script", "gradient-based-optimize.py"))
    scene = os.path.join(STAGE, "scene", T["scene"])
    state = os.path.join(STAGE, "state", T["state"])
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "standins"), os.path.join(ROOT, "difffr_b200"), os.path.dirname(script)])
    sim_out = os.path.join(out_dir, "script_output")
    shutil.rmtree(sim_out, ignore_errors=True)
    cmd = [sys.executable, script, "--scene", scene, "--state", state, "--no-gui", "--no-initial-pause", "--stopAt", "100",
           "--output-dir", sim_out] + T["args"]
    import time
    t0 = time.time()
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500)
    wall = time.time() - t0
    open(os.path.join(out_dir, "script_stdout.txt"), "w").write(" ".join(cmd) + "\n" + r.stdout + "\n---- stderr ----\n" + r.stderr)
    log_path = os.path.join(sim_out, "log", "SPH_log.txt")
    log = open(log_path).read() if os.path.exists(log_path) else ""
    log = re.compile(r"\x1b\[[0-9;]*m").sub("", log)
    open(os.path.join(out_dir, "SPH_log.txt"), "w").write(log)
    return r, log, wall, scene


def run_optimisation(task, out_dir):
    """Several optimiser iterations of the unmodified script: the loss per iteration and the wall time."""
    if task == "water":  # settle the fluid around the parked bunny first, then hand the script its particle file
        settled = os.path.join(out_dir, "water_rafting_settled.npz")
        os.makedirs(out_dir, exist_ok=True)
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "settle_scene.py"), os.path.join(ROOT, "gpurun_in", "water_rafting_scene.npz"),
                        settled, "5", "500", "1500"], check=True)
        write_state_bgeo("water", np.load(settled)["x"])
    r, log, wall, _ = run_script(task, out_dir)
    losses = [numbers(m)[0] for m in re.findall(r"\bloss = ([^\n]*)", log)]
    losses = losses[:: 2] if task != "stone" else losses  # two layers (v, omega) log the same loss once each per iteration
    if task == "billiards":  # this script prints its loss ("===== iter k, loss = [tensor([...])] ====") instead of logging it
        losses = [numbers(m)[0] for m in re.findall(r"iter \d+, loss = \[tensor\(\[([^\]]*)\]", r.stdout)]
    steps = [int(numbers(m)[0]) for m in re.findall(r"total timestep of a trajectory = ([^\n]*)", log)]
    grads = re.findall(r"(grad_[a-z_]+ = \[[^\]]*\])", log)
    res = {"task": task, "command_args": TASKS[task]["args"], "script_exit_code": r.returncode, "wall_seconds": wall, "iterations": len(losses),
           "seconds_per_iteration": wall / max(len(losses), 1), "loss_per_iteration": losses, "trajectory_steps": steps[:3],
           "first_iteration_gradients": grads[:4]}
    json.dump(res, open(os.path.join(out_dir, f"reference_script_{task}.json"), "w"), indent=1)
    print(json.dumps(res)[:2500])
    print("REFERENCE_SCRIPT_OK" if r.returncode == 0 and losses and all(np.isfinite(losses)) else "REFERENCE_SCRIPT_PROBLEM")


def run(out_dir):
    r, log, wall, scene = run_script("bottle", out_dir)
    got = {}
    for key in ("loss_x", "loss_rotation", "loss", "grad_init_v_rb", "grad_init_omega_rb"):
        m = re.findall(rf"\b{key} = ([^\n]*(?:\n[^\n=]*)?)", log)
        if m:
            got[key] = numbers(m[0])[: (1 if key.startswith("loss") else 3)]
    print("script exit code", r.returncode, "logged:", got)

    # ---- the same iteration through the C ABI ----
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "difffr_b200"))
    import pysplishsplash as sph
    from difffr_b200.cabi import Config, Context

    sc = sph._load_scene_full(scene, "")
    st = sph._read_bgeo(os.path.join(STAGE, "state", "bottle_flip", "state_54_particle_Fluid.bgeo"))
    state = os.path.join(STAGE, "state", TASKS["bottle"]["state"])
    cfg = Config.from_buffer_copy(sc["config"])
    ctx = Context(config=cfg, device=0)
    ctx.set_fluid(sc["fluid_x"], sc["fluid_v"])
    for b in sc["bodies"]:
        ctx.add_body(b["samples"], bool(b["dynamic"]), float(b["density"]), b["translation"], b["rotation"])
    for i, b in enumerate(sc["bodies"]):
        if b["dynamic"]:
            ctx.set_init_v_omega(i, b["init_v"], b["init_omega"])
    ctx.finalize()
    ctx.set_gradient_mode(1)
    ctx.load_fluid_state(st["x"], st["v"], st["kappa"], st["kappa_v"])
    steps = ctx.run_trajectory(100000)
    body = 1
    s = ctx.body_state(body)
    d = sph._load_scene_summary(scene, "")
    target_x = np.array(d["bodies"][body]["target_x"]) if "bodies" in d and "target_x" in d["bodies"][body] else None
    res = {"steps": int(steps), "x": s["x"].tolist(), "q": s["q"].tolist(), "script": got, "script_exit_code": r.returncode}
    gx_v0, gx_w0 = ctx.manager_grad(body, body, 0), ctx.manager_grad(body, body, 1)
    gq_v0, gq_w0 = ctx.manager_grad(body, body, 2), ctx.manager_grad(body, body, 3)
    res["manager_grad_x_to_v0"] = gx_v0.tolist()
    if target_x is not None:
        gl = s["x"] - target_x
        res["loss_x"] = float(0.5 * np.dot(gl, gl))
        res["grad_init_v_rb"] = (gx_v0.T @ gl).tolist()
    json.dump(res, open(os.path.join(out_dir, "reference_script_run.json"), "w"), indent=1)
    print(json.dumps(res)[:1500])
    ok = True
    if "loss_x" in got and "loss_x" in res:
        rel = abs(got["loss_x"][0] - res["loss_x"]) / max(abs(res["loss_x"]), 1e-300)
        print("loss_x: script", got["loss_x"][0], "C ABI", res["loss_x"], "rel diff", rel)
        ok &= rel < 1e-9
    if "grad_init_v_rb" in got and "grad_init_v_rb" in res:
        a, b_ = np.array(got["grad_init_v_rb"]), np.array(res["grad_init_v_rb"])
        rel = float(np.max(np.abs(a - b_)) / max(np.max(np.abs(b_)), 1e-300))
        print("grad_init_v_rb: script", a, "C ABI", b_, "rel diff", rel)
        ok &= rel < 1e-6  # the log prints ~8 significant digits
    print("REFERENCE_SCRIPT_OK" if ok and r.returncode == 0 else "REFERENCE_SCRIPT_MISMATCH")


if __name__ == "__main__":
    if sys.argv[1] == "stage":
        stage()
    elif sys.argv[1] == "run":
        run(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "refscript"))
    else:  # optimise <stone|water> [out_dir]
        run_optimisation(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "refscript_" + sys.argv[2]))
