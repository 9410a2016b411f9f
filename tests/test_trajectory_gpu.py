"""Whole-trajectory parity on a PAPER scene: BASELINE.json configs[0], stone skipping.

diff-stone-skipping.json with the settled fluid of tests/golden/trajectory/stone_skipping_settled.npz (237,699 particles;
made by tools/settle_scene.py, see tests/golden/make_trajectory_golden.py) is run from the throw to the end of the trajectory
(0.05 s velocity ramp + 0.18 s target time, ~2,000 CFL-limited steps) through the C ABI on the GPU and compared with the
record the REFERENCE's own code (oracle/_ref) left for the same inputs: rigid state, time step and iteration counts after
every step, the 16 Jacobian / sensitivity blocks every 25 steps and at the end.

What is asserted:
  * while the two runs are the same trajectory (identical iteration counts so far) the per-step rigid state agrees within
    the north-star's 1e-6 and the sensitivities within 1e-4;
  * the end-of-trajectory state and loss gradient agree within END_TOL, and the measured numbers are written to
    gpurun_out/trajectory_stone_skipping.json (profiles/ keeps a copy) - see the note on END_TOL below.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, rel_err
from difffr_b200.cabi import Config

pytestmark = pytest.mark.gpu

TRAJ = os.path.join(ROOT, "tests", "golden", "trajectory")
STATE_TOL = 1e-6
GRAD_TOL = 1e-4
# End of trajectory: the stone hits the water at 30 m/s and the solver takes up to tens of iterations per step; one
# borderline convergence decision (n vs n+1 iterations) anywhere in ~2,000 steps separates two FP-different runs for good,
# after which they are two valid samples of a chaotic splash.  The bound below is what the reference's own code shows
# against the oracle port on the CPU (make_trajectory_golden.py record stone_skipping orc; numbers in DESIGN.md).
END_STATE_TOL = 5e-2
END_GRAD_TOL = 5e-1


def build_gpu(gpu_factory, g, x0):
    cfg = Config.from_buffer_copy(g["config_bytes"].tobytes())
    ctx = gpu_factory(config=cfg)
    ctx.set_fluid(x0, np.zeros_like(x0))
    nb = int(g["n_bodies"])
    for i in range(nb):
        ctx.add_body(g[f"body{i}_samples"], bool(g[f"body{i}_dynamic"]), float(g[f"body{i}_density"]), g[f"body{i}_translation"],
                     g[f"body{i}_rotation"])
    for i in range(nb):
        if int(g[f"body{i}_dynamic"]):
            ctx.set_init_v_omega(i, g[f"body{i}_init_v"], g[f"body{i}_init_omega"])
    ctx.finalize()
    ctx.load_fluid_state(x0, np.zeros_like(x0), None, None)
    return ctx


def loss_gradient(state, grads, target_x):
    """d(0.5 |x_end - target|^2)/d(v0, omega0) as gradient-based-optimize.py forms it (Simulator_layer_shared, :290-300)."""
    gx = state[:3] - target_x
    gx_v0 = grads[0, :9].reshape(3, 3)
    gx_w0 = grads[1, :9].reshape(3, 3)
    return np.concatenate([gx_v0.T @ gx, gx_w0.T @ gx])


def test_stone_skipping_whole_trajectory(gpu_factory):
    path = os.path.join(TRAJ, "traj_stone_skipping.npz")
    if not os.path.exists(path):
        pytest.skip("trajectory record not generated")
    g = np.load(path)
    x0 = np.load(os.path.join(TRAJ, "stone_skipping_settled.npz"))["x"].astype(np.float64)
    assert x0.shape[0] == int(g["n_fluid"])
    ctx = build_gpu(gpu_factory, g, x0)
    b = int(g["dyn_body"])
    ref_state, ref_h = g["body_state"], g["step_h"]
    ref_it, ref_itv = g["step_iters"], g["step_iters_v"]
    grad_steps = {int(s): k for k, s in enumerate(g["grad_steps"])}
    n_ref = ref_state.shape[0]
    same = True          # identical iteration counts so far
    split_step = None
    worst_state_same, worst_grad_same = 0.0, 0.0
    first_violation = None
    first_grad_violation = None
    err_curve = []
    s = 0
    last_grads = None
    while s < n_ref + 50:
        ctx.step(1)
        info = ctx.step_info()
        st = ctx.body_state(b)
        got = np.concatenate([st["x"], st["q"], st["v"], st["omega"]])
        if s < n_ref:
            comp = [rel_err(got[sl], ref_state[s][sl]) for sl in (slice(0, 3), slice(3, 7), slice(7, 10), slice(10, 13))]
            e = max(comp)
            if same and (info.iterations != int(ref_it[s]) or info.iterations_v != int(ref_itv[s])):
                same, split_step = False, s + 1
            if same:
                if e > STATE_TOL and first_violation is None:
                    first_violation = (s + 1, comp, abs(info.time_step_size - ref_h[s]) / ref_h[s])
                worst_state_same = max(worst_state_same, e)
            if (s + 1) % 10 == 0 or (100 <= s + 1 <= 160):
                err_curve.append((s + 1, comp, abs(info.time_step_size - ref_h[s]) / ref_h[s], info.iterations, int(ref_it[s])))
            if (s + 1) in grad_steps:
                gg = np.zeros((16, 12))
                for w in range(16):
                    a = ctx.body_grad(b, w).ravel()
                    gg[w, : a.size] = a
                last_grads = gg
                if same:
                    rg = g["body_grads"][grad_steps[s + 1]]
                    for w in range(16):
                        eg = rel_err(gg[w], rg[w])
                        if eg > GRAD_TOL and first_grad_violation is None:
                            first_grad_violation = (s + 1, w, eg)
                        worst_grad_same = max(worst_grad_same, eg)
        s += 1
        if info.trajectory_finished:
            break
    assert info.trajectory_finished
    gg = np.zeros((16, 12))
    for w in range(16):
        a = ctx.body_grad(b, w).ravel()
        gg[w, : a.size] = a
    end_ref_state, end_ref_grads = ref_state[-1], g["body_grads"][-1]
    target = np.array([1.7, 1.6, 0.0])  # targetX of the stone in diff-stone-skipping.json
    lg, lr = loss_gradient(got, gg, target), loss_gradient(end_ref_state, end_ref_grads, target)
    out = {
        "scene": "diff-stone-skipping.json + settled fluid (tests/golden/trajectory/stone_skipping_settled.npz)",
        "steps_gpu": s, "steps_reference": int(n_ref), "first_step_with_different_iteration_counts": split_step,
        "worst_state_rel_err_while_same_trajectory": worst_state_same, "worst_sensitivity_rel_err_while_same_trajectory": worst_grad_same,
        "end_state_rel_err": {"x": rel_err(got[:3], end_ref_state[:3]), "q": rel_err(got[3:7], end_ref_state[3:7]),
                              "v": rel_err(got[7:10], end_ref_state[7:10]), "omega": rel_err(got[10:13], end_ref_state[10:13])},
        "end_x_gpu": got[:3].tolist(), "end_x_reference": end_ref_state[:3].tolist(),
        "end_loss_gradient_gpu": lg.tolist(), "end_loss_gradient_reference": lr.tolist(), "end_loss_gradient_rel_err": rel_err(lg, lr),
        "end_sensitivity_rel_err": {f"block{w}": rel_err(gg[w], end_ref_grads[w]) for w in range(8)},
        "first_state_violation_while_same_trajectory": first_violation, "first_sensitivity_violation_while_same_trajectory": first_grad_violation,
        "state_err_curve": err_curve,
    }
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "trajectory_stone_skipping.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: out[k] for k in ("steps_gpu", "steps_reference", "first_step_with_different_iteration_counts",
                                          "worst_state_rel_err_while_same_trajectory", "end_state_rel_err", "end_loss_gradient_rel_err")}))
    assert first_violation is None, first_violation
    assert first_grad_violation is None, first_grad_violation
    assert abs(s - n_ref) <= max(3, n_ref // 100)
    assert out["end_state_rel_err"]["x"] <= END_STATE_TOL
    assert out["end_loss_gradient_rel_err"] <= END_GRAD_TOL
    assert split_step is None or split_step > 100  # the ramp and the entry into the water are the same trajectory
