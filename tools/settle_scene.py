"""Settles the fluid of a paper scene on the GPU (the scripts' settled state files are not in the reference repository).

The scene is run forward from its lattice start with every dynamic body parked as a static one; every `cycle` steps the
velocities (and warm-start stiffnesses) are cleared, which takes the sloshing energy out quickly, then `free` more steps
run undisturbed.  The positions are kept as float32 in particle-id order (the precision of the reference's own .bgeo
state files).  The result is an INPUT for both sides of the parity tests (loaded with zero velocities).

  gpurun -- python tools/settle_scene.py <arrays.npz> <out.npz> [cycles] [cycle] [free]
<arrays.npz> comes from tests/golden/make_trajectory_golden.py dump <scene> (build container: the scene files live in
/root/reference, which does not exist on the GPU box).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from difffr_b200.cabi import Config, Context  # noqa: E402


def main():
    src, dst = sys.argv[1], sys.argv[2]
    cycles = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    cycle = int(sys.argv[4]) if len(sys.argv) > 4 else 600
    free = int(sys.argv[5]) if len(sys.argv) > 5 else 3000
    z = np.load(src)
    cfg = Config.from_buffer_copy(z["config_bytes"].tobytes())
    cfg.target_time = 1.0e9
    ctx = Context(config=cfg, device=0)
    ctx.set_fluid(z["fluid_x"], np.zeros_like(z["fluid_x"]))
    for i in range(int(z["n_bodies"])):
        ctx.add_body(z[f"body{i}_samples"], False, float(z[f"body{i}_density"]), z[f"body{i}_translation"], z[f"body{i}_rotation"])
    ctx.finalize()
    n = z["fluid_x"].shape[0]
    t0 = time.time()

    def report(tag):
        info = ctx.step_info()
        v = ctx.fluid("velocity")
        x = ctx.fluid("position")
        sp = np.linalg.norm(v, axis=1)
        print(f"{tag}: t {info.time:.3f} steps {info.step_count} h {info.time_step_size:.2e} max|v| {sp.max():.3f} mean|v| {sp.mean():.4f} "
              f"y99 {np.percentile(x[:, 1], 99):.4f} ymax {x[:, 1].max():.4f} wall {time.time() - t0:.0f}s", flush=True)
        return x

    # gentle start: a lattice block may overlap a parked body's samples, and the first pressure solves would shoot those
    # particles out of the tank; clearing the velocities every few steps lets them be pushed out slowly instead
    for c in range(40):
        ctx.step(5)
        x = ctx.fluid("position")
        ctx.load_fluid_state(x, np.zeros_like(x), np.zeros(n), np.zeros(n))
    report("relaxed")
    for c in range(cycles):
        ctx.step(cycle)
        x = report(f"cycle {c}")
        ctx.load_fluid_state(x, np.zeros_like(x), np.zeros(n), np.zeros(n))
    done = 0
    while done < free:
        ctx.step(500)
        done += 500
        x = report(f"free {done}")
    info = ctx.step_info()
    np.savez_compressed(dst, x=x.astype(np.float32), time=info.time, steps=info.step_count,
                        note="float32 positions in particle-id order; settled on the GPU path with the dynamic bodies parked "
                             f"({cycles} cycles of {cycle} steps with velocities cleared in between, then {free} free steps); load with zero velocities")
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
