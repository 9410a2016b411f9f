"""Developer check: run the CUDA path and the CPU oracle side by side on a small scene and print
per-field relative differences after each step (test infrastructure; not part of the product)."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from difffr_b200.cabi import Context, GRAD_NAMES
from difffr_b200 import scenes

def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = np.max(np.abs(a - b)) if a.size else 0.0
    s = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    return d / s

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    nbox = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    olib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    sc = scenes.dam_break_scene(n, n_boxes=nbox)
    kw = dict(surface_tension_method=2, surface_tension=0.2, target_time=0.05, max_error=0.05)
    orc = scenes.build_context(lambda **k: Context(lib=olib, prefix="orc_", **k), sc, **kw)
    gpu = scenes.build_context(lambda **k: Context(**k), sc, **kw)
    print("fluid", gpu.num_fluid, "bodies", [gpu.num_body_particles(b) for b in range(gpu.num_bodies)])
    for b in range(gpu.num_bodies):
        print(" body", b, "volume rel", rel(gpu.body_particles(b, "volume"), orc.body_particles(b, "volume")),
              "pos rel", rel(gpu.body_particles(b, "position"), orc.body_particles(b, "position")))
    for (a, b) in [(-1, -1), (-1, 0), (-1, 1), (1, -1)]:
        cg, ig = gpu.neighbors(a, b); co, io = orc.neighbors(a, b)
        print(" neighbours", (a, b), "counts equal", np.array_equal(cg, co), "indices equal", np.array_equal(ig, io), "total", int(co.sum()))
    for s in range(steps):
        t0 = time.time(); gpu.step(1); tg = time.time() - t0
        t0 = time.time(); orc.step(1); to = time.time() - t0
        ig, io = gpu.step_info(), orc.step_info()
        print(f"step {s}: t {ig.time:.6f}/{io.time:.6f} h {ig.time_step_size:.6e}/{io.time_step_size:.6e} it {ig.iterations}/{io.iterations} itV {ig.iterations_v}/{io.iterations_v}  wall gpu {tg*1e3:.1f} ms cpu {to*1e3:.1f} ms")
        for f in ["position", "velocity", "density", "factor", "kappa", "kappa_v", "density_adv", "acceleration", "sum_grad_p_k"]:
            print(f"   {f:14s} rel {rel(gpu.fluid(f), orc.fluid(f)):.3e}")
        for b in range(1, gpu.num_bodies):
            sg, so = gpu.body_state(b), orc.body_state(b)
            print("   body", b, {k: f"{rel(sg[k], so[k]):.2e}" for k in sg})
            pg, po = gpu.body_properties(b), orc.body_properties(b)
            print("   force", f"{rel(pg['force'], po['force']):.2e}", "torque", f"{rel(pg['torque'], po['torque']):.2e}", po['force'])
            print("   grads", {GRAD_NAMES[w][5:]: f"{rel(gpu.body_grad(b, w), orc.body_grad(b, w)):.1e}" for w in range(16)})
    ms, nl = gpu.device_time_ms()
    print("device ms", ms, "launches", nl)

if __name__ == "__main__":
    main()
