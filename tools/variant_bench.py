"""Tuning helper: per-kernel event timings of one library build (DFR_LIBRARY) on the bench scene."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from difffr_b200 import scenes
from difffr_b200.cabi import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_PARTICLES
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
sc = scenes.dam_break_scene(n, n_boxes=bench.N_BOXES)
ctx = scenes.build_context(lambda **k: Context(device=0, **k), sc, **bench.CFG)
ctx.step(warm)
ms0, _ = ctx.device_time_ms()
i0 = ctx.step_info()
ctx.step(steps)
ms1, _ = ctx.device_time_ms()
i1 = ctx.step_info()
ps = max(i1.total_particle_steps - i0.total_particle_steps, 1)
stats = (f"D {(i1.total_divergence_iterations - i0.total_divergence_iterations) / steps:.2f} "
         f"P {(i1.total_pressure_iterations - i0.total_pressure_iterations) / steps:.2f} "
         f"nbrs {(i1.total_fluid_neighbors - i0.total_fluid_neighbors) / ps:.1f} h {i1.time_step_size:.2e}")
ctx.set_profiling(True)
ctx.step(5)
prof = ctx.kernel_profile()
cls = {}
for k, (ms, cnt) in prof.items():
    if k.endswith("(idle)"):
        continue
    b = k.strip("()").split("<")[0]
    a = cls.setdefault(b, [0.0, 0])
    a[0] += ms; a[1] += cnt
top = sorted(cls.items(), key=lambda kv: -kv[1][0])[:7]
print(os.environ.get("DFR_LIBRARY", "default"), f"ms/step {(ms1-ms0)/steps:.3f} {stats} |", " ".join(f"{k}:{v[0]/v[1]*1e3:.0f}us" for k, v in top))
if os.environ.get("DFR_VARIANT_DETAIL"):
    det = sorted(((k, v) for k, v in prof.items() if not k.endswith("(idle)")), key=lambda kv: -kv[1][0])[:14]
    print("   detail:", " | ".join(f"{k.strip('()').replace('RHO_', '').replace('false', '0').replace('true', '1')}:{v[0]/v[1]*1e3:.0f}us x{v[1]/5:.0f}" for k, v in det))
