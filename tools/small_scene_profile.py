"""Per-kernel time of one step at a given scene size, next to the replayed (CUDA-graph) step time: where the time of a
paper-scale scene goes.  gpurun -- python tools/small_scene_profile.py <n_particles> [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from difffr_b200 import scenes  # noqa: E402
from difffr_b200.cabi import Context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 238000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
sc = scenes.dam_break_scene(n, n_boxes=bench.N_BOXES)
cfg = dict(bench.CFG, cfl_method=0, time_step_size=1.0e-3)
ctx = scenes.build_context(lambda **k: Context(device=0, **k), sc, **cfg)
ctx.step(20)
for label, env in (("graph", None), ("stream", "1")):
    if env:
        os.environ["DFR_NO_GRAPH"] = env
    ms0, l0 = ctx.device_time_ms()
    i0 = ctx.step_info()
    ctx.step(steps)
    ms1, l1 = ctx.device_time_ms()
    i1 = ctx.step_info()
    print(f"{label}: {ctx.num_fluid} particles {(ms1 - ms0) / steps:.4f} ms/step, {(l1 - l0) / steps:.1f} launches/step, "
          f"D {(i1.total_divergence_iterations - i0.total_divergence_iterations) / steps:.2f} "
          f"P {(i1.total_pressure_iterations - i0.total_pressure_iterations) / steps:.2f}")
ctx.set_profiling(True)
ctx.step(10)
prof = ctx.kernel_profile()
rows = sorted(((k, v) for k, v in prof.items()), key=lambda kv: -kv[1][0])
tot = sum(v[0] for _, v in rows)
print(f"profiled (stream path, one event pair per launch): {tot / 10:.4f} ms/step of kernel time")
for k, (ms, cnt) in rows:
    print(f"  {k.strip('()')[:70]:70s} {ms / 10 * 1e3:8.1f} us/step  x{cnt / 10:.1f}  avg {ms / cnt * 1e3:7.1f} us")
