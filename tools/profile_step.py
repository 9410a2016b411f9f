"""Profiling driver: build the bench scene and run a few steps (for `ncu` launch lists / --set full captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from difffr_b200 import scenes
from difffr_b200.cabi import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_PARTICLES
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = scenes.dam_break_scene(n, n_boxes=bench.N_BOXES)
ctx = scenes.build_context(lambda **k: Context(device=0, **k), sc, **bench.CFG)
ctx.step(steps)
info = ctx.step_info()
ms, launches = ctx.device_time_ms()
print(f"particles {ctx.num_fluid} steps {steps} device_ms {ms:.2f} launches {launches} P {info.total_pressure_iterations/steps:.2f} D {info.total_divergence_iterations/steps:.2f}")
