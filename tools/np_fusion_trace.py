import os, sys
sys.path.insert(0, "/root/repo")
import bench
from difffr_b200 import scenes
from difffr_b200.cabi import Context
n = int(sys.argv[1])
sc = scenes.dam_break_scene(n, n_boxes=bench.N_BOXES)
ctx = scenes.build_context(lambda **k: Context(device=0, **k), sc, **bench.CFG)
ctx.step(int(sys.argv[2]))
ctx.set_profiling(True)
prev = {}
i0 = ctx.step_info()
for s in range(12):
    ctx.step(1)
    i1 = ctx.step_info()
    prof = ctx.kernel_profile()
    cur = {k: v[1] for k, v in prof.items()}
    d = {k: cur[k] - prev.get(k, 0) for k in cur if cur[k] - prev.get(k, 0)}
    prev = cur
    print("D", i1.iterations_v, "P", i1.iterations, {k.replace("RHO_", "").replace("false", "0"): v for k, v in d.items() if "NONPRESSURE" in k or "nonpressure" in k or "apply" in k})
