"""A/B of the side-branch build of the dynamic-boundary grid in recorded steps (DFR_NO_FORK=1 switches it off):
gpurun -- python tools/fork_ab.py [sizes...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from difffr_b200 import scenes  # noqa: E402
from difffr_b200.cabi import Context  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [101840, 238000, 1048576]
for n in sizes:
    sc = scenes.dam_break_scene(n, n_boxes=bench.N_BOXES)
    cfg = dict(bench.CFG, cfl_method=0, time_step_size=1.0e-3)
    steps = 200 if n < 500000 else 60
    res = {}
    for label, env in ((("fork", "0"), ("serial", "1"), ("fork again", "0")) if n < 500000 else (("fork", "0"), ("serial", "1"))):
        os.environ["DFR_NO_FORK"] = env
        ctx = scenes.build_context(lambda **k: Context(device=0, **k), sc, **cfg)
        ctx.step(20)
        ms0, l0 = ctx.device_time_ms()
        ctx.step(steps)
        ms1, l1 = ctx.device_time_ms()
        res[label] = (ms1 - ms0) / steps
        st = ctx.body_state(1)
        print(f"{n} {label}: {res[label]:.4f} ms/step, {(l1 - l0) / steps:.1f} launches/step, x = {st['x'].tolist()}", flush=True)
        del ctx
