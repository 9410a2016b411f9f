"""Test-side helper (it runs the CPU ORACLE, which only tests/ may do): the bench scene family advanced a few steps,
dumped as positions + geometry for tools/gather_model.py (an offline model of the neighbour gathers' access pattern).
  python tests/dump_bench_state.py [n_particles] [steps] [out.npz]      (aspect 4:0.5:0.5 gives the 1M scene's ~90-cell rows)"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from difffr_b200 import scenes  # noqa: E402
from difffr_b200.cabi import Context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = sys.argv[3] if len(sys.argv) > 3 else "/tmp/bench_state.npz"
lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle_fast.so"))
sc = scenes.dam_break_scene(n, n_boxes=bench.N_BOXES, tank_aspect=(4.0, 0.5, 0.5))
ctx = scenes.build_context(lambda **k: Context(lib=lib, prefix="orc_", **k), sc, **bench.CFG)
ctx.step(steps)
np.savez_compressed(out, x=ctx.fluid("position"), radius=sc["radius"], steps=steps)
print("wrote", out, ctx.num_fluid, "particles after", steps, "steps")
