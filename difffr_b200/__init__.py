"""difffr_b200 — B200-native differentiable DFSPH time step (hot path of zhehaoli1999/DiffFR).

Only what the path needs lives here: `csrc/` (CUDA kernels + the C ABI of include/dfr.h),
`cabi.py` (ctypes binding of that ABI), `scenes.py` (synthetic scene inputs), and
`pysplishsplash/` (host-side mirror of the reference's pybind11 surface for this path).
"""
from .cabi import Config, Context, DfrError, StepInfo, GRAD_NAMES  # noqa: F401
