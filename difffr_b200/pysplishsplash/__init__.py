"""`pysplishsplash` surface of the DiffDFSPH path, backed by the B200 CUDA library.

Drop-in for the reference's pybind11 module of the same name (pySPlisHSPlasH/main.cpp:47-76): the
optimisation scripts do `import pysplishsplash as sph` and use `sph.Exec.SimulatorBase`,
`sph.Simulation.getCurrent()`, `TimeStepDiffDFSPH`, `BoundaryModelAkinci2012`,
`sph.Exec.RigidBodyGradientManager`, `sph.TimeManager`, `sph.GUI.Simulator_GUI_imgui`.  Put
`difffr_b200/` on `sys.path` (or `from difffr_b200 import pysplishsplash as sph`).

The compiled part is `_core` (bindings.cpp -> host/simulator_host.hpp -> include/dfr.h -> libdfr.so).
There is no CPU fallback: without the built extension this import fails; without a CUDA device
`SimulatorBase.initSimulation()` raises `DfrError`.
"""
import ctypes as _ctypes
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_lib = _os.path.join(_here, "..", "csrc", "libdfr.so")
if not _os.path.exists(_lib):
    raise ImportError(f"{_lib} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first (no CPU fallback)")
_ctypes.CDLL(_lib, mode=_ctypes.RTLD_GLOBAL)  # _core links against it by soname

from ._core import *  # noqa: E402,F401,F403
from ._core import Exec, GUI, Utilities, DfrError, _load_scene_summary, _load_scene_full, _read_bgeo, _write_bgeo, _bgeo_of_state_file  # noqa: E402,F401
