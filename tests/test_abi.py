"""The C-ABI library loads and exports every symbol include/dfr.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from difffr_b200 import cabi


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dfr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dfr_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_what_the_binding_lists():
    assert header_symbols() == sorted("dfr_" + s for s in cabi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = cabi.load_library()
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_oracle_exports_the_same_surface(oracle_lib):
    # no kernels to profile on the CPU side, and the CPU oracle is one domain (a slab run must equal it)
    skip = {"dfr_set_profiling", "dfr_get_kernel_profile", "dfr_slab_plan", "dfr_slab_unique_id", "dfr_slab_configure", "dfr_slab_info",
            "dfr_slab_local_ids", "dfr_load_fluid_state_local"}
    missing = [s for s in header_symbols() if s not in skip and not hasattr(oracle_lib, "orc_" + s[4:])]
    assert not missing, missing


def test_config_layout_and_defaults_agree(oracle_lib):
    """dfr_default_config / orc_default_config fill the ctypes mirror identically (catches layout drift)."""
    lib = cabi.load_library()
    a, b = cabi.Config(), cabi.Config()
    lib.dfr_default_config.argtypes = [ctypes.POINTER(cabi.Config)]
    lib.dfr_default_config.restype = None
    oracle_lib.orc_default_config.argtypes = [ctypes.POINTER(cabi.Config)]
    oracle_lib.orc_default_config.restype = None
    lib.dfr_default_config(ctypes.byref(a))
    oracle_lib.orc_default_config(ctypes.byref(b))
    assert bytes(a) == bytes(b)
    assert a.particle_radius == 0.025 and a.min_iterations == 2 and a.max_iterations == 100
    assert a.gradient_mode == 1 and a.cfl_method == 1 and abs(a.gravitation[1] + 9.81) < 1e-15


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(cabi.DfrError) as ei:
        cabi.Context(device=0)
    assert ei.value.code == -2  # DFR_ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    """Nothing under difffr_b200/ or include/ may import, link or call oracle/."""
    bad = []
    for base in ("difffr_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if f.endswith(".py"):
                        hit = re.search(r"liboracle|import\s+oracle|from\s+oracle|oracle[/\\]", txt)
                    else:  # comments may cite the oracle; code may not
                        code = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
                        code = re.sub(r"//[^\n]*", "", code)
                        hit = re.search(r"liboracle|oracle/|orc_[a-z]", code)
                    if hit:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
