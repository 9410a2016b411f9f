"""bench.py's roofline arithmetic (no GPU): the algorithmic-bytes model of SURVEY.md 8d / DESIGN.md 4."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_step_bytes_match_the_survey_formula():
    # B(n, D, P) = 1212 + 48 n + D (176 + 8 n) + P (184 + 8 n); the survey's nominal points
    assert bench.algorithmic_bytes_per_particle_step(35, 1, 2) == 4276
    assert bench.algorithmic_bytes_per_particle_step(35, 1, 11) == 8452


def test_kernel_bytes_follow_the_profile_names():
    n = 30.0
    plain = bench.kernel_bytes_per_particle("(k_rho<PRESSURE, RHO_ITER>)", n)
    assert plain == bench.KERNEL_BYTES["k_rho"](n, 0.0)
    # the fused passes stream the extra arrays they read / write
    assert bench.kernel_bytes_per_particle("(k_rho<false, RHO_WARM, RHO_X_DENSITY>)", n) == plain + 80
    assert bench.kernel_bytes_per_particle("(k_rho<false, RHO_PLAIN, RHO_X_NORMALS>)", n) == plain + 32
    assert bench.kernel_bytes_per_particle("(k_rho<false, RHO_ITER, RHO_X_NONPRESSURE>)", n) == plain + 64
    assert bench.kernel_bytes_per_particle("k_nbr_build", n) == 40 + 4 * n
    assert bench.kernel_bytes_per_particle("k_some_new_kernel", n) is None
    assert bench.kernel_class("(k_push<PRESSURE, true>) (idle)".replace(" (idle)", "")) == "k_push"


def test_workload_config_names_the_baseline_config():
    one = bench.workload_config(1)
    assert "configs[4]" in one["workload"] and one["particles_per_gpu"] == 1 << 20
    slab = bench.workload_config(8, "slab")
    assert "8 x 1048576" in slab["workload"] and "slab" in slab["parallelism"]
    assert bench.workload_config(8, "rollouts")["rollouts"] == 8


def test_clock_sampler_degrades_without_a_gpu():
    """bench.py samples clocks with NVML (nvidia-smi as fallback) during the timed region; without either it must still
    return the `clocks` object instead of failing the run."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present: the sampler's live path is exercised by bench.py itself")
    s = bench.ClockSampler(0)
    s.start()
    s.mark_begin()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
