"""The CUDA path (through the C ABI) against golden vectors recorded from the reference's own code."""
import pytest

import golden_util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_util.CASES)
def test_cuda_reproduces_reference_golden_vectors(gpu_factory, name):
    # north-star tolerances: state 1e-6, gradients 1e-4; the observed agreement is ~1e-13
    worst = golden_util.replay_and_compare(gpu_factory, name, state_tol=1e-6, grad_tol=1e-4)
    assert worst <= 1e-9
