"""Population evaluation with several contexts side by side on one GPU (difffr_b200/rollouts.py).

The reference evaluates a CMA-ES population one candidate after the other (opt-ng.py:174-211).  Here contexts are
independent, so a rank may drive several at once from host threads; the result of a candidate must not depend on that."""
import numpy as np
import pytest

from difffr_b200 import rollouts, scenes

pytestmark = pytest.mark.gpu

CFG = dict(surface_tension_method=2, surface_tension=0.2, max_error=0.05, target_time=0.012, uniform_acc_rb_time=0.004)


def test_concurrent_contexts_give_the_serial_results(gpu_factory):
    sc = scenes.dam_break_scene(3000, n_boxes=1)
    rng = np.random.default_rng(12)  # opt-ng.py: seed 12
    cands = [(rng.normal(size=3) * 0.5, rng.normal(size=3)) for _ in range(7)]

    def make():
        return scenes.build_context(gpu_factory, sc, **CFG)

    serial, steps1 = rollouts.run_population(make, cands, body=1)
    conc, steps3 = rollouts.run_population(make, cands, body=1, concurrency=3)
    assert np.array_equal(steps1, steps3) and steps1.min() > 3
    assert np.array_equal(serial, conc)  # bit for bit: deterministic kernels, no shared state between contexts
    assert np.abs(serial[0] - serial[1]).max() > 0  # candidates really differ
    # contexts handed in by the caller are reused (no re-finalize per population)
    ctxs = [make(), make()]
    again, _ = rollouts.run_population(None, cands, body=1, contexts=ctxs)
    assert np.array_equal(serial, again)
