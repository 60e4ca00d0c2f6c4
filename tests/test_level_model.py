"""The level-set propagation of the GPU (level.cu k_propagate: pushes from the front of the previous sweep, unsigned
atomicMin on the bit pattern or on an order-preserving key, first push claims, stop at the cutoff) as a pure-Python
model against the reference's Jacobi sweeps over all particles (simulation.rs:729-801) on random particle clouds:
same assigned set, bit-identical values, same sweep count; with the cutoff fewer sweeps and a bit-identical field after
the clamp of the smoothing pass (simulation.rs:833-836)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_front_pushes_equal_the_reference_sweeps():
    import level_model as m
    cut = 0
    for seed in range(8):  # even seeds: values <= 0 (EmptyAngle), odd seeds: either sign (CenterDiff)
        full, with_cutoff = m.run_case(seed, n=200 + 30 * (seed % 3))
        assert full >= 3
        cut += full - with_cutoff
    assert cut > 0  # the cutoff really ended some propagations early


def test_level_keys_preserve_the_order_of_floats_of_either_sign():
    import level_model as m
    assert m.key_order_ok(np.random.default_rng(1))
