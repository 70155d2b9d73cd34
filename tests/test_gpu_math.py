"""The branch-free IEEE arithmetic of the marching kernels (csrc/fargo_math.h) must be bit-identical to the plain
`/`, sqrt() and exp() wherever its validity flag says so — that is what makes the fused kernels reproduce the
reference CPU build's IEEE results.  Runs the device self-test on ~2.7e8 random / adversarial operand pairs."""
import pytest

import reftools

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wide", [False, True], ids=["physical-range", "full-exponent-range"])
def test_fast_paths_match_operators(wide):
    from fargocpt_b200 import HydroContext
    meta, z = reftools.load_golden("iso_star")
    ctx = HydroContext(reftools.make_params(meta["params"]), z["radii"])
    tot = {"div_mismatch": 0, "sqrt_mismatch": 0, "exp_mismatch": 0, "div_fast": 0, "pairs": 0}
    for seed in (1, 2):
        r = ctx.selftest_math(seed=seed, blocks=2048, per_thread=256, wide=wide)
        for k in tot:
            tot[k] += r[k]
    assert tot["div_mismatch"] == 0 and tot["sqrt_mismatch"] == 0 and tot["exp_mismatch"] == 0, tot
    # the fast path must actually be the common case for physical magnitudes
    if not wide:
        assert tot["div_fast"] > 0.8 * tot["pairs"], tot


def test_device_exp_is_the_hosts_libm_exp():
    """compression_heating's exp: the device evaluates glibc's algorithm operation by operation, so its results must be
    bit-identical with the libm exp of the host the oracle / reference run on (Python's math.exp calls libm)."""
    import math
    import numpy as np
    from fargocpt_b200 import HydroContext
    meta, z = reftools.load_golden("iso_star")
    ctx = HydroContext(reftools.make_params(meta["params"]), z["radii"])
    rng = np.random.default_rng(7)
    parts = [rng.uniform(-1e-2, 1e-2, 200000), rng.uniform(-1.0, 1.0, 100000), rng.uniform(-500.0, 500.0, 100000),
             rng.uniform(-1e-9, 1e-9, 50000), np.ldexp(rng.uniform(1.0, 2.0, 50000), rng.integers(-1074, 9, 50000)) * rng.choice([-1.0, 1.0], 50000),
             np.array([0.0, -0.0, 5e-324, -5e-324, 2.0 ** -54, -(2.0 ** -54), 511.9999999, -511.9999999])]
    x = np.concatenate(parts)
    x = x[np.abs(x) < 512.0]
    y = ctx.selftest_exp(x)
    ref = np.array([math.exp(v) for v in x])
    bad = np.nonzero(y.view(np.uint64) != ref.view(np.uint64))[0]
    assert bad.size == 0, (bad.size, x[bad[:5]], y[bad[:5]], ref[bad[:5]])
