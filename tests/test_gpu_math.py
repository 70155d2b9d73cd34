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
