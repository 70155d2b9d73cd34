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


def _adversarial_rings(ns, rng):
    """Rows whose sequential sums exercise every branch of the scan ring sum (csrc/kernels_ringsum.cuh)."""
    import numpy as np
    j = np.arange(ns)
    rows = []
    # disk-like rings: one sign, smooth + noise (few events per binade)
    for v0 in (0.7, -0.7, 1.3e-3, 4.1e5):
        rows.append(v0 * (1.0 + 1e-3 * np.sin(2 * np.pi * 3 * j / ns) + 1e-6 * rng.standard_normal(ns)))
    # corotation-like rings: partial sums hover around zero
    rows.append(1e-3 * np.sin(2 * np.pi * 2 * j / ns) + 1e-9 * rng.standard_normal(ns))
    rows.append(rng.standard_normal(ns))
    rows.append(rng.standard_normal(ns) * np.exp(rng.uniform(-30, 30, ns)))  # huge dynamic range
    # exact ties and powers of two: small integers times a power of two, all partial sums on grid points
    rows.append(rng.integers(1, 4, ns).astype(float) * 2.0 ** -3)
    rows.append(np.full(ns, 1.0))
    rows.append(np.full(ns, 1.0 + 2.0 ** -52))
    rows.append(np.where(j % 2 == 0, 1.0, 2.0 ** -53))   # every second add is a tie against a power of two or odd mantissa
    rows.append(np.where(j % 3 == 0, 1.0, 2.0 ** -54 * 3))
    rows.append(np.where(j == 0, 2.0 ** 60, rng.integers(1, 1 << 9, ns).astype(float)))  # increments near ulp / 2 of the sum
    rows.append(np.where(j == 0, 1.0, -2.0 ** -54))        # creeping down onto a power of two from above
    rows.append(np.where(j == 0, 1.0, -(2.0 ** -53) * (1 + (j % 2))))
    # zeros, signed zeros, denormals, cancellation to exactly zero, overflow, inf / nan
    rows.append(np.zeros(ns))
    rows.append(np.full(ns, -0.0))
    rows.append(np.full(ns, 5e-324) * rng.integers(0, 3, ns))
    rows.append(np.where(j % 2 == 0, 0.3, -0.3))
    rows.append(np.full(ns, 1.7e308))
    r = rng.standard_normal(ns)
    r[ns // 2] = np.inf
    rows.append(r)
    r = rng.standard_normal(ns)
    r[ns // 3] = np.nan
    rows.append(r)
    # random rows with random scales and offsets
    for _ in range(40):
        rows.append(rng.standard_normal(ns) * 10.0 ** rng.uniform(-8, 8) + rng.choice([0.0, 1.0, -1.0]) * 10.0 ** rng.uniform(-8, 8))
    return np.array(rows)


@pytest.mark.parametrize("ns", [16384, 4096, 1000, 377, 255, 7, 1])
def test_scan_ring_sums_are_the_sequential_sums(ns):
    """cfl.cpp:199-204 / TransportEuler.cpp:215-219 sum a ring strictly sequentially; dt and Nshift inherit every rounding of
    it.  The scan kernel (a warp per ring) and the chain kernel (a thread per ring) must both return exactly that sum."""
    import numpy as np
    from fargocpt_b200 import HydroContext
    meta, z = reftools.load_golden("iso_star")
    ctx = HydroContext(reftools.make_params(meta["params"]), z["radii"])
    x = _adversarial_rings(ns, np.random.default_rng(ns))
    with np.errstate(all="ignore"):
        # sum = 0.0; for j: sum += v[j]  (ufunc.accumulate adds strictly left to right; the leading 0.0 matters for -0.0)
        ref = np.cumsum(np.concatenate([np.zeros((x.shape[0], 1)), x], axis=1), axis=1)[:, -1]
    scan, chain = ctx.selftest_ringsum(x)

    def same(a, b):
        return (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
    assert same(chain, ref).all(), np.nonzero(~same(chain, ref))[0]
    bad = np.nonzero(~same(scan, ref))[0]
    assert bad.size == 0, (bad, scan[bad[:5]], ref[bad[:5]])
