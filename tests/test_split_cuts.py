"""SplitDomain (split.cpp:38-55) as this library cuts the grid: contiguous slabs whose COST is balanced (rings inside a damping
zone are also read and written by the damping kernel), never thinner than 2 x CPUOVERLAP rings, and the reference's equal
cuts when there is nothing to balance or FARGO_B200_SPLIT=equal asks for them.  Pure host arithmetic of the CUDA library:
runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from fargocpt_b200 import abi, synthetic


def _cuts(params, radii, np_ranks):
    lib = C.CDLL(abi.LIB_PATH)
    lib.fargo_split_cuts.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
    lib.fargo_split_cuts.restype = C.c_int
    radii = np.ascontiguousarray(radii, dtype=np.float64)
    out = (C.c_int * (np_ranks + 1))()
    rc = lib.fargo_split_cuts(C.byref(params), radii.ctypes.data_as(C.POINTER(C.c_double)), np_ranks, out)
    return rc, list(out)


def _equal(nrad, n):
    low, rem = divmod(nrad, n)
    return [(low + 1) * r if r < rem else (low + 1) * rem + (r - rem) * low for r in range(n + 1)]


@pytest.mark.parametrize("n", [2, 3, 4, 8])
def test_cost_balanced_cuts(n, monkeypatch):
    monkeypatch.delenv("FARGO_B200_SPLIT", raising=False)
    cfg = synthetic.make_config("adiabatic_planet", 8192, 16384)
    params, radii = synthetic.params_from_config(cfg), synthetic.radii_from_config(cfg)
    assert params.damping
    rc, cut = _cuts(params, radii, n)
    assert rc == 0 and cut[0] == 0 and cut[-1] == 8192
    sizes = np.diff(cut)
    assert (sizes >= 2 * abi.CPUOVERLAP).all()
    # the damping zones sit at the two ends: the edge ranks get fewer rings than the middle ones
    if n >= 3:
        assert sizes[0] < sizes[1] and sizes[-1] < sizes[-2]
    # cost per rank (1 per ring + the damping weight the library documents: 1.85 % per damped field) is level to within 2 rings
    rmid = 0.5 * (radii[:-1] + radii[1:])
    damped = (rmid < params.rmin * params.damping_inner_limit) | (rmid > params.rmax * params.damping_outer_limit)
    cost = 1.0 + 0.0185 * 4 * damped
    per_rank = [cost[cut[r]:cut[r + 1]].sum() for r in range(n)]
    assert max(per_rank) - min(per_rank) < 2.5, per_rank
    # FARGO_B200_SPLIT=equal: the reference's cut points
    monkeypatch.setenv("FARGO_B200_SPLIT", "equal")
    rc, cut = _cuts(params, radii, n)
    assert rc == 0 and cut == _equal(8192, n)


def test_equal_cuts_without_damping_and_refusal_of_thin_slabs(monkeypatch):
    monkeypatch.delenv("FARGO_B200_SPLIT", raising=False)
    cfg = synthetic.make_config("adiabatic_planet", 515, 64)
    params, radii = synthetic.params_from_config(cfg), synthetic.radii_from_config(cfg)
    params.damping = 0
    rc, cut = _cuts(params, radii, 4)
    assert rc == 0 and cut == _equal(515, 4)
    rc, _ = _cuts(params, radii, 64)  # 515 / 64 = 8 rings < 2 x CPUOVERLAP: the reference dies here (split.cpp:30-36)
    assert rc != 0
