"""fargocpt_b200 — B200-native (sm_100a) implementation of FargoCPT's per-timestep hydro step.

The product is the CUDA library fargocpt_b200/csrc/libfargo_b200.so behind the C ABI declared in
include/fargo_b200.h; this package only holds the ctypes plumbing used by tests and bench.py.
"""
from .abi import FargoParams, FargoBodies, HydroContext, load_library  # noqa: F401
