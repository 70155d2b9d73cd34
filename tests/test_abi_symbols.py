"""CPU-side checks of the drop-in boundary: libfargo_b200.so loads without a GPU, exports every function
include/fargo_b200.h declares, its POD structs have the layout the ctypes mirror assumes, and creating a context
without a CUDA device fails loudly (there is no CPU fallback).  The oracle library exports the same surface under
the fargo_oracle_ prefix (it is the checker, never the product)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fargo_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(fargo_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def _lib():
    from fargocpt_b200 import abi
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return C.CDLL(abi.LIB_PATH)


def test_every_declared_symbol_is_exported():
    lib = _lib()
    names = _declared_functions()
    assert len(names) >= 30, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layout_matches_header():
    """sizeof(fargo_params) / sizeof(fargo_bodies) as the C compiler sees them == the ctypes mirrors."""
    import subprocess
    import tempfile
    from fargocpt_b200 import abi
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "fargo_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n",'
                             'sizeof(fargo_params),sizeof(fargo_bodies),offsetof(fargo_params,damp_energy),'
                             'offsetof(fargo_bodies,omega_frame));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
        sp, sb, o1, o2 = map(int, subprocess.check_output([exe]).split())
    assert C.sizeof(abi.FargoParams) == sp
    assert C.sizeof(abi.FargoBodies) == sb
    assert abi.FargoParams.damp_energy.offset == o1
    assert abi.FargoBodies.omega_frame.offset == o2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is what this test is about")
    import reftools
    from fargocpt_b200 import HydroContext
    meta, z = reftools.load_golden("iso_star")
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA|CUDA"):
        HydroContext(reftools.make_params(meta["params"]), z["radii"])


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fargocpt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                # (abi.py's docstring mentions that the tests reuse its Handle class for the oracle; what must not exist
                # is code that loads or calls it)
                assert "reftools" not in txt and not re.search(r"CDLL\([^)]*oracle", txt) and "fargo_oracle_create" not in txt, \
                    os.path.join(dirpath, f)
