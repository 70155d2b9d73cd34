"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): the radial-slab path with its ghost-ring exchange (peer-memory
stores from the transport kernel, and the ncclSend / ncclRecv path) and the dt all-reduce must reproduce the single-GPU result bit for bit (the reference's own np-independence claim,
constants.h:17), for an isothermal and an adiabatic planet-disk.  Launches tools/multi_gpu_check.py under torchrun."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("halo", ["peer", "nccl"])
@pytest.mark.parametrize("physics", ["isothermal_planet", "adiabatic_planet"])
def test_nranks_equals_one_rank(physics, halo):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 4 if n >= 4 else 2
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--physics", physics,
           "--nrad", "256", "--naz", "512", "--steps", "12"]
    env = dict(os.environ)
    if halo == "nccl":  # the ncclSend / ncclRecv exchange instead of the transport kernel's peer-memory stores
        env["FARGO_B200_HALO"] = "nccl"
    else:
        env.pop("FARGO_B200_HALO", None)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert lines, res.stdout[-2000:] + res.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["dt_bit_equal"], out
    if halo == "nccl":
        assert out["halo_mode"] == "nccl send/recv", out
    print(physics, halo, "->", out["halo_mode"])
    for name, st in out["fields"].items():
        assert st["n_diff"] == 0, (name, st)
    assert res.returncode == 0
