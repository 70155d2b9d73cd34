#!/usr/bin/env python
"""Records monitor/Quantities.dat of the UNMODIFIED reference (oracle/_ref/fargocpt_exe_ieee) for golden fixtures that
already exist, into tests/golden/quantities.json.

The fixture's own .yml is re-run with WriteDiskQuantities switched on (output only: the fields, dt sequence and body states
of the run are the recorded ones) and OMP_NUM_THREADS=1, so that the reference's OpenMP sum reductions
(quantities.cpp:51-480) run in index order and the recorded numbers are reproducible bit for bit.  Every hydro step of
these configs is one monitor step (Nmonitor: 1), so row k of Quantities.dat belongs to snapshot k.

usage: python tests/golden/make_quantities.py   (needs /root/reference built by oracle/Makefile.ref)
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "fargocpt_exe_ieee")
CASES = {"adia_planet_100": (50, 100), "iso_planet_100": (50, 100), "rey_star": (1, 2), "adia_star": (3, 6)}
COLUMNS = {"mass": 3, "angular_momentum": 5, "internal_energy": 7, "kinetic_energy": 8, "radial_kinetic_energy": 10,
           "azimuthal_kinetic_energy": 11, "viscous_dissipation": 14, "luminosity": 15,
           # the mass-weighted columns (fargo_monitor_disk)
           "radius": 4, "eccentricity": 12, "periastron": 13, "aspect_ratio": 26,
           "advection_torque": 32, "viscous_torque": 33,
           # ... and the ones that read the POTENTIAL grid of the last step's start
           "total_energy": 6, "potential_energy": 9, "gravitational_torque": 34,
           # MassDelta's boundary flows since the previous row (TransportEuler.cpp:578-608)
           "inner_boundary_inflow": 17, "inner_boundary_outflow": 18, "outer_boundary_inflow": 19, "outer_boundary_outflow": 20}


def main():
    if not os.path.exists(EXE):
        raise SystemExit("build the reference first: make -f oracle/Makefile.ref -j8")
    out = {}
    for name, snaps in CASES.items():
        cfg = yaml.safe_load(open(os.path.join(HERE, name + ".yml")))
        for k in [k for k in cfg if k.startswith("_")]:
            cfg.pop(k)
        cfg["WriteDiskQuantities"] = "Yes"
        tmp = tempfile.mkdtemp(prefix="quant_" + name + "_")
        cfg["OutputDir"] = os.path.join(tmp, "out")
        ypath = os.path.join(tmp, "cfg.yml")
        yaml.safe_dump(cfg, open(ypath, "w"), sort_keys=False)
        res = subprocess.run([EXE, "start", ypath], cwd=tmp, env=dict(os.environ, OMP_NUM_THREADS="1"), capture_output=True, text=True)
        if res.returncode != 0:
            print(res.stdout[-2000:], res.stderr[-2000:])
            raise SystemExit(f"reference run failed for {name}")
        rows = {}
        for line in open(os.path.join(cfg["OutputDir"], "monitor", "Quantities.dat")):
            if line.startswith("#"):
                continue
            f = line.split()
            if int(f[0]) in snaps and int(f[0]) == int(f[1]):
                rows[int(f[0])] = {q: float.fromhex(float(f[c]).hex()) for q, c in COLUMNS.items()}  # %.16e round-trips a double
        # column 9 of monitor/nbody1.dat: the circumplanetary mass (ComputeCircumPlanetaryMasses, circumplanetary_mass.cpp:11-51)
        pfile = os.path.join(cfg["OutputDir"], "monitor", "nbody1.dat")
        if os.path.exists(pfile):
            for line in open(pfile):
                if line.startswith("#"):
                    continue
                f = line.split()
                if int(f[0]) in snaps and int(f[0]) == int(f[1]):
                    rows[int(f[0])]["mdcp"] = float(f[9])  # %.18g round-trips a double
        out[name] = {str(k): rows[k] for k in snaps}
        shutil.rmtree(tmp)
        print(name, {k: rows[k]["mass"] for k in snaps})
    json.dump({"source": "oracle/_ref/fargocpt_exe_ieee, OMP_NUM_THREADS=1, monitor/Quantities.dat (code units)", "quantities": out},
              open(os.path.join(HERE, "quantities.json"), "w"), indent=1)


if __name__ == "__main__":
    sys.exit(main())
