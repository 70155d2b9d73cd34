"""Synthetic protoplanetary disks for bench.py and the large-size property tests.

Host-side helper (numpy): builds the FargoParams, the interface radii and an equilibrium power-law disk
(Sigma ~ r^-p, locally isothermal temperature profile h = h0 r^f, pressure-supported Keplerian rotation)
with a small deterministic perturbation so limiter / upwind branches are exercised.  The formulas follow the
reference's profile initialisation (init.cpp:1257-1300 energy, Theo.cpp initial_locally_isothermal_smoothed_v_az)
but are NOT claimed bit-identical to it — parity tests use fixtures recorded from the reference instead.
"""
import math

import numpy as np

from . import abi

# code-unit constants of the reference for l0 = 30 au, m0 = 1 solMass (constants.yml written by the reference)
CONSTS_30AU = {"G": 0.99999999999999978, "R": 1.0, "sigma": 2565.02264205042, "c": 55129.195248966156,
               "temperature_unit_K": 3556.6728100870596}

# the BASELINE.json workloads.  Keys are FargoCPT YAML keys.
PHYSICS = {
    # C2: test/cold_disk_planet physics (ideal EOS, TW art-visc + dissipation, no physical viscosity, q = 2e-5)
    "cold_disk_planet": dict(EquationOfState="Ideal", AdiabaticIndex=1.4, ViscousAlpha=0.0, HeatingViscous="no",
                             CoolingBetaLocal="no", ArtificialViscosity="TW", ArtificialViscosityFactor=3.0,
                             planet_mass=2e-5),
    # C3/C5: adiabatic + alpha viscosity + viscous heating + beta cooling (+ Jupiter-mass planet for C5)
    "adiabatic_planet": dict(EquationOfState="Ideal", AdiabaticIndex=1.4, ViscousAlpha=1e-3, HeatingViscous="yes",
                             CoolingBetaLocal="yes", CoolingBeta=10, CoolingBetaReference="reference",
                             ArtificialViscosity="TW", ArtificialViscosityFactor=3.0, planet_mass=1e-3),
    # C4: examples/config.yml physics (isothermal, alpha 1e-3, TW, Jupiter, OmegaFrame 1)
    "isothermal_planet": dict(EquationOfState="Isothermal", ViscousAlpha=1e-3, ArtificialViscosity="TW",
                              ArtificialViscosityFactor=1.41, OmegaFrame=1.0, planet_mass=1e-3, Rmax=2.5,
                              DampingInnerLimit=1.1, DampingOuterLimit=0.9, DampingTimeFactor=0.1, FlaringIndex=0.0),
}

BASE = {
    "Sigma0": 0.005743125733951172, "SigmaSlope": 1.0, "SigmaFloor": 1e-7, "AspectRatio": 0.05,
    "FlaringIndex": 0.2857142857142857, "ArtificialViscosityDissipation": "Yes", "MinimumTemperature": "3 K",
    "MaximumTemperature": "1e100 K", "CFL": 0.5, "CFLmaxVar": 1.1, "HeatingCoolingCFLlimit": 1.0, "mu": 2.35,
    "ThicknessSmoothing": 0.6, "Transport": "FARGO", "Integrator": "Euler", "InnerBoundary": "Reflecting",
    "OuterBoundary": "Reflecting", "Damping": "Yes", "DampingInnerLimit": 1.311, "DampingOuterLimit": 0.763,
    "DampingTimeFactor": 0.05, "OmegaFrame": 0.0, "Rmin": 0.4, "Rmax": 2.0, "RadialSpacing": "Logarithmic",
    "BodyForceFromPotential": "Yes", "FirstDT": 0.1,
}
for _q in ("Energy", "VRadial", "VAzimuthal", "SurfaceDensity"):
    for _s in ("Inner", "Outer"):
        BASE[f"Damping{_q}{_s}"] = "Initial"


def make_config(physics, nrad, naz, **overrides):
    cfg = dict(BASE)
    cfg.update(PHYSICS[physics])
    cfg.update(overrides)
    cfg["Nrad"], cfg["Naz"] = int(nrad), int(naz)
    return cfg


def radii_from_config(cfg):
    """init.cpp:93-110: interface radii, ring 1 starts at Rmin (one ghost ring on each side)."""
    n, rmin, rmax = cfg["Nrad"], float(cfg["Rmin"]), float(cfg["Rmax"])
    k = np.arange(n + 1, dtype=np.float64)
    if str(cfg["RadialSpacing"]).lower().startswith("l"):
        f = math.pow(rmax / rmin, 1.0 / (n - 2.0))
        return np.array([rmin * math.pow(f, i - 1.0) for i in range(n + 1)])
    return rmin + (rmax - rmin) / (n - 2.0) * (k - 1.0)


def params_from_config(cfg, consts=CONSTS_30AU):
    from .config import params_from_config as _pfc
    d = _pfc(cfg, consts, cfg["Nrad"], cfg["Naz"], consts["temperature_unit_K"])
    return abi.FargoParams.from_dict(d)


def disk_fields(cfg, radii, consts=CONSTS_30AU, perturb=1e-3):
    """Returns dict Sigma, vrad, vazi, energy (global arrays)."""
    nrad, naz = cfg["Nrad"], cfg["Naz"]
    G, M = consts["G"], 1.0
    rs, ri = radii[1:], radii[:-1]
    rmed = 2.0 / 3.0 * (rs ** 3 - ri ** 3) / (rs ** 2 - ri ** 2)
    h0, fl, p = float(cfg["AspectRatio"]), float(cfg["FlaringIndex"]), float(cfg["SigmaSlope"])
    gamma = float(cfg.get("AdiabaticIndex", 1.4))
    omega_f = float(cfg.get("OmegaFrame", 0.0))
    phi = 2.0 * np.pi / naz * np.arange(naz)
    sigma_r = float(cfg["Sigma0"]) * rmed ** (-p)
    pert = 1.0 + perturb * np.outer(np.cos(11.0 * np.log(rmed)), np.sin(3.0 * phi)) \
        + 0.5 * perturb * np.outer(np.sin(29.0 * np.log(rmed)), np.cos(17.0 * phi))
    sigma = sigma_r[:, None] * pert
    h = h0 * rmed ** fl
    cs2 = h * h * G * M / rmed
    energy = sigma * cs2[:, None] / (gamma - 1.0)
    # pressure-supported rotation: v^2 = v_K^2 [1 - h^2 (1 + p - 2 f)]
    vk = np.sqrt(G * M / rmed)
    vphi = vk * np.sqrt(np.maximum(1.0 - h * h * (1.0 + p - 2.0 * fl), 0.0)) - omega_f * rmed
    vazi = np.repeat(vphi[:, None], naz, axis=1) * (1.0 + 0.1 * perturb * np.outer(np.ones(nrad), np.cos(5.0 * phi)))
    vrad = np.zeros((nrad + 1, naz))
    return {"Sigma": np.ascontiguousarray(sigma), "vrad": vrad, "vazi": np.ascontiguousarray(vazi),
            "energy": np.ascontiguousarray(energy)}


class PlanetOrbit:
    """Host-side two-body coupling for the synthetic workloads: star at the origin (HydroFrameCenter: primary),
    one planet on a circular orbit; indirect term = minus the star's acceleration by the planet
    (frame_of_reference.cpp:107-128), expressed in the frame rotating with OmegaFrame."""

    def __init__(self, cfg, consts=CONSTS_30AU):
        self.G, self.mp = consts["G"], float(cfg.get("planet_mass", 0.0))
        self.a, self.omega_frame = 1.0, float(cfg.get("OmegaFrame", 0.0))
        self.omega = math.sqrt(self.G * (1.0 + self.mp) / self.a ** 3)

    def bodies(self, t, dt=0.0):
        if self.mp <= 0.0:
            return abi.FargoBodies.make([0.0], [0.0], [1.0], omega_frame=self.omega_frame)
        ang = (self.omega - self.omega_frame) * t
        x, y = self.a * math.cos(ang), self.a * math.sin(ang)
        acc = self.G * self.mp / self.a ** 2  # star is pulled towards the planet
        return abi.FargoBodies.make([0.0, x], [0.0, y], [1.0, self.mp], indirect=(-acc * x / self.a, -acc * y / self.a),
                                    omega_frame=self.omega_frame)
