"""FargoCPT YAML config -> fargo_params (host-side mirror of the subset of parameters.cpp / Interpret.cpp /
boundary_conditions/config.cpp / damping.cpp the hot path reads).  Key names, defaults and quirks follow the
reference; dimensional values are converted to code units with the unit factors the caller supplies."""
from . import abi

# config -> params (mirror of the subset of parameters.cpp / Interpret.cpp / boundary_conditions/config.cpp
# the hot path reads).  `consts` = code-unit constants as the reference printed them (constants.yml).

def _flag(v, default=False):
    if v is None:
        return default
    if isinstance(v, bool):
        return v
    return str(v).strip().lower()[0] in ("y", "t", "1")


def _num(v, unit_cgs=None):
    """'3 K' with unit_cgs=<code temperature unit in K> -> 3/unit; plain numbers are code units."""
    if isinstance(v, (int, float)):
        return float(v)
    parts = str(v).split()
    x = float(parts[0])
    if len(parts) > 1 and unit_cgs is not None:
        return x / unit_cgs
    return x


def params_from_config(cfg, consts, nrad, naz, temp_unit_K=1.0, units=None):
    """units: code -> cgs factors of units.yml ({"density": ..., "opacity": ...}); only the radiative terms need them."""
    g = {k.lower(): v for k, v in cfg.items()}

    def get(key, default=None):
        return g.get(key.lower(), default)

    d = {}
    d["nrad"], d["naz"] = int(nrad), int(naz)
    spacing = {"l": "logarithmic", "a": "arithmetic", "e": "exponential"}.get(
        str(get("RadialSpacing", "Arithmetic")).lower()[:1], "custom")
    d["radial_spacing"] = abi.SPACING[spacing]
    d["rmin"], d["rmax"] = float(get("Rmin")), float(get("Rmax"))
    eos = str(get("EquationOfState", "Isothermal")).lower()
    d["adiabatic"] = 1 if eos in ("ideal", "adiabatic", "perfect", "pvte", "pvtelaw") else 0
    d["pvte"] = 1 if eos in ("pvte", "pvtelaw") else 0  # Interpret.cpp:453-491
    d["gamma"] = float(get("AdiabaticIndex", 1.4))
    d["mu"] = float(get("mu", 1.0))
    d["aspectratio_ref"] = float(get("AspectRatio", 0.05))
    d["flaring_index"] = float(get("FlaringIndex", 0.0))
    d["sigma0"] = _num(get("Sigma0", 173.0))
    d["sigma_floor"] = float(get("SigmaFloor", 1e-9))
    d["sigma_slope"] = float(get("SigmaSlope", 0.0))
    d["minimum_temperature"] = _num(get("MinimumTemperature", "3 K"), temp_unit_K)
    d["maximum_temperature"] = _num(get("MaximumTemperature", "1e100 K"), temp_unit_K)
    d["G"], d["Rgas"], d["sigma_sb"], d["c_light"] = consts["G"], consts["R"], consts["sigma"], consts["c"]
    d["hydro_center_mass"] = consts.get("hydro_center_mass", 1.0)
    d["cfl"] = float(get("CFL", 0.5))
    d["cfl_max_var"] = float(get("CFLmaxVar", 1.1))
    d["heating_cooling_cfl_limit"] = float(get("HeatingCoolingCFLlimit", 10.0))  # parameters.cpp:797
    integ = str(get("Integrator", "Euler")).lower()
    d["leapfrog"] = 0 if integ.startswith("e") else 1
    d["fast_transport"] = 1 if str(get("Transport", "FARGO")).lower().startswith("f") else 0
    # Interpret.cpp:640-664 compares case-sensitively: only the exact strings "mc" / "m" select MC
    d["flux_limiter"] = 1 if str(get("FluxLimiter", "VanLeer")) in ("mc", "m") else 0
    d["artificial_viscosity"] = abi.ARTVISC[str(get("ArtificialViscosity", "SN")).lower()]
    d["artificial_viscosity_factor"] = float(get("ArtificialViscosityFactor", 1.41))
    d["artificial_viscosity_dissipation"] = int(_flag(get("ArtificialViscosityDissipation"), True))
    d["viscous_alpha"] = float(get("ViscousAlpha", 0.0))
    d["constant_viscosity"] = _num(get("ConstantViscosity", 0.0))
    d["alpha_mode"] = int(get("AlphaMode", 0))  # parameters.cpp:704-706
    d["alpha_cold"] = float(get("AlphaCold", 0.01))
    d["alpha_hot"] = float(get("AlphaHot", 0.1))
    d["stabilize_viscosity"] = int(get("StabilizeViscosity", 0))
    d["radial_viscosity_factor"] = float(get("RadialViscosityFactor", 1.0))
    d["heating_viscous"] = int(_flag(get("HeatingViscous"), True))  # parameters.cpp:561
    d["heating_viscous_factor"] = float(get("HeatingViscousFactor", 1.0))
    d["cooling_beta"] = int(_flag(get("CoolingBetaLocal"), False))
    d["cooling_beta_value"] = float(get("CoolingBeta", 1.0))
    d["cooling_beta_ramp_up"] = _num(get("CoolingBetaRampUp", 0.0))
    d["cooling_beta_reference"] = abi.BETA_REF[str(get("CoolingBetaReference", "zero")).lower()]
    # radiative surface cooling / stellar irradiation (parameters.cpp:389-435, 628-632; planetary_system.cpp:137-146)
    sc = str(get("SurfaceCooling", "No")).lower()
    if sc not in ("no", "off", "false", "thermal", "scurve"):
        raise ValueError("SurfaceCooling: %s is outside this path" % sc)
    d["cooling_surface"] = int(sc == "thermal")
    # parameters.cpp:374-403: ScurveType Kimura (default) | Ichikawa
    d["cooling_scurve"] = 0 if sc != "scurve" else {"ichikawa": 1, "kimura": 2}[str(get("ScurveType", "Kimura")).lower()]
    d["surface_cooling_factor"] = float(get("CoolingRadiativeFactor", 1.0))
    d["heating_star"] = int(any(_num(b.get("temperature", 0.0)) > 0 for b in (get("nbody") or [])))
    d["opacity"] = abi.OPACITY[str(get("Opacity", "Lin")).lower()]
    units = units or {}
    d["kappa_const"] = _num(get("KappaConst", 1.0), units.get("opacity"))
    d["kappa_factor"] = float(get("KappaFactor", 1.0))
    d["tau_factor"] = float(get("TauFactor", 0.5))
    d["tau_min"] = float(get("TauMin", 0.01))
    d["density_factor"] = float(get("DensityFactor", (2.0 * 3.141592653589793) ** 0.5))
    d["temperature_cgs"] = float(temp_unit_K)
    d["density_cgs"] = float(units.get("density", 1.0))
    d["opacity_code"] = 1.0 / float(units.get("opacity", 1.0))
    d["energy_density_cgs"] = float(units.get("energy surface density", 1.0))
    d["surface_density_cgs"] = float(units.get("mass surface density", 1.0))
    d["length_cgs"] = float(units.get("length", 1.0))
    d["mass_cgs"] = float(units.get("mass", 1.0))
    d["energy_flux_cgs"] = float(units.get("energy flux", 1.0))
    d["sigma_sb_cgs"] = float(consts.get("sigma_cgs", 0.0))
    d["G_cgs"] = float(consts.get("G_cgs", 0.0))
    d["body_force_from_potential"] = int(_flag(get("BodyForceFromPotential"), True))
    if not d["body_force_from_potential"]:  # SourceEuler.cpp:348-353, 406-413: the kicks would read the acceleration grids
        raise ValueError("BodyForceFromPotential: no is outside this path")
    d["thickness_smoothing"] = float(get("ThicknessSmoothing", 0.6))
    d["imposed_disk_drift"] = float(get("ImposedDiskDrift", 0.0))

    # boundaries: composite names first (boundary_conditions/config.cpp:345-436), then the individual keys, which overwrite what the
    # composite set (get_type, config.cpp:75-94).  The reference infers the INNER energy type from the OUTER side's name, and an explicit
    # InnerBoundaryEnergy overwrites that name too (config.cpp:147): reproduced.
    comp = {"zerogradient": ("zerogradient", "zerogradient", "zerogradient"),
            "outflow": ("zerogradient", "zerogradient", "outflow"),
            "reflecting": ("zerogradient", "zerogradient", "reflecting"),
            "reference": ("reference", "reference", "reference"),
            "viscous": ("zerogradient", "zerogradient", "viscous"),
            "individual": ("", "", "")}
    names = {}
    if get("OuterBoundary") is None:  # Interpret.cpp:290-293
        raise ValueError("OuterBoundary doesn't exist. Old parameter file?")
    for side in ("Inner", "Outer"):
        c = str(get(side + "Boundary", "individual")).lower()
        if c not in comp or (c == "viscous" and side == "Outer"):
            raise ValueError("%sBoundary: %s is outside this path" % (side, c))
        names[side + "Sigma"], names[side + "Energy"], names[side + "Vrad"] = comp[c]

    def get_type(key, name):
        v = get(key)
        if v is not None:
            names[name] = str(v).lower()
        elif not names[name]:
            raise ValueError("Can not infer '%s' when 'InnerBoundary/OuterBoundary' is set to 'individual'" % key)
        return names[name]

    s = [get_type("InnerBoundarySigma", "InnerSigma"), get_type("OuterBoundarySigma", "OuterSigma")]
    e = [get_type("InnerBoundaryEnergy", "OuterEnergy"), get_type("OuterBoundaryEnergy", "OuterEnergy")]  # sic, in this order
    vr = [get_type("InnerBoundaryVrad", "InnerVrad"), get_type("OuterBoundaryVrad", "OuterVrad")]
    for side, name in ((0, "Inner"), (1, "Outer")):
        va = str(get(name + "BoundaryVazi", "keplerian")).lower()
        d.setdefault("bc_sigma", [0, 0])[side] = abi.BC[s[side]]
        d.setdefault("bc_energy", [0, 0])[side] = abi.BC[e[side]]
        d.setdefault("bc_vrad", [0, 0])[side] = abi.BC[vr[side]]
        d.setdefault("bc_vazi", [0, 0])[side] = abi.BC[va]
        d.setdefault("keplerian_azimuthal_factor", [1.0, 1.0])[side] = float(
            get(name + "BoundaryVaziKeplerianFactor", 1.0))
    d["balanced_vazi_sq"] = [0.0, 0.0]  # filled by balanced_vazi_sq() once the radii are known
    d["keplerian_radial_factor"] = [float(get("InnerBoundaryVradKeplerianFactor", 0.1)), float(get("OuterBoundaryVradKeplerianFactor", 0.1))]
    d["viscous_outflow_speed"] = float(get("ViscousOutflowSpeed", 1.0))
    d["correct_disk_selfgravity"] = int(_flag(get("CorrectDiskSelfgravity"), not _flag(get("SelfGravity"), False)))
    d["damping"] = int(_flag(get("Damping"), False))
    d["damping_inner_limit"] = float(get("DampingInnerLimit", 1.05))
    d["damping_outer_limit"] = float(get("DampingOuterLimit", 0.95))
    d["damping_time_factor"] = float(get("DampingTimeFactor", 1.0))
    d["damping_time_radius_outer"] = float(get("DampingTimeRadiusOuter", d["rmax"]))
    for key, name in (("damp_vrad", "VRadial"), ("damp_vazi", "VAzimuthal"), ("damp_sigma", "SurfaceDensity"),
                      ("damp_energy", "Energy")):
        d[key] = [abi.DAMP[str(get("Damping" + name + side, "None")).lower()] for side in ("Inner", "Outer")]
    return d




def balanced_vazi_sq(d, cfg, radii):
    """v_sq of boundary_conditions::balanced_boundary (balanced.cpp:23-52) for the two ghost rings: pow(v_K, 2) x (pressure support
    + smoothing-derivative support), Theo.cpp:122-148; no profile cut-off, no quadrupole term, no self-gravity."""
    import math
    out = []
    for i in (0, len(radii) - 2):
        ri, rs = float(radii[i]), float(radii[i + 1])
        R = 2.0 / 3.0 * (math.pow(rs, 3) - math.pow(ri, 3))
        R = R / (math.pow(rs, 2) - math.pow(ri, 2))
        vk_2 = math.pow(math.sqrt(d["G"] * d["hydro_center_mass"] / R), 2)
        h = d["aspectratio_ref"] * math.pow(R, d["flaring_index"])
        eps = d["thickness_smoothing"]
        support = 0.0
        support += (2.0 * d["flaring_index"] - 1.0 - d["sigma_slope"]) * math.pow(h, 2)
        support += (1.0 + (d["flaring_index"] + 1.0) * math.pow(h * eps, 2)) / math.pow(math.sqrt(1 + math.pow(h * eps, 2)), 3)
        out.append(vk_2 * support)
    return out
