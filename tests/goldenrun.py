"""Drive a hydro context (oracle or CUDA) through a golden fixture exactly like the reference's
main()/sim::run did when the fixture was recorded (main.cpp:115-158, simulation.cpp:462-558)."""
import numpy as np

import reftools
from fargocpt_b200 import abi

STATE = ((abi.SIGMA, "Sigma"), (abi.VRAD, "vrad"), (abi.VAZI, "vazi"), (abi.ENERGY, "energy"))


def indirect_term_euler(G, bl):
    """refframe::ComputeIndirectTermNbodyEuler (frame_of_reference.cpp:112-132) over ComputeNbodyOnNbodyAccel
    (Pframeforce.cpp:225-251) for a hydro frame centred on body 0, same operation order (IEEE doubles; math.pow is
    the libm pow the reference's std::pow(dist, 3) resolves to; std::pow(x, 2) is folded to x * x by the compiler)."""
    import math
    mass0, x, y = bl[0][0], bl[0][1], bl[0][2]
    ax = ay = 0.0
    for b in bl[1:]:
        mass, xo, yo = b[0], b[1], b[2]
        dist = math.sqrt((x - xo) * (x - xo) + (y - yo) * (y - yo))
        ax -= G * mass / math.pow(dist, 3) * (x - xo)
        ay -= G * mass / math.pow(dist, 3) * (y - yo)
    ix = iy = 0.0
    mass_center = 0.0
    ix -= mass0 * ax
    iy -= mass0 * ay
    mass_center += mass0
    return ix / mass_center, iy / mass_center


def bodies_at(meta, k, omega_frame):
    bl = meta["bodies"][k]
    indirect = (0.0, 0.0)
    if len(bl) > 1:
        assert int(meta["config"].get("IndirectTermMode", 0)) == 1, "only the Euler indirect term is restated here"
        indirect = indirect_term_euler(meta["consts"]["G"], bl)
        if str(meta["config"].get("DiskFeedback", "yes")).lower()[0] == "y":
            # refframe::ComputeIndirectTermDisk (frame_of_reference.cpp:69-90) from the recorded disk-on-body acceleration
            # of the centre body, + ComputeIndirectTermFully (:166-169)
            # with DiskFeedback the record written at snapshot k+1 holds the acceleration computed at the START of the
            # step k -> k+1 (simulation.cpp:155-156; handle_outputs does not refresh it, simulation.cpp:59-61)
            acc = meta["bodies"][min(k + 1, len(meta["bodies"]) - 1)][0]
            dx = dy = 0.0
            dx -= bl[0][0] * acc[5]
            dy -= bl[0][0] * acc[6]
            dx /= bl[0][0]
            dy /= bl[0][0]
            indirect = (dx + indirect[0], dy + indirect[1])
    masses = [b[0] for b in bl]
    if len(bl) > 1 and str(meta["config"].get("DiskFeedback", "yes")).lower()[0] == "y":
        # an accreting body that feels the disk has swallowed this step's gas (update_planet, accretion.cpp:60-82) before the
        # potential is evaluated (simulation.cpp:150-172): its mass is the one recorded at the END of the step
        nxt = meta["bodies"][min(k + 1, len(meta["bodies"]) - 1)]
        masses = [nxt[i][0] if (len(b) > 7 and b[7] > 0.0) else b[0] for i, b in enumerate(bl)]
    rad = {}
    if len(bl[0]) >= 16:  # fixtures recorded with the irradiation members of the planet records (SourceEuler.cpp:538-564)
        import math
        t = meta["misc"][k]["time"]
        rad = dict(temperature=[b[12] for b in bl], radius=[b[13] for b in bl],
                   ramp=[1.0 - math.pow(math.cos(t * math.pi / 2.0 / b[14]), 2) if t < b[14] else 1.0 for b in bl],
                   rsm=[b[10] * b[9] * b[15] for b in bl])  # Pframeforce.cpp:33-35: l1 * cubic smoothing factor
    return abi.FargoBodies.make([b[1] for b in bl], [b[2] for b in bl], masses, indirect=indirect,
                                omega_frame=omega_frame, **rad)


def start_from_snapshot0(ctx, meta, z):
    for fid, name in STATE:
        ctx.upload(fid, z[name + "_0"])
    omega = float(meta["config"].get("OmegaFrame", 0.0))
    ctx.set_bodies(bodies_at(meta, 0, omega))
    ctx.set_time(0.0)
    if ctx.params.pvte:  # init_eos_arrays (init.cpp:290-292) before init_euler
        ctx.set_pvte(float(meta["config"].get("HydrogenMassFraction", 0.75)))
    ctx.init_derived()
    # the reference evaluates Q+/- for the first CFL before the velocities exist (init.cpp:330-331 ->
    # SourceEuler.cpp:284); take its stored values instead of recomputing them from the full state
    if "Qplus_0" in z and ctx.params.adiabatic:
        ctx.upload(abi.QPLUS, z["Qplus_0"])
        ctx.upload(abi.QMINUS, z["Qminus_0"])
    if ctx.params.pvte and "gammaeff_0" in z:
        # the PVTE grids are state of their own: the reference looked them up BEFORE its first boundary conditions changed the
        # ghost rings of the snapshot-0 energy, and the scale height the next lookup reads was stored then too
        for fid, name in ((abi.GAMMAEFF, "gammaeff"), (abi.MU, "mu"), (abi.GAMMA1, "gamma1"), (abi.SCALE_HEIGHT, "scale_height")):
            ctx.upload(fid, z[name + "_0"])
    if ctx.params.alpha_mode and "Temperature_0" in z:
        # AlphaMode 1 reads the TEMPERATURE grid one refresh late: the reference's dates from init_euler, before the first boundary
        # conditions changed the ghost rings of the snapshot-0 energy
        ctx.upload(abi.TEMPERATURE, z["Temperature_0"])
    ctx.copy_initial_values()
    loop = reftools.TimeLoop(ctx, meta["first_dt"], meta["monitor_timestep"])
    loop.calculate_time_step()  # main.cpp:117
    ctx.stage("boundary", 0.0, 0)  # sim::init, simulation.cpp:463
    loop.calculate_time_step()  # simulation.cpp:467
    return loop, omega


def accretion_inputs(meta, k, body, dt):
    """What accretion::AccreteOntoSinglePlanet (accretion.cpp:104-118) reads off body `body` as recorded at snapshot k, for a
    step of length dt starting there: (x, y, RHill, facc, frac).  The orbital period is t_planet::calculate_orbital_elements'
    (planet.cpp:516-517) from the recorded semi-major axis; std::pow / std::log are the libm functions Python calls."""
    import math
    rec, primary = meta["bodies"][k][body], meta["bodies"][k][0]
    acc_eff, dist_primary, roche, a = rec[7], rec[9], rec[10], rec[11]
    G = meta["consts"]["G"]
    m = primary[0] + rec[0]
    period = 2.0 * math.pi * math.sqrt(math.pow(a, 3) / (m * G))
    facc = dt * acc_eff / period * math.log(2)
    if meta["config"]["nbody"][body].get("accretion method", "kley") == "viscous":
        facc = dt * 3.0 * math.pi * acc_eff  # accretion.cpp:355
    r_hill = roche * dist_primary
    frac = float(meta["config"].get("MassAccretionRadius", 1.0))
    return rec[1], rec[2], r_hill, facc, frac


def run_fixture(ctx, meta, z, nsteps=None, on_snapshot=None, accreted=None):
    """Returns list of per-snapshot dicts of downloaded state fields.  Bodies with an accretion efficiency accrete first
    thing in every step (simulation.cpp:150-153); `accreted` (a list) receives (snapshot, body, dM, dPx, dPy)."""
    loop, omega = start_from_snapshot0(ctx, meta, z)
    nsnap = meta["nsnap"] if nsteps is None else nsteps
    out = []
    accretors = [b for b, rec in enumerate(meta["bodies"][0]) if len(rec) > 7 and rec[7] > 0.0]

    def before_step(k, dt):
        for b in accretors:
            method = meta["config"]["nbody"][b].get("accretion method", "kley")
            d = ctx.accrete_kley(*accretion_inputs(meta, k - 1, b, dt), method=method)
            if accreted is not None:
                accreted.append((k, b) + tuple(d))
        return bodies_at(meta, k - 1, omega)

    for k in range(1, nsnap + 1):
        guard = 0
        while True:
            hit = loop.advance(lambda t, dt: before_step(k, dt))
            guard += 1
            assert guard < 1000
            if hit:
                break
        keep = meta["config"].get("_keep")
        snap = {name: ctx.download(fid) for fid, name in STATE} if (keep is None or k in keep) else {}
        snap["time"], snap["n_iter"], snap["last_dt"] = loop.time, loop.n_iter, loop.last_dt
        out.append(snap)
        if on_snapshot:
            on_snapshot(k, snap)
    return out
