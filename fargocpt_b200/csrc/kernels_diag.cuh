// kernels_diag.cuh — full-grid work the reference does outside the gas update proper but on the same fields:
//   k_reynolds_means / k_reynolds_cells   stress::calculate_Reynolds_stress (stress.cpp:34-70), output diagnostic
//   k_disk_on_body / k_disk_on_body_final ComputeDiskOnPlanetAccel (Force.cpp:23-122), per step with DiskFeedback
#pragma once
#include "fargo_dev.h"

// Which cells credit their accreted gas to the planet: the reference's condition is `radial_first_active < i < radial_active_size`
// (accretion.cpp:186-187, strictly), which at np = 1 leaves out the first active ring next to the inner boundary — and at np > 1
// also the first OWNED ring of every further rank, whose gas is then removed from the disk but credited to nobody: the one place
// where the reference's result depends on np.  Here N GPUs reproduce the np = 1 result: only rank 0 skips its first active ring.
__device__ __forceinline__ bool accrete_counts(const DevView &c, const int i)
{
    return (c.rank == 0 ? c.first_active < i : c.first_active <= i) && i < c.active_size;
}
#include "kernels_source.cuh"

// ring means of the cell-centred velocities, summed strictly in index order like the reference's serial inner loop
// (stress.cpp:48-58).  One thread per ring: a diagnostic evaluated when an output asks for it, not per step.
__global__ void __launch_bounds__(64) k_reynolds_means(const DevView c, const double *__restrict__ vr, const double *__restrict__ vp,
							 double *__restrict__ vr_mean, double *__restrict__ vp_mean)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.nr)
	return;
    double a = 0.0, b = 0.0;
    for (int j = 0; j < c.ns; ++j) {
	const int jn = (j == c.ns - 1) ? 0 : j + 1;
	a += 0.5 * (AT(vr, i, j) + AT(vr, i + 1, j));
	b += 0.5 * (AT(vp, i, j) + AT(vp, i, jn));
    }
    vp_mean[i] = b / (double)c.ns;
    vr_mean[i] = a / (double)c.ns;
}
__global__ void __launch_bounds__(256) k_reynolds_cells(const DevView c, const double *__restrict__ sigma, const double *__restrict__ vr,
							 const double *__restrict__ vp, const double *__restrict__ vr_mean,
							 const double *__restrict__ vp_mean, double *__restrict__ out)
{
    CELL_INDEX(c.nr);
    AT(out, i, j) = AT(sigma, i, j) * (0.5 * (AT(vr, i, j) + AT(vr, i + 1, j)) - vr_mean[i]) *
		    (0.5 * (AT(vp, i, j) + AT(vp, i, jp)) - vp_mean[i]);
}

// ---------------------------------------------------------------------------------------------
// ComputeDiskOnPlanetAccel (Force.cpp:23-122): acceleration of body nb by the gas of the active rings, split into
// the parts from inside / outside the body's orbit {axi, ayi, axo, ayo}.  The reference sums with an OpenMP
// reduction + MPI_Allreduce, i.e. in no defined order; here the order is FIXED (per-thread serial over a column
// strip, shuffle tree, one partial per block, partials added in block order by the final kernel), so the result is
// reproducible from run to run and independent of the launch geometry of other kernels.
// ComputeAverageDensity (Pframeforce.cpp:174-188): ring mean of Sigma, serial sum in index order
__global__ void __launch_bounds__(64) k_sigma_ring_mean(const DevView c, const double *__restrict__ sigma, double *__restrict__ sigma1d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.nr)
	return;
    double sum = 0;
    for (int j = 0; j < c.ns; ++j)
	sum += AT(sigma, i, j);
    sigma1d[i] = sum / c.ns;
}
struct BodyForceIn {
    double x, y, a;	     // position, distance to the origin (planet.get_r())
    double klahr_factor, r_sm; // cubic smoothing factor and l1 * factor (0: off)
};
#define DOB_THREADS 256
__global__ void __launch_bounds__(DOB_THREADS) k_disk_on_body(const DevView c, const double *__restrict__ sigma,
								const double *__restrict__ energy, const double *__restrict__ sigma1d,
								const BodyForceIn B, double *__restrict__ partials /* [gridDim.y * gridDim.x][4] */)
{
    const int i = c.first_active + blockIdx.y;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < c.active_size) {
	const double rmed = c.g.rmed[i], surf = c.g.surf[i];
	const bool inner = rmed < B.a;
	const double s1d = c.p.correct_disk_selfgravity ? sigma1d[i] : 0.0;
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < c.ns; j += gridDim.x * blockDim.x) {
	    const double s = AT(sigma, i, j), e = c.p.adiabatic ? AT(energy, i, j) : 0.0;
	    const size_t cell_ = (size_t)i * c.ns + j;
	    const double smooth = c.p.thickness_smoothing * eos_H_at(c, i, cell_, eos_cs_at(c, i, cell_, s, e)); // compute_smoothing, Force.cpp:124-159
	    const double xc = rmed * c.g.cosphi[j], yc = rmed * c.g.sinphi[j];
	    double cell_sigma = s;
	    if (c.p.correct_disk_selfgravity)
		cell_sigma -= s1d;
	    const double cellmass = surf * cell_sigma;
	    const double dx = xc - B.x, dy = yc - B.y;
	    const double dist_2 = dx * dx + dy * dy;
	    const double dist_sm_2 = dist_2 + smooth * smooth;
	    const double dist_sm = sqrt(dist_sm_2);
	    const double dist_sm_3 = dist_sm_2 * dist_sm;
	    const double inv_dist_sm_3 = 1.0 / dist_sm_3;
	    double smooth_factor_klahr = 1.0;
	    if (B.klahr_factor > 0.0 && dist_sm < B.r_sm)
		smooth_factor_klahr = -(3.0 * pow(dist_sm / B.r_sm, 4.0) - 4.0 * pow(dist_sm / B.r_sm, 3.0));
	    const double fx = c.p.G * cellmass * dx * inv_dist_sm_3 * smooth_factor_klahr;
	    const double fy = c.p.G * cellmass * dy * inv_dist_sm_3 * smooth_factor_klahr;
	    acc[inner ? 0 : 2] += fx;
	    acc[inner ? 1 : 3] += fy;
	}
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	    acc[q] += __shfl_down_sync(0xffffffffu, acc[q], o);
    __shared__ double sh[DOB_THREADS / 32][4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
	for (int q = 0; q < 4; ++q)
	    sh[w][q] = acc[q];
    __syncthreads();
    if (threadIdx.x < 4) {
	double s = 0.0;
	for (int k = 0; k < DOB_THREADS / 32; ++k)
	    s += sh[k][threadIdx.x];
	partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = s;
    }
}
// adds the block partials in block order: 4 warps, one per component, each lane a strided serial sum, then a shuffle tree
__global__ void __launch_bounds__(128) k_disk_on_body_final(const double *__restrict__ partials, const int nblocks, double *__restrict__ out4)
{
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32)
	s += partials[(size_t)b * 4 + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
	s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0)
	out4[q] = s;
}

// ---------------------------------------------------------------------------------------------
// Global disk quantities of monitor/Quantities.dat (quantities.cpp:51-78 mass, :242-276 angular momentum, :281-304
// internal energy, :306-355 viscous dissipation and luminosity, :357-480 kinetic energies; written by
// output::write_quantities, output.cpp:326-520): sums over the active cells with Rmed <= radius_limit.
// The reference adds them with an OpenMP reduction (no defined order); here: per-block partials by warp shuffles, then one
// block adding the partials in block order — a fixed order, so the result is reproducible run to run.
// q: 0 mass, 1 angular momentum, 2 internal energy, 3 kinetic energy, 4 radial kinetic, 5 azimuthal kinetic,
//    6 viscous dissipation (sum Surf Q+), 7 luminosity (sum Surf Q-)
#define MQ_N 8
#define MQ_THREADS 128
__global__ void __launch_bounds__(MQ_THREADS)
    k_monitor_quantities(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy,
			 const double *__restrict__ vr, const double *__restrict__ vp, const double *__restrict__ qplus,
			 const double *__restrict__ qminus, const double radius_limit, double *__restrict__ partials)
{
    const int i = c.first_active + blockIdx.y;
    double acc[MQ_N];
#pragma unroll
    for (int q = 0; q < MQ_N; ++q)
	acc[q] = 0.0;
    if (i < c.active_size && c.g.rmed[i] <= radius_limit) {
	const double rmed = c.g.rmed[i], surf = c.g.surf[i], rinf = c.g.rinf[i], rsup = c.g.rsup[i];
	const double OmegaF = c.b.omega_frame;
	const int ns = c.ns;
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ns; j += gridDim.x * blockDim.x) {
	    const int jm = (j == 0) ? ns - 1 : j - 1, jp = (j == ns - 1) ? 0 : j + 1;
	    const double s = AT(sigma, i, j);
	    acc[0] += surf * s;
	    acc[1] += surf * 0.5 * (s + AT(sigma, i, jm)) * rmed * (AT(vp, i, j) + OmegaF * rmed);
	    if (c.p.adiabatic) {
		acc[2] += surf * AT(energy, i, j);
		acc[6] += surf * AT(qplus, i, j);
		acc[7] += surf * AT(qminus, i, j);
	    }
	    double v_radial_center = (rmed - rinf) * AT(vr, i + 1, j) + (rsup - rmed) * AT(vr, i, j);
	    v_radial_center /= (rsup - rinf);
	    const double v_azimuthal_center = 0.5 * (AT(vp, i, j) + AT(vp, i, jp)) + rmed * OmegaF;
	    acc[3] += 0.5 * surf * s * (v_radial_center * v_radial_center + v_azimuthal_center * v_azimuthal_center);
	    acc[4] += 0.5 * surf * s * (v_radial_center * v_radial_center);
	    acc[5] += 0.5 * surf * s * (v_azimuthal_center * v_azimuthal_center);
	}
    }
#pragma unroll
    for (int q = 0; q < MQ_N; ++q)
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	    acc[q] += __shfl_down_sync(0xffffffffu, acc[q], o);
    __shared__ double sh[MQ_THREADS / 32][MQ_N];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
	for (int q = 0; q < MQ_N; ++q)
	    sh[w][q] = acc[q];
    __syncthreads();
    if (threadIdx.x < MQ_N) {
	double t = 0.0;
	for (int k = 0; k < MQ_THREADS / 32; ++k)
	    t += sh[k][threadIdx.x];
	partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * MQ_N + threadIdx.x] = t;
    }
}
// adds the block partials in block order: one warp per quantity, each lane a strided serial sum, then a shuffle tree
__global__ void __launch_bounds__(32 * MQ_N) k_monitor_final(const double *__restrict__ partials, const int nblocks, double *__restrict__ out)
{
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32)
	s += partials[(size_t)b * MQ_N + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
	s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0)
	out[q] = s;
}

// ---------------------------------------------------------------------------------------------
// The mass-weighted columns of monitor/Quantities.dat (output.cpp:373-423): disk radius (quantities::gas_disk_radius,
// quantities.cpp:191-237), disk eccentricity / periastron (calculate_disk_ecc_vector :481-550 + gas_reduce_mass_average
// :145-182) and the mean aspect ratio (compute_aspectratio mode 0, :784-806).  Per-ring sums, so that the host side of
// fargo_monitor_disk can walk the rings in order like the reference's root does:
// q: 0 ring mass sum(Surf Sigma) of every ring, and over the active cells with Rmed <= radius_limit:
//    1 mass sum(Sigma Surf), 2 sum(e_x m), 3 sum(e_y m) (eccentricity vector rotated by the frame angle), 4 sum(H / Rb m),
//    5 advection torque, 6 viscous torque (gas_torques.cpp:11-115 summed by gas_quantity_reduce, quantities.cpp:80-105, 1000-1018),
//    7 sum(Phi m) of the POTENTIAL grid as stored (output.cpp:413-414), 8 gravitational torque (gas_torques.cpp:122-153)
#define MD_N 9
__global__ void __launch_bounds__(MQ_THREADS)
    k_monitor_disk(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy, const double *__restrict__ vr,
		   const double *__restrict__ vp, const double *__restrict__ pot, const double radius_limit, const double cosF,
		   const double sinF, double *__restrict__ partials)
{
    const int i = blockIdx.y;
    double acc[MD_N];
#pragma unroll
    for (int q = 0; q < MD_N; ++q)
	acc[q] = 0.0;
    const double rmed = c.g.rmed[i], surf = c.g.surf[i];
    const bool in_means = i >= c.first_active && i < c.active_size && rmed <= radius_limit;
    const double OmegaF = c.b.omega_frame;
    const int ns = c.ns;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ns; j += gridDim.x * blockDim.x) {
	const double s = AT(sigma, i, j);
	acc[0] += surf * s;
	if (in_means) {
	    const int jp = (j == ns - 1) ? 0 : j + 1;
	    const size_t cell = (size_t)i * ns + j;
	    const double cell_mass = s * surf;
	    const double total_mass = c.p.hydro_center_mass + s * surf;
	    const double cosa = c.g.cosphi[j], sina = c.g.sinphi[j];
	    const double r_x = rmed * cosa, r_y = rmed * sina;
	    const double dist = sqrt(r_x * r_x + r_y * r_y);
	    const double vrc = 0.5 * (AT(vr, i, j) + AT(vr, i + 1, j));
	    const double vpc = 0.5 * (AT(vp, i, j) + AT(vp, i, jp)) + OmegaF * rmed;
	    const double v_xmed = cosa * vrc - sina * vpc;
	    const double v_ymed = sina * vrc + cosa * vpc;
	    const double jz = r_x * v_ymed - r_y * v_xmed;
	    const double e_x = jz * v_ymed / (c.p.G * total_mass) - r_x / dist;
	    const double e_y = -1.0 * jz * v_xmed / (c.p.G * total_mass) - r_y / dist;
	    const double e = c.p.adiabatic ? AT(energy, i, j) : 0.0;
	    const double H = c.pv.H ? c.pv.H[cell] : eos_H_at(c, i, cell, eos_cs_at(c, i, cell, s, e));
	    acc[1] += cell_mass;
	    acc[2] += (e_x * cosF - e_y * sinF) * cell_mass;
	    acc[3] += (e_y * cosF + e_x * sinF) * cell_mass;
	    acc[4] += H / rmed * cell_mass;
	    { // calculate_advection_torque (gas_torques.cpp:11-43)
		double vr_cell = (rmed - c.g.rinf[i]) * AT(vr, i + 1, j) + (c.g.rsup[i] - rmed) * AT(vr, i, j);
		vr_cell *= c.g.invdiffrsup[i];
		const double vazi_cell = 0.5 * (AT(vp, i, j) + AT(vp, i, jp));
		acc[5] += -(rmed * rmed) * s * vr_cell * vazi_cell;
	    }
	    { // the stored potential: mass-weighted sum and calculate_gravitational_torque (:122-153, BodyForceFromPotential)
		const int jm = (j == 0) ? ns - 1 : j - 1;
		acc[7] += AT(pot, i, j) * cell_mass;
		const double gradphi = (AT(pot, i, jp) - AT(pot, i, jm)) * c.invdphi * 0.5;
		acc[8] += -s * gradphi * surf;
	    }
	    if (i >= 1 && i < c.nr - 1) { // calculate_viscous_torque (:45-115) fills rings 1 .. max_radial - 1
		const int jm = (j == 0) ? ns - 1 : j - 1;
		const double inv_dr = c.g.invdiffrsup[i];
		const double dvr_dphi_top = (AT(vr, i + 1, jp) - AT(vr, i + 1, jm)) * 0.5 * c.invdphi;
		const double dvr_dphi_bot = (AT(vr, i, jp) - AT(vr, i, jm)) * 0.5 * c.invdphi;
		double dvr_dphi = (rmed - c.g.rinf[i]) * dvr_dphi_top + (c.g.rsup[i] - rmed) * dvr_dphi_bot;
		dvr_dphi *= inv_dr;
		const double phi_dot_top = 0.5 * (AT(vp, i + 1, jp) + AT(vp, i + 1, j)) / c.g.rmed[i + 1];
		const double phi_dot = 0.5 * (AT(vp, i, jp) + AT(vp, i, j)) / rmed;
		const double phi_dot_bot = 0.5 * (AT(vp, i - 1, jp) + AT(vp, i - 1, j)) / c.g.rmed[i - 1];
		const double dphi_dot_dr_top = (phi_dot_top - phi_dot) * c.g.invdiffrmed[i + 1];
		const double dphi_dot_dr_bot = (phi_dot - phi_dot_bot) * c.g.invdiffrmed[i];
		double dphi_dot_dr = (rmed - c.g.rinf[i]) * dphi_dot_dr_top + (c.g.rsup[i] - rmed) * dphi_dot_dr_bot;
		dphi_dot_dr *= inv_dr;
		const double nu = eos_nu_at(c, i, cell, s, e);
		acc[6] += -fm_pow3(rmed) * nu * s * (dphi_dot_dr + 1.0 / (rmed * rmed) * dvr_dphi);
	    }
	}
    }
#pragma unroll
    for (int q = 0; q < MD_N; ++q)
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	    acc[q] += __shfl_down_sync(0xffffffffu, acc[q], o);
    __shared__ double sh[MQ_THREADS / 32][MD_N];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
	for (int q = 0; q < MD_N; ++q)
	    sh[w][q] = acc[q];
    __syncthreads();
    if (threadIdx.x < MD_N) {
	double t = 0.0;
	for (int k = 0; k < MQ_THREADS / 32; ++k)
	    t += sh[k][threadIdx.x];
	partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * MD_N + threadIdx.x] = t;
    }
}
// block partials of a ring added in block order into rings[q * nrad_global + imin + i] — only the rings this rank owns
// (write2D's rule), the others stay 0 for the sum over ranks
__global__ void __launch_bounds__(128) k_monitor_disk_rings(const DevView c, const double *__restrict__ partials, const int gx,
							     const int nrad_global, double *__restrict__ rings)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.nr)
	return;
    const bool first = c.rank == 0, last = c.rank == c.nranks - 1;
    const bool owned = i >= (first ? 0 : FARGO_CPUOVERLAP) && i < c.nr - (last ? 0 : FARGO_CPUOVERLAP);
    if (!owned)
	return;
#pragma unroll
    for (int q = 0; q < MD_N; ++q) {
	double t = 0.0;
	for (int b = 0; b < gx; ++b)
	    t += partials[((size_t)i * gx + b) * MD_N + q];
	rings[(size_t)q * nrad_global + c.imin + i] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// ComputeCircumPlanetaryMasses (circumplanetary_mass.cpp:11-51): sum(Surf Sigma) over the active cells whose centre lies inside
// the body's Roche radius; rings [ring_lo, ring_hi) can reach it.  Block partials in k_accrete_final's layout (3 per block).
__global__ void __launch_bounds__(128)
    k_circumplanetary_mass(const DevView c, const double *__restrict__ sigma, const double x, const double y, const double roche_radius,
			   const int ring_lo, double *__restrict__ partials)
{
    const int i = ring_lo + blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0.0;
    if (j < c.ns && i >= c.first_active && i < c.active_size) {
	const double rmed = c.g.rmed[i];
	const double cx = rmed * c.g.cosphi[j], cy = rmed * c.g.sinphi[j];
	const double dist = sqrt((cx - x) * (cx - x) + (cy - y) * (cy - y));
	if (dist < roche_radius)
	    m = c.g.surf[i] * AT(sigma, i, j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
	m += __shfl_down_sync(0xffffffffu, m, o);
    __shared__ double sh[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
	sh[w] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
	double *p = partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3;
	p[0] = (sh[0] + sh[1]) + (sh[2] + sh[3]);
	p[1] = p[2] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// accretion::AccreteOntoSinglePlanet (accretion.cpp:84-221, "kley" accretion onto a planet): see fargo_b200.h.
// One thread per cell of the rings [ring_lo, ring_hi) that can reach into the accretion radius; the cells inside it are
// changed in place exactly as the reference changes them (same operations, same order); mass and momentum taken from active
// cells are summed per block (shuffles) and then in block order (k_monitor_final's scheme).
struct AccretionIn {
    double x, y, r_hill, facc1, facc2, frac1, frac2, density_floor;
    int ring_lo, ring_hi;
};
#define ACC_THREADS 128
__global__ void __launch_bounds__(ACC_THREADS)
    k_accrete_kley(const DevView c, double *__restrict__ sigma, double *__restrict__ energy, const double *__restrict__ vr,
		   const double *__restrict__ vp, const AccretionIn a, double *__restrict__ partials)
{
    const int i = a.ring_lo + blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[3] = {0.0, 0.0, 0.0};
    if (i < a.ring_hi && j < c.ns) {
	const double rmed = c.g.rmed[i];
	const double xc = rmed * c.g.cosphi[j], yc = rmed * c.g.sinphi[j];
	const double dx = a.x - xc, dy = a.y - yc;
	const double distance = sqrt(dx * dx + dy * dy);
	if (distance < a.frac1 * a.r_hill) {
	    const int jp = (j == c.ns - 1) ? 0 : j + 1;
	    const double vtcell = 0.5 * (AT(vp, i, j) + AT(vp, i, jp)) + rmed * c.b.omega_frame;
	    const double vrcell = 0.5 * (AT(vr, i, j) + AT(vr, i + 1, j));
	    const double vxcell = (vrcell * xc - vtcell * yc) / rmed;
	    const double vycell = (vrcell * yc + vtcell * xc) / rmed;
	    double s = AT(sigma, i, j);
	    double e = c.p.adiabatic ? AT(energy, i, j) : 0.0;
	    const double facc_max = 1 - a.density_floor / s;
	    const bool active = accrete_counts(c, i);
	    {
		const double facc_ceil = stdmin(a.facc1, facc_max);
		const double deltaM = facc_ceil * s * c.g.surf[i];
		s *= 1.0 - facc_ceil;
		e *= 1.0 - facc_ceil;
		if (active) {
		    acc[1] += deltaM * vxcell;
		    acc[2] += deltaM * vycell;
		    acc[0] += deltaM;
		}
	    }
	    if (distance < a.frac2 * a.r_hill) {
		const double facc_ceil = stdmin(a.facc2, facc_max);
		const double deltaM = facc_ceil * s * c.g.surf[i];
		s *= 1.0 - facc_ceil;
		e *= 1.0 - a.facc2;
		if (active) {
		    acc[1] += deltaM * vxcell;
		    acc[2] += deltaM * vycell;
		    acc[0] += deltaM;
		}
	    }
	    AT(sigma, i, j) = s;
	    if (c.p.adiabatic)
		AT(energy, i, j) = e;
	}
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	    acc[q] += __shfl_down_sync(0xffffffffu, acc[q], o);
    __shared__ double sh[ACC_THREADS / 32][3];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
	for (int q = 0; q < 3; ++q)
	    sh[w][q] = acc[q];
    __syncthreads();
    if (threadIdx.x < 3) {
	double t = 0.0;
	for (int k = 0; k < ACC_THREADS / 32; ++k)
	    t += sh[k][threadIdx.x];
	partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
    }
}
// accretion::AccreteOntoSinglePlanetViscous (accretion.cpp:335-417, "accretion method: viscous"): one zone of radius
// d_max = frac * RHill; a cell loses the fraction facc * nu * 3 / (pi d_max^2) * (1 - distance / d_max), nu being the VISCOSITY
// grid the previous step stored — i.e. nu of the state before any accretion of this step.  This path stores no derived fields:
// nu is evaluated from the kept pre-accretion rows (fargo_dev.h:PreState; accrete_zones keeps this body's band before the
// launch, so every ring of the launch is inside it).  a.facc1 = dt * 3 pi * efficiency, a.frac1 = frac; a.facc2 = f_const,
// a.frac2 = d_max (formed on the host with the reference's expressions).
__global__ void __launch_bounds__(ACC_THREADS)
    k_accrete_viscous(const DevView c, double *__restrict__ sigma, double *__restrict__ energy, const double *__restrict__ vr,
		      const double *__restrict__ vp, const AccretionIn a, const PreState pre, double *__restrict__ partials)
{
    const int i = a.ring_lo + blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[3] = {0.0, 0.0, 0.0};
    if (i < a.ring_hi && j < c.ns) {
	const double rmed = c.g.rmed[i];
	const double xc = rmed * c.g.cosphi[j], yc = rmed * c.g.sinphi[j];
	const double dx = a.x - xc, dy = a.y - yc;
	const double distance = sqrt(dx * dx + dy * dy);
	if (distance < a.frac1 * a.r_hill) {
	    const int jp = (j == c.ns - 1) ? 0 : j + 1;
	    const bool kept = pre_has(pre, i);
	    const double s_nu = kept ? AT(pre.sigma, i, j) : AT(sigma, i, j);
	    const double e_nu = c.p.adiabatic ? (kept ? AT(pre.energy, i, j) : AT(energy, i, j)) : 0.0;
	    const double nu = eos_nu_at(c, i, (size_t)i * c.ns + j, s_nu, e_nu);
	    const double spread = a.facc2 * (1.0 - distance / a.frac2);
	    const double vtcell = 0.5 * (AT(vp, i, j) + AT(vp, i, jp)) + rmed * c.b.omega_frame;
	    const double vrcell = 0.5 * (AT(vr, i, j) + AT(vr, i + 1, j));
	    const double vxcell = (vrcell * xc - vtcell * yc) / rmed;
	    const double vycell = (vrcell * yc + vtcell * xc) / rmed;
	    double s = AT(sigma, i, j);
	    const double facc_max = 1 - a.density_floor / s;
	    const double facc_tmp = a.facc1 * nu * spread;
	    const double facc_ceil = stdmin(facc_tmp, facc_max);
	    const double deltaM = facc_ceil * s * c.g.surf[i];
	    AT(sigma, i, j) = s * (1.0 - facc_ceil);
	    if (c.p.adiabatic)
		AT(energy, i, j) = AT(energy, i, j) * (1.0 - facc_ceil);
	    if (accrete_counts(c, i)) {
		acc[1] += deltaM * vxcell;
		acc[2] += deltaM * vycell;
		acc[0] += deltaM;
	    }
	}
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	    acc[q] += __shfl_down_sync(0xffffffffu, acc[q], o);
    __shared__ double sh[ACC_THREADS / 32][3];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
	for (int q = 0; q < 3; ++q)
	    sh[w][q] = acc[q];
    __syncthreads();
    if (threadIdx.x < 3) {
	double t = 0.0;
	for (int k = 0; k < ACC_THREADS / 32; ++k)
	    t += sh[k][threadIdx.x];
	partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
    }
}
__global__ void __launch_bounds__(96) k_accrete_final(const double *__restrict__ partials, const int nblocks, double *__restrict__ out3)
{
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32)
	s += partials[(size_t)b * 3 + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
	s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0)
	out3[q] = s;
}
