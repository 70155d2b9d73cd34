// kernels_ringsum.cuh — the reference's strictly sequential ring sums, bit for bit, without the sequential chain.
//
// The mean azimuthal velocity of a ring is  sum = 0; for j: sum += v[j]  in both cfl.cpp:199-204 and
// TransportEuler.cpp:215-219; the CFL dt and the integer FARGO shifts follow from it and must be bit-exact, so the
// rounding of EVERY partial sum has to be reproduced.  One thread per ring doing the Ns dependent DADDs
// (k_ring_mean, kernels_ring.cuh) costs Ns x ~20 cycles whatever the slab size: 0.17 ms at Ns = 16384, the largest
// fixed cost of a step once the grid is split over 8 GPUs (33 warps busy on the whole GPU).
//
// k_ring_sum_scan gives a warp to each ring and turns the chain into scans.  While a partial sum S stays inside one
// binade [2^e, 2^(e+1)) every double there is a multiple of u = 2^(e-52), so
//     RN(S + w) = S + RN_u(w)           (RN_u: round w to the nearest multiple of u)
// for every w that is not a tie, and RN_u(w) does not depend on S: it is (w + M) - M with M = 1.5 * 2^e (an exact
// rounding-error computation, valid for |w| < 2^(e-1)).  A tie (|RN_u(w) - w| == u/2) rounds to the even neighbour of
// S + w, which only needs the PARITY of S / u; that parity evolves through the elements as a composition of the maps
// "xor with the parity of the increment" (no tie) and "becomes even" (tie) — an associative operation, i.e. one more
// scan.  Sums of multiples of u inside the binade are exact in FP64 in any order.  So a tile of 32 x RS_TL consecutive
// elements is processed as: round every element independently, scan the parities, fix the ties, scan the sums, and
// find the FIRST "event" — an element too large for the trick or a partial sum leaving (lo, hi) — everything before it
// is exactly what the sequential loop computes; the event element is then added with one plain FP64 addition and the
// remainder of the tile is redone from there (new binade).  A ring of same-sign velocities has about one event per
// binade it crosses (~15 per 16384 elements).  Tiles that keep producing events (partial sums hovering around zero near
// corotation) are finished by the plain sequential loop after RS_MAX_ITERS rounds, and a ring whose last tile needed
// that gives the next tile one round only, so no input costs much more than the chain.
// Sign: the scheme runs on |S| with w = sign(S) * v (negation is exact and RN is sign-symmetric); S == 0, tiny,
// infinite or NaN partial sums take the plain addition element by element.
//
// fargo_selftest_ringsum / tests/test_gpu_math.py compare it with the sequential sum on adversarial rings.
#pragma once
#include "fargo_dev.h"
#include "kernels_ring.cuh"

#ifndef RS_TL
#define RS_TL 8 // elements per lane per tile (lane-major: lane l owns tile elements [l * RS_TL, (l + 1) * RS_TL))
#endif
#define RS_MAX_ITERS 8 // scan rounds per tile before falling back to the plain chain for that tile

__device__ __forceinline__ double rs_pow2(const int e) { return __hiloint2double((e + 1023) << 20, 0); } // 2^e, normal range
// parity maps p -> p ^ b ("xor", code 2 + b) and p -> b ("const", code b); rs_then(f, g) = first f, then g
__device__ __forceinline__ int rs_then(const int f, const int g) { return (g & 2) ? ((f & 2) | ((f ^ g) & 1)) : g; }
__device__ __forceinline__ int rs_apply(const int f, const int p) { return (f & 2) ? ((p ^ f) & 1) : (f & 1); }

// One scan round over the pending elements [pos, tile_n) of a tile held lane-major in v.  Returns the tile index of the
// first event (0x7fffffff: none; then *s_out is the sum after the whole tile) and in *s_out the partial sum just before
// that element, sign restored.  FULL: pos == 0 and the tile is complete (no activity masks).
template <bool FULL>
__device__ __forceinline__ int rs_round(const double (&v)[RS_TL], const double s, const int ebits, const int pos, const int tile_n,
					 const int lane, double *s_out)
{
    const unsigned full = 0xffffffffu;
    const int NONE = 0x7fffffff;
    const double as = fabs(s);
    const int e = ebits - 1023;
    const double lo = rs_pow2(e), hi = rs_pow2(e + 1), big = rs_pow2(e - 1), hu = rs_pow2(e - 53), u = rs_pow2(e - 52);
    const double M = lo + big; // 1.5 * 2^e
    const bool neg = s < 0.0;
    double P[RS_TL];
    unsigned evmask = 0u, tiemask = 0u, upmask = 0u;
    int fk[RS_TL];
    int F = 2; // identity
#pragma unroll
    for (int k = 0; k < RS_TL; ++k) {
	const int t = lane * RS_TL + k;
	const bool active = FULL || ((t >= pos) && (t < tile_n));
	const double w = neg ? -v[k] : v[k];
	const double tt = w + M;
	const double rr = tt - M;
	const double d = rr - w;
	const bool tie = active && (fabs(d) == hu);
	const bool up = d > 0.0;			// a tie that (w + M) rounded upwards: floor is one grid step below
	const int bpar = __double2loint(tt) & 1;	// parity of the rounded increment (M / u is even)
	P[k] = active ? rr : 0.0;
	if (active && !(fabs(w) < big)) // too large for the trick (or NaN)
	    evmask |= 1u << k;
	if (tie)
	    tiemask |= 1u << k;
	if (tie && up)
	    upmask |= 1u << k;
	fk[k] = !active ? 2 : (tie ? 0 : (2 | bpar)); // after a tie the sum is even
	F = rs_then(F, fk[k]);
    }
    if (__any_sync(full, tiemask != 0u)) { // parity scan, then the ties: floor + 1 where (S + floor) is odd
	int G = F;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
	    const int y = __shfl_up_sync(full, G, off);
	    if (lane >= off)
		G = rs_then(y, G);
	}
	int ex = __shfl_up_sync(full, G, 1);
	if (lane == 0)
	    ex = 2;
	int p = rs_apply(ex, __double2loint(as) & 1); // parity of the partial sum in front of this lane's first element
#pragma unroll
	for (int k = 0; k < RS_TL; ++k) {
	    if (tiemask & (1u << k)) {
		const int bpar = __double2loint((neg ? -v[k] : v[k]) + M) & 1;
		const bool isup = (upmask >> k) & 1u;
		const double fl = isup ? P[k] - u : P[k];
		const int flpar = bpar ^ (isup ? 1 : 0);
		P[k] = ((p ^ flpar) & 1) ? fl + u : fl;
	    }
	    p = rs_apply(fk[k], p);
	}
    }
    // in-lane inclusive prefix sums as a tree (every partial sum is a difference of two in-range partial sums of the chain,
    // hence exact, up to and including the lane's first event)
#pragma unroll
    for (int off = 1; off < RS_TL; off <<= 1) {
#pragma unroll
	for (int k = RS_TL - 1; k >= off; --k)
	    P[k] += P[k - off];
    }
    double incl = P[RS_TL - 1];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
	const double y = __shfl_up_sync(full, incl, off);
	if (lane >= off)
	    incl += y;
    }
    // (the exclusive prefix is taken from the lower neighbour's inclusive one: a lane's own total may be inexact)
    const double below = __shfl_up_sync(full, incl, 1);
    const double z0 = (lane == 0) ? as : as + below; // |s| plus everything pending in lower lanes
    unsigned badmask = evmask;
#pragma unroll
    for (int k = 0; k < RS_TL; ++k) {
	const int t = lane * RS_TL + k;
	const double zn = z0 + P[k];
	if ((FULL || ((t >= pos) && (t < tile_n))) && !(zn > lo && zn < hi))
	    badmask |= 1u << k;
    }
    const int first = badmask ? lane * RS_TL + (__ffs(badmask) - 1) : NONE;
    int fmin = first;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
	fmin = min(fmin, __shfl_xor_sync(full, fmin, off));
    if (fmin == NONE) {
	const double zl = __shfl_sync(full, z0 + P[RS_TL - 1], 31);
	*s_out = neg ? -zl : zl;
    } else {
	const int owner = fmin / RS_TL, kk = fmin % RS_TL;
	double zb = z0; // partial sum just before the lane's element kk
#pragma unroll
	for (int k = 1; k < RS_TL; ++k)
	    zb = (kk == k) ? z0 + P[k - 1] : zb;
	zb = __shfl_sync(full, zb, owner);
	*s_out = neg ? -zb : zb;
    }
    return fmin;
}

// mode 0 / 1: as k_ring_mean (CFL mean / transport mean + Nshift + constant residual); mode 2: raw sums into vmean (self-test)
__global__ void __launch_bounds__(128)
    k_ring_sum_scan(const DevView c, const double *__restrict__ vp, double *__restrict__ vmean, int *__restrict__ nshift,
		    double *__restrict__ vconst, const double dt, const int mode, const int nrings, const int ns)
{
    const int lane = threadIdx.x & 31;
    const int ring = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ring >= nrings)
	return; // whole warp
    const double *__restrict__ row = vp + (size_t)ring * ns;
    const bool vec = (ns & 1) == 0;
    const unsigned full = 0xffffffffu;
    const int NONE = 0x7fffffff;
    double s = 0.0;
    double vn[RS_TL]; // next tile (software pipelining: its loads fly while the current tile is scanned)
    auto fetch = [&](const int base, double(&x)[RS_TL]) {
	const int j0 = base + lane * RS_TL;
	if (vec && j0 + RS_TL <= ns) {
#pragma unroll
	    for (int k = 0; k < RS_TL; k += 2) {
		const double2 a = *reinterpret_cast<const double2 *>(row + j0 + k);
		x[k] = a.x;
		x[k + 1] = a.y;
	    }
	} else {
#pragma unroll
	    for (int k = 0; k < RS_TL; ++k)
		x[k] = (j0 + k < ns) ? row[j0 + k] : 0.0;
	}
    };
    fetch(0, vn);
    bool hard = false; // the previous tile had to be finished by the plain chain
    for (int base = 0; base < ns; base += 32 * RS_TL) {
	double v[RS_TL];
#pragma unroll
	for (int k = 0; k < RS_TL; ++k)
	    v[k] = vn[k];
	if (base + 32 * RS_TL < ns)
	    fetch(base + 32 * RS_TL, vn);
	const int tile_n = min(32 * RS_TL, ns - base); // uniform
	const int budget = hard ? 1 : RS_MAX_ITERS;
	hard = false;
	int pos = 0; // tile elements [0, pos) are already in s
	int iter = 0;
	while (pos < tile_n) {
	    const int ebits = (__double2hiint(fabs(s)) >> 20) & 0x7ff;
	    // the scan needs a normal, not tiny, finite |s| (u/2 = 2^(e-53) must be a normal number)
	    if (!((ebits > 60) && (ebits < 0x7ff))) { // one plain addition
		const int owner = pos / RS_TL, kk = pos % RS_TL;
		double x = 0.0;
#pragma unroll
		for (int k = 0; k < RS_TL; ++k)
		    x = (kk == k) ? v[k] : x;
		s += __shfl_sync(full, x, owner);
		++pos;
		continue;
	    }
	    if (iter >= budget) { // events keep coming: the rest of the tile by the plain chain, lane after lane
		for (int l = pos / RS_TL; l * RS_TL < tile_n; ++l) {
		    double sl = s;
#pragma unroll
		    for (int k = 0; k < RS_TL; ++k) {
			const int t = lane * RS_TL + k;
			if (t >= pos && t < tile_n)
			    sl += v[k];
		    }
		    s = __shfl_sync(full, sl, l);
		}
		pos = tile_n;
		hard = true;
		continue;
	    }
	    ++iter;
	    double sb;
	    const int fmin = (pos == 0 && tile_n == 32 * RS_TL) ? rs_round<true>(v, s, ebits, pos, tile_n, lane, &sb)
								: rs_round<false>(v, s, ebits, pos, tile_n, lane, &sb);
	    s = sb;
	    if (fmin == NONE) {
		pos = tile_n;
	    } else { // everything before element fmin is exact; add that element the plain way and go on behind it
		const int owner = fmin / RS_TL, kk = fmin % RS_TL;
		double x = 0.0;
#pragma unroll
		for (int k = 0; k < RS_TL; ++k)
		    x = (kk == k) ? v[k] : x;
		s += __shfl_sync(full, x, owner);
		pos = fmin + 1;
	    }
	}
    }
    if (lane == 0) {
	if (mode == 2)
	    vmean[ring] = s;
	else
	    ring_mean_finish(c, ring, s, vmean, nshift, vconst, dt, mode);
    }
}

// the plain chain on raw rows (self-test reference on the device; the host compares both with its own sequential sum)
__global__ void __launch_bounds__(128) k_ring_sum_chain(const double *__restrict__ x, double *__restrict__ sums, const int nrows, const int ns)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows)
	return;
    const double *row = x + (size_t)r * ns;
    double s = 0.0;
    for (int j = 0; j < ns; ++j)
	s += row[j];
    sums[r] = s;
}
