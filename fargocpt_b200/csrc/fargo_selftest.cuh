// fargo_selftest.cuh — device self-test of fargo_math.h: the branch-free fast paths against the plain operators.
#pragma once
#include "fargo_dev.h"
#include "fargo_math.h"

__device__ __forceinline__ unsigned long long st_rng(unsigned long long &s)
{
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return s;
}
__device__ __forceinline__ double st_make(unsigned long long m, int e, int mode)
{
    if (mode == 1)
	m |= 0xFFFFFFFFFF000ull; // significand nearly all ones
    else if (mode == 2)
	m &= 0xFFFull; // significand nearly a power of two
    return __longlong_as_double((long long)(((unsigned long long)(e + 1023) << 52) | (m & 0xFFFFFFFFFFFFFull)));
}
// counts[0] division mismatches, [1] sqrt, [2] exp, [3] pairs whose fast path was declared valid (of n)
__global__ void k_selftest_math(const unsigned long long seed, const int per_thread, const int wide, unsigned long long *counts)
{
    unsigned long long s = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long bad_div = 0, bad_sqrt = 0, bad_exp = 0, nvalid = 0;
    for (int it = 0; it < per_thread; ++it) {
	const int mode = (int)(st_rng(s) & 3);
	const int span = wide ? 2040 : 80; // exponent range: full double range or "physical" magnitudes
	const int ea = (int)(st_rng(s) % span) - span / 2, eb = (int)(st_rng(s) % span) - span / 2;
	double a = st_make(st_rng(s), ea, mode), b = st_make(st_rng(s), eb, (mode + 1) & 3);
	if (st_rng(s) & 1)
	    a = -a;
	if (st_rng(s) & 2)
	    b = -b;
	if (mode == 3 && (it & 7) == 0)
	    a = 0.0;
	{
	    bool ok;
	    const double q = fm_div(a, b, ok);
	    const double qr = a / b;
	    if (ok) {
		++nvalid;
		if (__double_as_longlong(q) != __double_as_longlong(qr))
		    ++bad_div;
	    }
	    FmAcc A;
	    const double q2 = MathP<true>::div(a, b, A);
	    if (fm_acc_ok(A) != ok || (ok && __double_as_longlong(q2) != __double_as_longlong(qr)))
		++bad_div;
	}
	{ // the key-free van Leer limiter (fargo_dev.h:limiter_nb) against the plain operators on the operand range its
	  // callers guarantee — 0 or +-[2^-483, 2^431]: same signs (the branch where the division counts), the operands as
	  // drawn (half of them of opposite sign, some zero), and the 0.5 * limiter form of the azimuthal sweep
	    const bool in_range = (a == 0.0 || (fabs(a) >= 0x1p-483 && fabs(a) <= 0x1p431)) && fabs(b) >= 0x1p-483 && fabs(b) <= 0x1p431;
	    const double bb = (a < 0.0) == (b < 0.0) ? b : -b;
	    const double bs[2] = {bb, b};
	    for (int v = 0; v < 2 && in_range; ++v) {
		const double q = limiter_nb<FARGO_LIMITER_VANLEER>(a, bs[v]);
		const double qh = limiter_nb<FARGO_LIMITER_VANLEER, true>(a, bs[v]);
		const double ref = flux_limiter<FARGO_LIMITER_VANLEER>(a, bs[v]);
		if (__double_as_longlong(q) != __double_as_longlong(ref))
		    ++bad_div;
		if (__double_as_longlong(qh) != __double_as_longlong(0.5 * ref))
		    ++bad_div;
	    }
	}
	{
	    bool ok;
	    const double x = fabs(a);
	    const double r = fm_sqrt(x, ok);
	    if (ok && __double_as_longlong(r) != __double_as_longlong(sqrt(x)))
		++bad_sqrt;
	}
	{
	    unsigned key;
	    const double x = wide ? a : ldexp(a, -ea + (int)(st_rng(s) % 24) - 14); // |x| in [2^-14, 2^10)
	    const double r = fm_exp_raw(x, key);
	    if (key < 0x7ca00000u && __double_as_longlong(r) != __double_as_longlong(exp_ref(x))) // one function, two entry points
		++bad_exp;
	}
    }
    atomicAdd(&counts[0], bad_div);
    atomicAdd(&counts[1], bad_sqrt);
    atomicAdd(&counts[2], bad_exp);
    atomicAdd(&counts[3], nvalid);
}

// exp on caller-provided arguments (the host compares the results with its libm: tests/test_gpu_math.py)
__global__ void k_selftest_exp(const int n, const double *__restrict__ x, double *__restrict__ y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
	y[i] = exp_ref(x[i]);
}
