// FP64 + other-pipe co-issue micro-benchmark (B200 sm_100a): does a DFMA (16 lanes / clk / SM sub-partition, i.e. two
// pipe cycles per warp instruction) leave the sub-partition's issue port free for ALU / FMA-pipe / XU instructions?
// Every thread runs NF independent DFMA chains and NO independent chains of the "other" instruction per iteration.
//   MODE 0: none   1: IADD3 (ALU)   2: FSEL-like select (ALU)   3: IMAD (FMA pipe)   4: FFMA (FMA pipe)   5: SHFL
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NO, int MODE> __global__ void k(double *out, int iters, double a, double b, int ia, float fa)
{
    double x[NF > 0 ? NF : 1];
    x[0] = 0.0;
    int n[NO > 0 ? NO : 1];
    float f[NO > 0 ? NO : 1];
#pragma unroll
    for (int i = 0; i < NF; ++i)
	x[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < NO; ++i) {
	n[i] = threadIdx.x + i;
	f[i] = threadIdx.x * 0.5f + i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
	for (int i = 0; i < (NF > NO ? NF : NO); ++i) {
	    if (i < NF)
		asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
	    if (i < NO) {
		if (MODE == 1)
		    asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(ia));
		if (MODE == 2)
		    asm volatile("{.reg .pred p; setp.gt.s32 p, %1, 0; selp.b32 %0, %0, %1, p;}" : "+r"(n[i]) : "r"(ia));
		if (MODE == 3)
		    asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(n[i]) : "r"(ia));
		if (MODE == 4)
		    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fa));
		if (MODE == 5)
		    asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(n[i]));
	    }
	}
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i)
	s += x[i];
#pragma unroll
    for (int i = 0; i < NO; ++i)
	s += n[i] + f[i];
    if (s == 123.456)
	out[0] = s;
}

template <int NF, int NO, int MODE> void run(int warps_per_sm, double *d, const char *what)
{
    const int iters = 4096;
    const int blocks = 148, threads = warps_per_sm * 32;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<NF, NO, MODE><<<blocks, threads>>>(d, 16, 1.0000001, 1e-9, 3, 1.0001f);
    cudaEventRecord(e0);
    k<NF, NO, MODE><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9, 3, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9;
    const double wps = warps_per_sm / 4.0; // warps per sub-partition
    printf("%-6s NF=%d NO=%2d warps/SMSP=%4.1f : %6.2f clk / iteration / SMSP  -> DFMA %.3f + other %.3f warp-inst/clk/SMSP\n", what, NF, NO,
	   wps, cyc / iters, NF * wps * iters / cyc, NO * wps * iters / cyc);
}

int main()
{
    double *d;
    cudaMalloc(&d, 8);
    for (int w : {4, 12, 16}) {
	run<8, 0, 0>(w, d, "none");
	run<8, 4, 1>(w, d, "iadd");
	run<8, 8, 1>(w, d, "iadd");
	run<8, 16, 1>(w, d, "iadd");
	run<8, 8, 2>(w, d, "sel");
	run<8, 8, 3>(w, d, "imad");
	run<8, 16, 3>(w, d, "imad");
	run<8, 8, 4>(w, d, "ffma");
	run<8, 8, 5>(w, d, "shfl");
	run<0, 16, 1>(w, d, "iadd");
	run<0, 16, 3>(w, d, "imad");
    }
    return 0;
}
