// FP64 pipe micro-benchmark: DFMA / DMUL+DADD issue rate vs ILP and warps per SM (B200 sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, bool FMA> __global__ void k(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
	for (int i = 0; i < ILP; ++i) {
	    if (FMA) x[i] = fma(x[i], a, b);
	    else x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
	}
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}
template <int ILP, bool FMA> void run(int warps_per_sm, double *d)
{
    int iters = 4096;
    int blocks = 148, threads = warps_per_sm * 32;
    if (threads > 1024) { blocks = 148 * (threads / 1024); threads = 1024; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP, FMA><<<blocks, threads>>>(d, 16, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<ILP, FMA><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ninst = (double)iters * ILP * (FMA ? 1 : 2) * warps_per_sm;   // warp instructions per SM
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%s ILP=%d warps/SM=%2d : %.3f warp-inst/clk/SM  (%.2f clk per dependent step)\n", FMA ? "DFMA     " : "DMUL+DADD", ILP, warps_per_sm,
	   ninst / cyc, cyc / iters / (FMA ? 1 : 2));
}
int main()
{
    double *d; cudaMalloc(&d, 8);
    int ws[] = {4, 8, 12, 16, 32};
    for (int w : ws) { run<1, true>(w, d); run<2, true>(w, d); run<4, true>(w, d); run<8, true>(w, d); }
    for (int w : ws) { run<1, false>(w, d); run<4, false>(w, d); }
    return 0;
}
