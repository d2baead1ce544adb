// Are the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64) and the FP64 FMA pipe concurrent on B200?
// Three kernels with the same loop count: DFMA only, DMMA only, both interleaved (RD DFMA per DMMA).  If the pipes are
// independent, t(mixed) ~ max(t_dfma, t_dmma); if they share the datapath, t(mixed) ~ t_dfma + t_dmma.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mix tools/micro/dmma_dfma_mix.cu && /tmp/mix
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF, int NM> __global__ void __launch_bounds__(128) mix(double* out, int iters, double a, double b)
{
    double f[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { f[i] = threadIdx.x * 1e-3 + i; c[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (NM) {
#pragma unroll
                for (int i = 0; i < NM; i++) dmma(c[2 * (i & 3)], c[2 * (i & 3) + 1], a, b);
            }
            if (NF) {
#pragma unroll
                for (int i = 0; i < NF; i++) f[i & 7] = fma(f[i & 7], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += f[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, int NM> float run(double* out, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NF, NM><<<148 * 4, 128>>>(out, 10, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    mix<NF, NM><<<148 * 4, 128>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = 148.0 * 4 * 4, n = (double)iters * 4;
    printf("DFMA/iter %2d  DMMA/iter %2d : %8.3f ms   DFMA %6.2f TFLOP/s   DMMA %6.2f TFLOP/s\n", NF, NM, ms,
           warps * n * NF * 32 * 2 / ms * 1e-9, warps * n * NM * 512 / ms * 1e-9);
    return ms;
}

int main()
{
    double* out;
    cudaMalloc(&out, 148 * 4 * 128 * 8);
    const int iters = 20000;
    run<16, 0>(out, iters);
    run<0, 2>(out, iters);
    run<0, 4>(out, iters);
    run<16, 1>(out, iters);
    run<16, 2>(out, iters);
    run<16, 4>(out, iters);
    run<8, 4>(out, iters);
    run<32, 2>(out, iters);
    return 0;
}
