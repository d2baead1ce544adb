// measures the relative error of MUFU.RCP64H (rcp.approx.ftz.f64) and of the Newton refinements built on it
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double* x, double* e0, double* e3, double* e5, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = x[i], r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double exact = 1.0 / a;
    e0[i] = fabs(r - exact) / exact;
    double e = fma(-a, r, 1.0);
    e = fma(e, e, e);
    double r1 = fma(r, e, r);
    e3[i] = fabs(r1 - exact) / exact;
    e = fma(-a, r1, 1.0);
    double r2 = fma(r1, e, r1);
    e5[i] = fabs(r2 - exact) / exact;
}
int main()
{
    const int n = 1 << 22;
    double *x, *e0, *e3, *e5;
    cudaMallocManaged(&x, n * 8); cudaMallocManaged(&e0, n * 8); cudaMallocManaged(&e3, n * 8); cudaMallocManaged(&e5, n * 8);
    for (int i = 0; i < n; i++) x[i] = ldexp(1.0 + (double)i / n, (i % 61) - 30) * (1.0 + 1e-9 * (i % 977));
    k<<<(n + 255) / 256, 256>>>(x, e0, e3, e5, n);
    cudaDeviceSynchronize();
    double m0 = 0, m3 = 0, m5 = 0;
    for (int i = 0; i < n; i++) { m0 = fmax(m0, e0[i]); m3 = fmax(m3, e3[i]); m5 = fmax(m5, e5[i]); }
    printf("max rel err: RCP64H seed %.3e (2^%.1f), +cubic step %.3e, +quadratic step %.3e\n", m0, log2(m0), m3, m5);
    return 0;
}
