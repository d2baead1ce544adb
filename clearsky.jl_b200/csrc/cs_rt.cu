// cs_rt.cu -- K6 fused optical-depth quadrature + Schwarzschild sweeps, K7 spectral reduction.
//
// Replaces, for every wavenumber at once, the body of the reference's threaded loop
//     d-depth!(tau_j, P, T, mu, A, j, C, nlobatto); d-monoflux!(M+_j, M-_j, tau_j, P, B_j, nu_j, fS, fa, theta_s, nstream)
// (src/fluxes.jl:270-277; src/core/discretized.jl:136-177, 249-326), planckevaluations
// (core/discretized.jl:46-58, src/radiation.jl:48-54) and integral-F! (src/core/shared.jl:125-137).
//
// One thread owns one wavenumber.  Sigma at the Lobatto nodes is read coalesced from the workspace
// ([node][nu], nu fastest); the per-layer tau and per-level Planck values needed again by the upward sweep go
// through an L2-resident scratch ([layer][nu]); all streams advance together in registers.  The spectral
// integral is fused: each level's monochromatic flux is multiplied by the point's trapezoid weight and reduced
// with warp shuffles, then across warps in shared memory, and finally across CTAs by a second, fixed-order
// kernel (run-to-run bit-stable; no atomics).
#include "cs_internal.cuh"
#include <cstring>
#include <algorithm>
#include <cmath>

namespace {

constexpr int RT_THREADS = 128;
constexpr int RT_WARPS = RT_THREADS / 32;

struct RtArgs {
    const double* sig;     // [nnode][nnu]
    const double* nu;      // [nnu]
    const double* w;       // [nnu] trapezoid weights
    const double* fS;      // [nnu] or null (== 0)
    const double* fa;      // [nnu] or null (== 0)
    const double* small;   // packed: P[np], mu[nlob*(np-1)], wl[nlob], 1/(kT)[np], m[ns], W[ns], 1/m[ns]
    int64_t nnu;
    int np, nlob, ns;
    double Cg, cos_s;
    double tau_floor;      // floor on the vertical layer depth (1e-6 in the reference, discretized.jl:174)
    double* tau_s;         // scratch [np-1][nnu]
    double* B_s;           // scratch [np][nnu]
    double* tau_out;       // optional, Julia tau[i,j] -> i + (np-1)*j
    double* Mup_out;       // optional, Julia M[i,j]  -> i + np*j
    double* Mdn_out;
    double* part;          // [nblocks][2][np] CTA partial sums (0: up, 1: down); [nblocks*RT_WARPS][2][np] when red_global
    int red_global;        // 1: per-warp partial sums go straight to `part` (no shared-memory accumulators: lifts the level cap)
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// layerplanck (core/discretized.jl:85): B2 (1-t) - (B1-B2) t + (1-t)(B1-B2)/tau, with 1/tau supplied by the caller
// (one reciprocal per layer shared by all streams: 1/(tau m_k) = (1/tau)(1/m_k); differences are at the 1e-16 level)
__device__ __forceinline__ double layerplanck(double B1, double B2, double rtau, double t)
{
    double omt = 1.0 - t, dB = B1 - B2;
    return fma(omt * dB, rtau, fma(B2, omt, -dB * t));
}

template <int NS>
__global__ void __launch_bounds__(RT_THREADS) rt_kernel(RtArgs a)
{
    extern __shared__ double sm[];
    const int np = a.np, L = np - 1, nlob = a.nlob;
    const int ns = (NS > 0) ? NS : a.ns;
    // shared layout: small arrays, then per-warp partial sums [2][np][RT_WARPS]
    const int nsmall = np + nlob * L + nlob + np + 3 * ns;
    double* sP = sm;
    double* smu = sP + np;
    double* swl = smu + nlob * L;
    double* skT = swl + nlob;
    double* sm_m = skT + np;
    double* sW = sm_m + ns;
    double* srm = sW + ns;      // 1/m_k
    double* red = sm + nsmall;
    for (int t = threadIdx.x; t < nsmall; t += RT_THREADS) sm[t] = a.small[t];
    if (!a.red_global)
        for (int t = threadIdx.x; t < 2 * np * RT_WARPS; t += RT_THREADS) red[t] = 0.0;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a warp contributes exactly one partial sum per (direction, level): in shared memory (summed over the warps below) or,
    // for level counts whose accumulators would not fit, straight in the second stage's input
    double* gred = a.part + ((size_t)blockIdx.x * RT_WARPS + warp) * 2 * np;
    auto put = [&](int idx, double r) {
        if (lane == 0) {
            if (a.red_global) gred[idx] = r;
            else red[idx * RT_WARPS + warp] += r;
        }
    };
    const int64_t jraw = (int64_t)blockIdx.x * RT_THREADS + threadIdx.x;
    const bool live = jraw < a.nnu;
    const int64_t j = live ? jraw : a.nnu - 1;
    const int64_t nnu = a.nnu;
    const double wj = live ? a.w[j] : 0.0;
    const double nuj = a.nu[j];
    // planck (radiation.jl:48-54), same operation order, no expm1
    const double num = 100.0 * nuj;
    const double hcn = CS_H * CS_C * num;
    const double pref = 2 * CS_H * (CS_C * CS_C) * (num * num * num);
    const double Cg = a.Cg;
    const double c = a.cos_s;
    const double rc = 1.0 / c;

    constexpr int NSMAX = (NS > 0) ? NS : CS_MAX_STREAMS;
    double I[NSMAX];
#pragma unroll
    for (int k = 0; k < NSMAX; k++) I[k] = 0.0;

    // ---- downward: optical depth of each layer on the fly, diffuse streams, stellar beam
    double beta1 = Cg * (a.sig[j] / smu[0]);
    double Bprev = 100.0 * pref / (exp(hcn * skT[0]) - 1.0);
    a.B_s[j] = Bprev;
    double beam = c * (a.fS ? a.fS[j] : 0.0);
    double Mdn = beam;                                   // M-[1] = c*fS(nu)   (discretized.jl:299)
    if (a.Mdn_out && live) a.Mdn_out[(size_t)np * j] = Mdn;
    {
        put(np + 0, warp_sum(wj * Mdn));
    }
    // the end-node cross-section of the NEXT layer is loaded one iteration ahead: the load is the only long-latency
    // operation of the loop and everything after it depends on it
    double sig_end = a.sig[(size_t)(nlob - 1) * nnu + j];
    for (int i = 0; i < L; i++) {
        double dP = sP[i + 1] - sP[i];
        double ti = 0.0;
        ti += (dP * swl[0]) * beta1;
        const double sig_cur = sig_end;
        if (i + 1 < L) sig_end = a.sig[(size_t)((nlob - 1) * (i + 2)) * nnu + j];
        for (int n = 1; n < nlob - 1; n++) {
            double bn = Cg * (a.sig[(size_t)(n + (nlob - 1) * i) * nnu + j] / smu[n + nlob * i]);
            ti += (dP * swl[n]) * bn;
        }
        double bn = Cg * (sig_cur / smu[(nlob - 1) + nlob * i]);
        ti += (dP * swl[nlob - 1]) * bn;
        beta1 = bn;
        double tau = fmax(ti, a.tau_floor);              // floor on the vertical depth (discretized.jl:174)
        a.tau_s[(size_t)i * nnu + j] = tau;
        if (a.tau_out && live) a.tau_out[(size_t)L * j + i] = tau;
        double Bnext = 100.0 * pref / (exp(hcn * skT[i + 1]) - 1.0);
        a.B_s[(size_t)(i + 1) * nnu + j] = Bnext;
        double Msum = 0.0;
        const double rtau = cs_rcp(tau);
#pragma unroll
        for (int k = 0; k < NSMAX; k++) {
            if (k < ns) {
                double tk = tau * sm_m[k];
                double tr = exp(-tk);
                double Be = layerplanck(Bprev, Bnext, rtau * srm[k], tr);
                I[k] = I[k] * tr + Be;
                Msum += sW[k] * I[k];
            }
        }
        beam *= exp(-tau * rc);                          // discretized.jl:302
        Mdn = Msum + beam;
        if (a.Mdn_out && live) a.Mdn_out[(size_t)np * j + i + 1] = Mdn;
        put(np + i + 1, warp_sum(wj * Mdn));
        Bprev = Bnext;
    }
    // ---- surface: Lambertian reflection + emission (discretized.jl:309-310)
    const double Is = Mdn * (a.fa ? a.fa[j] : 0.0) / CS_PI + Bprev;
    {
        double Mup = Is * CS_PI;
        if (a.Mup_out && live) a.Mup_out[(size_t)np * j + L] = Mup;
        put(L, warp_sum(wj * Mup));
    }
    // ---- upward
#pragma unroll
    for (int k = 0; k < NSMAX; k++) I[k] = Is;
    double B1 = Bprev;
    double tau_n = a.tau_s[(size_t)(L - 1) * nnu + j], B_n = a.B_s[(size_t)(L - 1) * nnu + j];
    for (int i = L - 1; i >= 0; i--) {
        const double tau = tau_n, B2 = B_n;
        if (i > 0) { tau_n = a.tau_s[(size_t)(i - 1) * nnu + j]; B_n = a.B_s[(size_t)(i - 1) * nnu + j]; }
        double Msum = 0.0;
        const double rtau = cs_rcp(tau);
#pragma unroll
        for (int k = 0; k < NSMAX; k++) {
            if (k < ns) {
                double tk = tau * sm_m[k];
                double tr = exp(-tk);
                double Be = layerplanck(B1, B2, rtau * srm[k], tr);
                I[k] = I[k] * tr + Be;
                Msum += sW[k] * I[k];
            }
        }
        if (a.Mup_out && live) a.Mup_out[(size_t)np * j + i] = Msum;
        put(i, warp_sum(wj * Msum));
        B1 = B2;
    }
    if (a.red_global) return;
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * np; t += RT_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < RT_WARPS; q++) s += red[t * RT_WARPS + q];
        a.part[(size_t)blockIdx.x * 2 * np + t] = s;
    }
}

// ---- batched K6: NB temperature profiles over the SAME optical depths (Sigma does not depend on T for an
// AcceleratedAbsorber, absorbers.jl:203; jacobian! is np+1 such solves, radiative_convective.jl:154-171).  The layer
// depths and the (2 ns + 1) transmittances per layer -- all the exponentials except Planck's -- are computed once and
// shared by the NB recurrences a thread carries in registers.  Only the spectrally integrated fluxes are produced.
constexpr int RTB_NB = 8;
struct RtBatchArgs {
    RtArgs a;
    int nb;                // profiles in this launch (<= RTB_NB)
    const double* rkT;     // [nb][np]  1/(k T) per profile and level
    double* B_b;           // scratch [RTB_NB][np][nnu]
    double* part_b;        // [nblocks][nb][2][np]
};

template <int NS>
__global__ void __launch_bounds__(RT_THREADS) rt_batch_kernel(RtBatchArgs ba)
{
    extern __shared__ double sm[];
    const RtArgs& a = ba.a;
    const int np = a.np, L = np - 1, nlob = a.nlob, nb = ba.nb;
    const int ns = (NS > 0) ? NS : a.ns;
    const int nsmall = np + nlob * L + nlob + np + 3 * ns;
    double* sP = sm;
    double* smu = sP + np;
    double* swl = smu + nlob * L;
    double* sm_m = swl + nlob + np;
    double* sW = sm_m + ns;
    double* srm = sW + ns;
    double* skT = sm + nsmall;                 // [nb][np]
    double* red = skT + RTB_NB * np;           // [RTB_NB][2][np][RT_WARPS]
    for (int t = threadIdx.x; t < nsmall; t += RT_THREADS) sm[t] = a.small[t];
    for (int t = threadIdx.x; t < RTB_NB * np; t += RT_THREADS) skT[t] = t < nb * np ? ba.rkT[t] : ba.rkT[t % np];
    for (int t = threadIdx.x; t < RTB_NB * 2 * np * RT_WARPS; t += RT_THREADS) red[t] = 0.0;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t jraw = (int64_t)blockIdx.x * RT_THREADS + threadIdx.x;
    const bool live = jraw < a.nnu;
    const int64_t j = live ? jraw : a.nnu - 1;
    const int64_t nnu = a.nnu;
    const double wj = live ? a.w[j] : 0.0;
    const double nuj = a.nu[j];
    const double num = 100.0 * nuj;
    const double hcn = CS_H * CS_C * num;
    const double pref = 2 * CS_H * (CS_C * CS_C) * (num * num * num);
    const double Cg = a.Cg;
    const double c = a.cos_s;
    const double rc = 1.0 / c;
    constexpr int NSMAX = (NS > 0) ? NS : CS_MAX_STREAMS;
    auto planck = [&](int b, int lev) { return 100.0 * pref / (exp(hcn * skT[b * np + lev]) - 1.0); };
    auto RED = [&](int b, int dir, int lev) -> double& { return red[((b * 2 + dir) * np + lev) * RT_WARPS + warp]; };

    double I[RTB_NB][NSMAX], Bp[RTB_NB];
#pragma unroll
    for (int b = 0; b < RTB_NB; b++) {
#pragma unroll
        for (int k = 0; k < NSMAX; k++) I[b][k] = 0.0;
        Bp[b] = planck(b, 0);
        ba.B_b[((size_t)b * np) * nnu + j] = Bp[b];
    }
    double beta1 = Cg * (a.sig[j] / smu[0]);
    double beam = c * (a.fS ? a.fS[j] : 0.0);
    {
        double r = warp_sum(wj * beam);
        if (lane == 0)
            for (int b = 0; b < nb; b++) RED(b, 1, 0) += r;
    }
    double Mdn[RTB_NB];
    for (int i = 0; i < L; i++) {
        double dP = sP[i + 1] - sP[i];
        double ti = (dP * swl[0]) * beta1;
        for (int n = 1; n < nlob - 1; n++) {
            double bn = Cg * (a.sig[(size_t)(n + (nlob - 1) * i) * nnu + j] / smu[n + nlob * i]);
            ti += (dP * swl[n]) * bn;
        }
        double bn = Cg * (a.sig[(size_t)((nlob - 1) * (i + 1)) * nnu + j] / smu[(nlob - 1) + nlob * i]);
        ti += (dP * swl[nlob - 1]) * bn;
        beta1 = bn;
        double tau = fmax(ti, a.tau_floor);
        a.tau_s[(size_t)i * nnu + j] = tau;
        const double rtau = cs_rcp(tau);
        double Bn[RTB_NB], dB[RTB_NB];
#pragma unroll
        for (int b = 0; b < RTB_NB; b++) {
            Bn[b] = planck(b, i + 1);
            ba.B_b[((size_t)b * np + i + 1) * nnu + j] = Bn[b];
            dB[b] = Bp[b] - Bn[b];
            Mdn[b] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < NSMAX; k++) {
            if (k < ns) {
                // layerplanck(B1,B2) = B2 (1-t) + (B1-B2) ((1-t)/tau_k - t): the two stream factors are shared by all profiles
                const double tr = exp(-(tau * sm_m[k]));
                const double c1 = 1.0 - tr, c2 = fma(c1, rtau * srm[k], -tr);
#pragma unroll
                for (int b = 0; b < RTB_NB; b++) {
                    I[b][k] = fma(I[b][k], tr, fma(Bn[b], c1, dB[b] * c2));
                    Mdn[b] = fma(sW[k], I[b][k], Mdn[b]);
                }
            }
        }
        beam *= exp(-tau * rc);
#pragma unroll
        for (int b = 0; b < RTB_NB; b++) {
            Mdn[b] += beam;
            double r = warp_sum(wj * Mdn[b]);
            if (lane == 0 && b < nb) RED(b, 1, i + 1) += r;
            Bp[b] = Bn[b];
        }
    }
    const double alb = a.fa ? a.fa[j] : 0.0;
#pragma unroll
    for (int b = 0; b < RTB_NB; b++) {
        const double Is = Mdn[b] * alb / CS_PI + Bp[b];
        double r = warp_sum(wj * (Is * CS_PI));
        if (lane == 0 && b < nb) RED(b, 0, L) += r;
#pragma unroll
        for (int k = 0; k < NSMAX; k++) I[b][k] = Is;
    }
    for (int i = L - 1; i >= 0; i--) {
        const double tau = a.tau_s[(size_t)i * nnu + j];
        const double rtau = cs_rcp(tau);
        double Bn[RTB_NB], Ms[RTB_NB], dB[RTB_NB];
#pragma unroll
        for (int b = 0; b < RTB_NB; b++) {
            Bn[b] = ba.B_b[((size_t)b * np + i) * nnu + j];
            dB[b] = Bp[b] - Bn[b];
            Ms[b] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < NSMAX; k++) {
            if (k < ns) {
                const double tr = exp(-(tau * sm_m[k]));
                const double c1 = 1.0 - tr, c2 = fma(c1, rtau * srm[k], -tr);
#pragma unroll
                for (int b = 0; b < RTB_NB; b++) {
                    I[b][k] = fma(I[b][k], tr, fma(Bn[b], c1, dB[b] * c2));
                    Ms[b] = fma(sW[k], I[b][k], Ms[b]);
                }
            }
        }
#pragma unroll
        for (int b = 0; b < RTB_NB; b++) {
            double r = warp_sum(wj * Ms[b]);
            if (lane == 0 && b < nb) RED(b, 0, i) += r;
            Bp[b] = Bn[b];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nb * 2 * np; t += RT_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < RT_WARPS; q++) s += red[t * RT_WARPS + q];
        ba.part_b[(size_t)blockIdx.x * nb * 2 * np + t] = s;
    }
}

// K7 second stage: fixed-order sum over CTAs; F[0..np) = up, F[np..2np) = down.  One warp per output: lanes
// stride over the CTAs in a fixed pattern and finish with a fixed shuffle tree -> run-to-run bit-stable.
__global__ void __launch_bounds__(128) flux_reduce_kernel(const double* __restrict__ part, int nblocks, int n2, double* F)
{
    const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= n2) return;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += part[(size_t)b * n2 + t];
    s = warp_sum(s);
    if (lane == 0) F[t] = s;
}

// total slant optical depth, d-depth (core/discretized.jl:92-134): no floor, tau += tau_i*m per layer
__global__ void __launch_bounds__(RT_THREADS) depth_kernel(const double* sig, const double* small, int64_t nnu, int np,
                                                           int nlob, double Cg, double mfac, double* out)
{
    extern __shared__ double sm[];
    const int L = np - 1;
    const int nsmall = np + nlob * L + nlob;
    for (int t = threadIdx.x; t < nsmall; t += RT_THREADS) sm[t] = small[t];
    __syncthreads();
    const double* sP = sm;
    const double* smu = sP + np;
    const double* swl = smu + nlob * L;
    int64_t j = (int64_t)blockIdx.x * RT_THREADS + threadIdx.x;
    if (j >= nnu) return;
    double beta1 = Cg * (sig[j] / smu[0]);
    double acc = 0.0;
    for (int i = 0; i < L; i++) {
        double dP = sP[i + 1] - sP[i];
        double ti = 0.0;
        ti += (dP * swl[0]) * beta1;
        for (int n = 1; n < nlob - 1; n++) {
            double bn = Cg * (sig[(size_t)(n + (nlob - 1) * i) * nnu + j] / smu[n + nlob * i]);
            ti += (dP * swl[n]) * bn;
        }
        double bn = Cg * (sig[(size_t)((nlob - 1) * (i + 1)) * nnu + j] / smu[(nlob - 1) + nlob * i]);
        ti += (dP * swl[nlob - 1]) * bn;
        beta1 = bn;
        acc += ti * mfac;
    }
    out[j] = acc;
}

template <int NS> int32_t launch_rt(cs_ctx* ctx, const RtArgs& a, int nblocks, size_t smem)
{
    if (smem > 48 * 1024)
        CS_CUDA(cudaFuncSetAttribute(rt_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rt_kernel<NS><<<nblocks, RT_THREADS, smem, ctx->stream>>>(a);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    return CS_OK;
}

int32_t check_profile_args(cs_sigma* s, int64_t np, const double* P, int32_t nlob)
{
    CS_REQUIRE(np >= 2, CS_ERR_ARG, "need at least two pressure levels");
    CS_REQUIRE(nlob >= 2 && nlob <= CS_MAX_LOBATTO, CS_ERR_ARG, "nlobatto must be in [2,%d]", CS_MAX_LOBATTO);
    CS_REQUIRE(s->nnode == (np - 1) * (nlob - 1) + 1, CS_ERR_ARG,
               "sigma workspace has %lld nodes, expected (np-1)*(nlobatto-1)+1 = %lld", (long long)s->nnode,
               (long long)((np - 1) * (nlob - 1) + 1));
    // issorted(P) (fluxes.jl:257)
    for (int64_t i = 1; i < np; i++)
        CS_REQUIRE(P[i] >= P[i - 1], CS_ERR_ARG, "pressure coordinates must be in ascending order (sorted)");
    return CS_OK;
}

int32_t fluxes_impl(cs_sigma* s, int64_t np, const double* P, int32_t nlob, const double* wlob, const double* mu,
                    const double* Tlev, double g, const double* fS, const double* fa, double theta_s,
                    int32_t nstream, const double* m, const double* W, const double* nu_weights, double* tau,
                    double* Mup, double* Mdn, double* h_F, double* d_F)
{
    CS_REQUIRE(s && P && wlob && mu && Tlev && m && W, CS_ERR_ARG, "null argument");
    cs_ctx* ctx = s->ctx;
    CS_TRY(check_profile_args(s, np, P, nlob));
    CS_REQUIRE(nstream >= 1 && nstream <= CS_MAX_STREAMS, CS_ERR_ARG, "nstream must be in [1,%d]", CS_MAX_STREAMS);
    // checkazimuth (fluxes.jl:4-6)
    CS_REQUIRE(theta_s >= 0 && theta_s < CS_PI / 2, CS_ERR_ARG, "azimuth angle theta must be in [0,pi/2)");
    CS_REQUIRE(g > 0, CS_ERR_ARG, "gravity must be positive");
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t nnu = s->nnu;
    const int L = (int)np - 1;

    // pack the small per-level arrays
    std::vector<double> small;
    small.reserve((size_t)(2 * np + nlob * L + nlob + 3 * nstream));
    small.insert(small.end(), P, P + np);
    small.insert(small.end(), mu, mu + (size_t)nlob * L);
    small.insert(small.end(), wlob, wlob + nlob);
    for (int64_t i = 0; i < np; i++) small.push_back(1.0 / (CS_KB * Tlev[i]));   // x = h c nu / (k T) as a product
    small.insert(small.end(), m, m + nstream);
    small.insert(small.end(), W, W + nstream);
    for (int k = 0; k < nstream; k++) small.push_back(1.0 / m[k]);

    // trapezoid weights: the workspace's own (computed at creation) unless the caller supplies global ones
    const double* wsrc = nu_weights;
    const int nblocks = (int)((nnu + RT_THREADS - 1) / RT_THREADS);
    size_t off_small = 0;
    size_t off_w = ((small.size() * sizeof(double) + 255) / 256) * 256;
    size_t off_fS = off_w + (((size_t)nnu * sizeof(double) + 255) / 256) * 256;
    size_t off_fa = off_fS + (((size_t)nnu * sizeof(double) + 255) / 256) * 256;
    size_t off_F = off_fa + (((size_t)nnu * sizeof(double) + 255) / 256) * 256;
    size_t total = off_F + sizeof(double) * 2 * (size_t)np;
    CS_TRY(ctx->s_misc.reserve(total));
    char* base = ctx->s_misc.as<char>();
    CS_TRY(cs_stage_h2d(ctx, base + off_small, small.data(), small.size() * sizeof(double)));
    if (wsrc) CS_CUDA(cudaMemcpyAsync(base + off_w, wsrc, sizeof(double) * (size_t)nnu, cudaMemcpyHostToDevice, st));
    if (fS) CS_CUDA(cudaMemcpyAsync(base + off_fS, fS, sizeof(double) * (size_t)nnu, cudaMemcpyHostToDevice, st));
    if (fa) CS_CUDA(cudaMemcpyAsync(base + off_fa, fa, sizeof(double) * (size_t)nnu, cudaMemcpyHostToDevice, st));
    CS_TRY(ctx->s_tau.reserve(sizeof(double) * (size_t)L * nnu));
    CS_TRY(ctx->s_planck.reserve(sizeof(double) * (size_t)np * nnu));
    // per-CTA tables: the small arrays, plus the per-warp accumulators while they fit; beyond that (Radau-equivalent
    // refinements with thousands of levels) the warps write their partial sums straight to the second stage's input
    const size_t smem_small = sizeof(double) * small.size();
    const bool red_global = smem_small + sizeof(double) * (size_t)2 * np * RT_WARPS > 96 * 1024;
    const int nparts = red_global ? nblocks * RT_WARPS : nblocks;
    CS_TRY(ctx->s_part.reserve(sizeof(double) * (size_t)nparts * 2 * np));
    if (tau) CS_TRY(ctx->s_out0.reserve(sizeof(double) * (size_t)L * nnu));
    if (Mup) CS_TRY(ctx->s_out1.reserve(sizeof(double) * (size_t)np * nnu));
    if (Mdn) CS_TRY(ctx->s_out2.reserve(sizeof(double) * (size_t)np * nnu));

    RtArgs a;
    a.sig = s->sig; a.nu = s->nu; a.w = wsrc ? (const double*)(base + off_w) : s->w;
    a.fS = fS ? (const double*)(base + off_fS) : nullptr;
    a.fa = fa ? (const double*)(base + off_fa) : nullptr;
    a.small = (const double*)(base + off_small);
    a.nnu = nnu; a.np = (int)np; a.nlob = nlob; a.ns = nstream;
    a.Cg = 1e-4 * CS_NA / g;          // fluxes.jl:259
    a.cos_s = cos(theta_s);
    a.tau_floor = ctx->tau_floor;
    a.tau_s = ctx->s_tau.as<double>(); a.B_s = ctx->s_planck.as<double>();
    a.tau_out = tau ? ctx->s_out0.as<double>() : nullptr;
    a.Mup_out = Mup ? ctx->s_out1.as<double>() : nullptr;
    a.Mdn_out = Mdn ? ctx->s_out2.as<double>() : nullptr;
    a.part = ctx->s_part.as<double>();
    a.red_global = red_global ? 1 : 0;
    size_t smem = red_global ? smem_small : smem_small + sizeof(double) * (size_t)2 * np * RT_WARPS;
    CS_REQUIRE(smem <= 200 * 1024, CS_ERR_ARG, "too many pressure levels for one flux call (%lld): per-CTA tables need %zu bytes", (long long)np, smem);

    const int sp_rt = cs_span_begin(ctx, CS_T_RT, true);
    switch (nstream) {
    case 1: CS_TRY(launch_rt<1>(ctx, a, nblocks, smem)); break;
    case 2: CS_TRY(launch_rt<2>(ctx, a, nblocks, smem)); break;
    case 3: CS_TRY(launch_rt<3>(ctx, a, nblocks, smem)); break;
    case 4: CS_TRY(launch_rt<4>(ctx, a, nblocks, smem)); break;
    case 5: CS_TRY(launch_rt<5>(ctx, a, nblocks, smem)); break;
    case 6: CS_TRY(launch_rt<6>(ctx, a, nblocks, smem)); break;
    case 7: CS_TRY(launch_rt<7>(ctx, a, nblocks, smem)); break;
    case 8: CS_TRY(launch_rt<8>(ctx, a, nblocks, smem)); break;
    default: CS_TRY(launch_rt<0>(ctx, a, nblocks, smem)); break;
    }
    cs_span_end(ctx, sp_rt);
    const int sp_red = cs_span_begin(ctx, CS_T_REDUCE, true);
    double* dF = d_F ? d_F : (double*)(base + off_F);
    flux_reduce_kernel<<<(2 * (int)np + 3) / 4, 128, 0, st>>>(a.part, nparts, 2 * (int)np, dF);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    cs_span_end(ctx, sp_red);

    if (!h_F && !tau && !Mup && !Mdn) {
        // device-side result only (cs_fluxes_device): nothing to wait for -- the caller's collective or read-back is
        // stream-ordered behind the reduction
        cs_spans_collect(ctx, false);
        return CS_OK;
    }
    if (h_F) CS_CUDA(cudaMemcpyAsync(h_F, dF, sizeof(double) * 2 * (size_t)np, cudaMemcpyDeviceToHost, st));
    if (tau) CS_CUDA(cudaMemcpyAsync(tau, a.tau_out, sizeof(double) * (size_t)L * nnu, cudaMemcpyDeviceToHost, st));
    if (Mup) CS_CUDA(cudaMemcpyAsync(Mup, a.Mup_out, sizeof(double) * (size_t)np * nnu, cudaMemcpyDeviceToHost, st));
    if (Mdn) CS_CUDA(cudaMemcpyAsync(Mdn, a.Mdn_out, sizeof(double) * (size_t)np * nnu, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    cs_spans_collect(ctx, false);
    return CS_OK;
}

}  // namespace

extern "C" int32_t cs_fluxes(cs_sigma* s, int64_t np, const double* P, int32_t nlob, const double* wlob,
                             const double* mu, const double* Tlev, double g, const double* fS, const double* fa,
                             double theta_s, int32_t nstream, const double* m, const double* W,
                             const double* nu_weights, double* tau, double* Mup, double* Mdn, double* Fup,
                             double* Fdn, double* Fnet)
{
    CS_REQUIRE(Fup && Fdn && Fnet, CS_ERR_ARG, "Fup/Fdn/Fnet outputs are required");
    std::vector<double> F(2 * (size_t)(np > 0 ? np : 1));
    CS_TRY(fluxes_impl(s, np, P, nlob, wlob, mu, Tlev, g, fS, fa, theta_s, nstream, m, W, nu_weights, tau, Mup, Mdn,
                       F.data(), nullptr));
    for (int64_t i = 0; i < np; i++) {
        Fup[i] = F[(size_t)i];
        Fdn[i] = F[(size_t)(np + i)];
        Fnet[i] = Fup[i] - Fdn[i];     // fluxes.jl:380
    }
    return CS_OK;
}

extern "C" int32_t cs_fluxes_device(cs_sigma* s, int64_t np, const double* P, int32_t nlob, const double* wlob,
                                    const double* mu, const double* Tlev, double g, const double* fS,
                                    const double* fa, double theta_s, int32_t nstream, const double* m,
                                    const double* W, const double* nu_weights, double* d_F)
{
    CS_REQUIRE(d_F, CS_ERR_ARG, "null device output");
    return fluxes_impl(s, np, P, nlob, wlob, mu, Tlev, g, fS, fa, theta_s, nstream, m, W, nu_weights, nullptr, nullptr,
                       nullptr, nullptr, d_F);
}

// fluxes for nbatch temperature profiles over the same Sigma workspace: F[b][0..np) = F+, F[b][np..2np) = F-
extern "C" int32_t cs_fluxes_batch(cs_sigma* s, int64_t np, const double* P, int32_t nlob, const double* wlob,
                                   const double* mu, int64_t nbatch, const double* Tlev, double g, const double* fS,
                                   const double* fa, double theta_s, int32_t nstream, const double* m, const double* W,
                                   const double* nu_weights, double* F)
{
    CS_REQUIRE(s && P && wlob && mu && Tlev && m && W && F, CS_ERR_ARG, "null argument");
    CS_REQUIRE(nbatch >= 1, CS_ERR_ARG, "empty batch");
    cs_ctx* ctx = s->ctx;
    CS_TRY(check_profile_args(s, np, P, nlob));
    CS_REQUIRE(nstream >= 1 && nstream <= CS_MAX_STREAMS, CS_ERR_ARG, "nstream must be in [1,%d]", CS_MAX_STREAMS);
    CS_REQUIRE(theta_s >= 0 && theta_s < CS_PI / 2, CS_ERR_ARG, "azimuth angle theta must be in [0,pi/2)");
    CS_REQUIRE(g > 0, CS_ERR_ARG, "gravity must be positive");
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t nnu = s->nnu;
    const int L = (int)np - 1;
    std::vector<double> small;
    small.insert(small.end(), P, P + np);
    small.insert(small.end(), mu, mu + (size_t)nlob * L);
    small.insert(small.end(), wlob, wlob + nlob);
    small.insert(small.end(), (size_t)np, 0.0);                 // slot of the single-profile 1/(kT) table (unused here)
    small.insert(small.end(), m, m + nstream);
    small.insert(small.end(), W, W + nstream);
    for (int k = 0; k < nstream; k++) small.push_back(1.0 / m[k]);
    std::vector<double> rkT((size_t)nbatch * np);
    for (size_t i = 0; i < rkT.size(); i++) rkT[i] = 1.0 / (CS_KB * Tlev[i]);

    const int nblocks = (int)((nnu + RT_THREADS - 1) / RT_THREADS);
    auto al = [](size_t b) { return ((b + 255) / 256) * 256; };
    const size_t off_w = al(small.size() * sizeof(double));
    const size_t off_fS = off_w + al((size_t)nnu * sizeof(double));
    const size_t off_fa = off_fS + al((size_t)nnu * sizeof(double));
    const size_t off_kT = off_fa + al((size_t)nnu * sizeof(double));
    const size_t off_F = off_kT + al(rkT.size() * sizeof(double));
    CS_TRY(ctx->s_misc.reserve(off_F + sizeof(double) * 2 * (size_t)np * RTB_NB));
    char* base = ctx->s_misc.as<char>();
    CS_TRY(cs_stage_h2d(ctx, base, small.data(), small.size() * sizeof(double)));
    if (nu_weights) CS_CUDA(cudaMemcpyAsync(base + off_w, nu_weights, sizeof(double) * (size_t)nnu, cudaMemcpyHostToDevice, st));
    if (fS) CS_CUDA(cudaMemcpyAsync(base + off_fS, fS, sizeof(double) * (size_t)nnu, cudaMemcpyHostToDevice, st));
    if (fa) CS_CUDA(cudaMemcpyAsync(base + off_fa, fa, sizeof(double) * (size_t)nnu, cudaMemcpyHostToDevice, st));
    CS_TRY(cs_stage_h2d(ctx, base + off_kT, rkT.data(), rkT.size() * sizeof(double)));
    CS_TRY(ctx->s_tau.reserve(sizeof(double) * (size_t)L * nnu));
    CS_TRY(ctx->s_planck.reserve(sizeof(double) * (size_t)RTB_NB * np * nnu));
    CS_TRY(ctx->s_part.reserve(sizeof(double) * (size_t)nblocks * RTB_NB * 2 * np));

    RtBatchArgs ba;
    RtArgs& a = ba.a;
    a.sig = s->sig; a.nu = s->nu; a.w = nu_weights ? (const double*)(base + off_w) : s->w;
    a.fS = fS ? (const double*)(base + off_fS) : nullptr;
    a.fa = fa ? (const double*)(base + off_fa) : nullptr;
    a.small = (const double*)base;
    a.nnu = nnu; a.np = (int)np; a.nlob = nlob; a.ns = nstream;
    a.Cg = 1e-4 * CS_NA / g;
    a.cos_s = cos(theta_s);
    a.tau_floor = ctx->tau_floor;
    a.tau_s = ctx->s_tau.as<double>(); a.B_s = nullptr;
    a.tau_out = nullptr; a.Mup_out = nullptr; a.Mdn_out = nullptr; a.part = nullptr; a.red_global = 0;
    ba.B_b = ctx->s_planck.as<double>();
    ba.part_b = ctx->s_part.as<double>();
    const size_t smem = sizeof(double) * (small.size() + (size_t)RTB_NB * np + (size_t)RTB_NB * 2 * np * RT_WARPS);
    CS_REQUIRE(smem <= 200 * 1024, CS_ERR_ARG, "too many pressure levels for a batched flux call (%lld): per-CTA tables need %zu bytes",
               (long long)np, smem);
    CS_CUDA(cudaFuncSetAttribute(rt_batch_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CS_CUDA(cudaFuncSetAttribute(rt_batch_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sp_rt = cs_span_begin(ctx, CS_T_RT, true);
    double* dF = (double*)(base + off_F);
    for (int64_t b0 = 0; b0 < nbatch; b0 += RTB_NB) {
        ba.nb = (int)std::min<int64_t>(RTB_NB, nbatch - b0);
        ba.rkT = (const double*)(base + off_kT) + (size_t)b0 * np;
        if (nstream == 5) rt_batch_kernel<5><<<nblocks, RT_THREADS, smem, st>>>(ba);
        else rt_batch_kernel<0><<<nblocks, RT_THREADS, smem, st>>>(ba);
        CS_CUDA(cudaGetLastError());
        const int n2 = ba.nb * 2 * (int)np;
        flux_reduce_kernel<<<(n2 + 3) / 4, 128, 0, st>>>(ba.part_b, nblocks, n2, dF);
        CS_CUDA(cudaGetLastError());
        cs_count_launch(ctx, 2);
        CS_CUDA(cudaMemcpyAsync(F + (size_t)b0 * 2 * np, dF, sizeof(double) * (size_t)n2, cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaStreamSynchronize(st));       // dF and the scratch are reused by the next chunk
    }
    cs_span_end(ctx, sp_rt);
    cs_spans_collect(ctx, false);
    return CS_OK;
}

extern "C" int32_t cs_opticaldepth(cs_sigma* s, int64_t np, const double* P, int32_t nlob, const double* wlob,
                                   const double* mu, double g, double theta, double* tau_total)
{
    CS_REQUIRE(s && P && wlob && mu && tau_total, CS_ERR_ARG, "null argument");
    cs_ctx* ctx = s->ctx;
    CS_TRY(check_profile_args(s, np, P, nlob));
    CS_REQUIRE(theta >= 0 && theta < CS_PI / 2, CS_ERR_ARG, "azimuth angle theta must be in [0,pi/2)");
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int L = (int)np - 1;
    std::vector<double> small;
    small.insert(small.end(), P, P + np);
    small.insert(small.end(), mu, mu + (size_t)nlob * L);
    small.insert(small.end(), wlob, wlob + nlob);
    size_t off_out = ((small.size() * sizeof(double) + 255) / 256) * 256;
    CS_TRY(ctx->s_misc.reserve(off_out + sizeof(double) * (size_t)s->nnu));
    char* base = ctx->s_misc.as<char>();
    CS_TRY(cs_stage_h2d(ctx, base, small.data(), small.size() * sizeof(double)));
    int nblocks = (int)((s->nnu + RT_THREADS - 1) / RT_THREADS);
    const size_t dsm = small.size() * sizeof(double);
    CS_REQUIRE(dsm <= 200 * 1024, CS_ERR_ARG, "too many pressure levels for one optical-depth call (%lld): per-CTA tables need %zu bytes",
               (long long)np, dsm);
    if (dsm > 48 * 1024) CS_CUDA(cudaFuncSetAttribute(depth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
    depth_kernel<<<nblocks, RT_THREADS, dsm, st>>>(
        s->sig, (const double*)base, s->nnu, (int)np, nlob, 1e-4 * CS_NA / g, 1 / cos(theta), (double*)(base + off_out));
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    CS_CUDA(cudaMemcpyAsync(tau_total, base + off_out, sizeof(double) * (size_t)s->nnu, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}

// ================================================================================================
// Device-resident radiative-convective loop.
//
// Replaces the per-step body of the reference's RCM (src/radiative_convective.jl): heating! :109-144 = radiate! on the
// radiative levels Pr with the AcceleratedAbsorber, Fnet interpolated to the cell edges (:123-124), cell and surface
// heating rates (:129-140); step! :147-151 = T += dt*H.  The reference never calls update!(A, T) inside heating!, so
// Sigma -- an AcceleratedAbsorber ignores T (absorbers.jl:203) -- is the same at every step, and with it every layer
// optical depth, every stream transmittance exp(-tau*m_k) and the whole stellar beam.  cs_rcm_create computes those once
// (rcm_static_kernel); a step then only evaluates Planck at the new level temperatures and runs the stream recurrences
// on stored transmittances (rcm_rt_kernel: HBM-bound, (NS+1) doubles per (layer, nu) and sweep, instead of
// (2 NS + 1) exponentials per (layer, nu)), reduces spectrally (flux_reduce_kernel) and updates the column
// (rcm_update_kernel).  The three kernels of a step are captured in ONE CUDA graph; the host sees the temperatures only
// when it asks (cs_rcm_state).
namespace {

struct RcmStaticArgs {
    const double* sig;     // [nnode][nnu]
    const double* w;       // [nnu]
    const double* fS;      // [nnu] or null
    const double* small;   // packed: P[np], mu[nlob*(np-1)], wl[nlob], m[ns]
    int64_t nnu;
    int np, nlob, ns;
    double Cg, cos_s, tau_floor;
    double* tau;           // [np-1][nnu]
    double* tr;            // [np-1][ns][nnu]
    double* beam_surf;     // [nnu] stellar beam at the surface (or null)
    double* part;          // [nblocks][np] partial sums of w*beam per level (or null)
};

__global__ void __launch_bounds__(RT_THREADS) rcm_static_kernel(RcmStaticArgs a)
{
    extern __shared__ double sm[];
    const int np = a.np, L = np - 1, nlob = a.nlob, ns = a.ns;
    const int nsmall = np + nlob * L + nlob + ns;
    double* sP = sm;
    double* smu = sP + np;
    double* swl = smu + nlob * L;
    double* sm_m = swl + nlob;
    double* red = sm + nsmall;      // [np][RT_WARPS]
    for (int t = threadIdx.x; t < nsmall; t += RT_THREADS) sm[t] = a.small[t];
    for (int t = threadIdx.x; t < np * RT_WARPS; t += RT_THREADS) red[t] = 0.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t jraw = (int64_t)blockIdx.x * RT_THREADS + threadIdx.x;
    const bool live = jraw < a.nnu;
    const int64_t j = live ? jraw : a.nnu - 1;
    const int64_t nnu = a.nnu;
    const double wj = live ? a.w[j] : 0.0;
    const double Cg = a.Cg, rc = 1.0 / a.cos_s;
    // same operations, in the same order, as the downward loop of rt_kernel (d-depth!, discretized.jl:136-177)
    double beta1 = Cg * (a.sig[j] / smu[0]);
    double beam = a.cos_s * (a.fS ? a.fS[j] : 0.0);        // M-[1] = c*fS(nu)   (discretized.jl:299)
    if (a.part) {
        double r = warp_sum(wj * beam);
        if (lane == 0) red[0 * RT_WARPS + warp] += r;
    }
    for (int i = 0; i < L; i++) {
        double dP = sP[i + 1] - sP[i];
        double ti = 0.0;
        ti += (dP * swl[0]) * beta1;
        for (int n = 1; n < nlob - 1; n++) {
            double bn = Cg * (a.sig[(size_t)(n + (nlob - 1) * i) * nnu + j] / smu[n + nlob * i]);
            ti += (dP * swl[n]) * bn;
        }
        double bn = Cg * (a.sig[(size_t)((nlob - 1) * (i + 1)) * nnu + j] / smu[(nlob - 1) + nlob * i]);
        ti += (dP * swl[nlob - 1]) * bn;
        beta1 = bn;
        double tau = fmax(ti, a.tau_floor);
        if (live) a.tau[(size_t)i * nnu + j] = tau;
        for (int k = 0; k < ns; k++) {
            double tk = tau * sm_m[k];
            if (live) a.tr[((size_t)i * ns + k) * nnu + j] = exp(-tk);
        }
        beam *= exp(-tau * rc);                              // discretized.jl:302
        if (a.part) {
            double r = warp_sum(wj * beam);
            if (lane == 0) red[(i + 1) * RT_WARPS + warp] += r;
        }
    }
    if (a.beam_surf && live) a.beam_surf[j] = beam;
    if (a.part) {
        __syncthreads();
        for (int t = threadIdx.x; t < np; t += RT_THREADS) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < RT_WARPS; q++) s += red[t * RT_WARPS + q];
            a.part[(size_t)blockIdx.x * np + t] = s;
        }
    }
}

struct RcmRtArgs {
    const double* nu;          // [nnu]
    const double* w;           // [nnu]
    const double* fa;          // [nnu] or null
    const double* beam_surf;   // [nnu] or null
    const double* tau;         // [L][nnu]
    const double* tr;          // [L][ns][nnu]
    const double* rkT;         // [np] 1/(k T) at the radiative levels (device, rewritten every step)
    const double* stream;      // W[ns], 1/m[ns]
    int64_t nnu;
    int np, ns;
    double* B_s;               // scratch [np][nnu]
    double* part;              // [nblocks][2][np]
};

template <int NS>
__global__ void __launch_bounds__(RT_THREADS) rcm_rt_kernel(RcmRtArgs a)
{
    extern __shared__ double sm[];
    const int np = a.np, L = np - 1;
    const int ns = (NS > 0) ? NS : a.ns;
    double* skT = sm;                    // [np]
    double* sW = skT + np;               // [ns]
    double* srm = sW + ns;               // [ns]
    double* red = srm + ns;              // [2][np][RT_WARPS]
    for (int t = threadIdx.x; t < np; t += RT_THREADS) skT[t] = a.rkT[t];
    for (int t = threadIdx.x; t < 2 * ns; t += RT_THREADS) sW[t] = a.stream[t];
    for (int t = threadIdx.x; t < 2 * np * RT_WARPS; t += RT_THREADS) red[t] = 0.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t jraw = (int64_t)blockIdx.x * RT_THREADS + threadIdx.x;
    const bool live = jraw < a.nnu;
    const int64_t j = live ? jraw : a.nnu - 1;
    const int64_t nnu = a.nnu;
    const double wj = live ? a.w[j] : 0.0;
    const double nuj = a.nu[j];
    const double num = 100.0 * nuj;
    const double hcn = CS_H * CS_C * num;
    const double pref = 2 * CS_H * (CS_C * CS_C) * (num * num * num);
    constexpr int NSMAX = (NS > 0) ? NS : CS_MAX_STREAMS;
    double I[NSMAX], tk[NSMAX], tn[NSMAX];
#pragma unroll
    for (int k = 0; k < NSMAX; k++) I[k] = 0.0;
    // ---- downward (diffuse part; the stellar beam is static and added by rcm_update_kernel)
    double Bprev = 100.0 * pref / (exp(hcn * skT[0]) - 1.0);
    a.B_s[j] = Bprev;
    double tau_n = a.tau[j];
#pragma unroll
    for (int k = 0; k < NSMAX; k++) tn[k] = (k < ns) ? a.tr[(size_t)k * nnu + j] : 0.0;
    // operands of the layer after next as well: with few wavenumbers per device (a nu slice of an 8-way split) there are only
    // ~8 warps per SM, and one layer of loads in flight per thread does not cover the HBM latency
    double tau_n2 = 0.0, tn2[NSMAX];
#pragma unroll
    for (int k = 0; k < NSMAX; k++) tn2[k] = 0.0;
    if (L > 1) {
        tau_n2 = a.tau[nnu + j];
#pragma unroll
        for (int k = 0; k < NSMAX; k++)
            if (k < ns) tn2[k] = a.tr[((size_t)ns + k) * nnu + j];
    }
    double Msum = 0.0;
    for (int i = 0; i < L; i++) {
        const double tau = tau_n;
#pragma unroll
        for (int k = 0; k < NSMAX; k++) tk[k] = tn[k];
        tau_n = tau_n2;
#pragma unroll
        for (int k = 0; k < NSMAX; k++) tn[k] = tn2[k];
        if (i + 2 < L) {      // operands two layers ahead
            tau_n2 = a.tau[(size_t)(i + 2) * nnu + j];
#pragma unroll
            for (int k = 0; k < NSMAX; k++)
                if (k < ns) tn2[k] = a.tr[((size_t)(i + 2) * ns + k) * nnu + j];
        }
        const double Bnext = 100.0 * pref / (exp(hcn * skT[i + 1]) - 1.0);
        a.B_s[(size_t)(i + 1) * nnu + j] = Bnext;
        const double rtau = cs_rcp(tau);
        Msum = 0.0;
#pragma unroll
        for (int k = 0; k < NSMAX; k++) {
            if (k < ns) {
                double Be = layerplanck(Bprev, Bnext, rtau * srm[k], tk[k]);
                I[k] = I[k] * tk[k] + Be;
                Msum += sW[k] * I[k];
            }
        }
        double r = warp_sum(wj * Msum);
        if (lane == 0) red[(np + i + 1) * RT_WARPS + warp] += r;
        Bprev = Bnext;
    }
    // ---- surface (discretized.jl:309-310): M-[end] = diffuse + beam
    const double Mdn_s = Msum + (a.beam_surf ? a.beam_surf[j] : 0.0);
    const double Is = Mdn_s * (a.fa ? a.fa[j] : 0.0) / CS_PI + Bprev;
    {
        double r = warp_sum(wj * (Is * CS_PI));
        if (lane == 0) red[L * RT_WARPS + warp] += r;
    }
    // ---- upward
#pragma unroll
    for (int k = 0; k < NSMAX; k++) I[k] = Is;
    double B1 = Bprev;
    tau_n = a.tau[(size_t)(L - 1) * nnu + j];
    double B_n = a.B_s[(size_t)(L - 1) * nnu + j];
#pragma unroll
    for (int k = 0; k < NSMAX; k++) tn[k] = (k < ns) ? a.tr[((size_t)(L - 1) * ns + k) * nnu + j] : 0.0;
    double B_n2 = 0.0;
    if (L > 1) {
        tau_n2 = a.tau[(size_t)(L - 2) * nnu + j];
        B_n2 = a.B_s[(size_t)(L - 2) * nnu + j];
#pragma unroll
        for (int k = 0; k < NSMAX; k++)
            if (k < ns) tn2[k] = a.tr[((size_t)(L - 2) * ns + k) * nnu + j];
    }
    for (int i = L - 1; i >= 0; i--) {
        const double tau = tau_n, B2 = B_n;
#pragma unroll
        for (int k = 0; k < NSMAX; k++) tk[k] = tn[k];
        tau_n = tau_n2;
        B_n = B_n2;
#pragma unroll
        for (int k = 0; k < NSMAX; k++) tn[k] = tn2[k];
        if (i > 1) {
            tau_n2 = a.tau[(size_t)(i - 2) * nnu + j];
            B_n2 = a.B_s[(size_t)(i - 2) * nnu + j];
#pragma unroll
            for (int k = 0; k < NSMAX; k++)
                if (k < ns) tn2[k] = a.tr[((size_t)(i - 2) * ns + k) * nnu + j];
        }
        const double rtau = cs_rcp(tau);
        double Ms = 0.0;
#pragma unroll
        for (int k = 0; k < NSMAX; k++) {
            if (k < ns) {
                double Be = layerplanck(B1, B2, rtau * srm[k], tk[k]);
                I[k] = I[k] * tk[k] + Be;
                Ms += sW[k] * I[k];
            }
        }
        double r = warp_sum(wj * Ms);
        if (lane == 0) red[i * RT_WARPS + warp] += r;
        B1 = B2;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * np; t += RT_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < RT_WARPS; q++) s += red[t * RT_WARPS + q];
        a.part[(size_t)blockIdx.x * 2 * np + t] = s;
    }
}

// column update: everything of heating!/step! that is O(np) (radiative_convective.jl:123-150), in the reference's
// operation order and WITHOUT fused multiply-adds (Julia does not contract a*b+c), then the level temperatures of the
// next step (AtmosphericProfile(P, T) at Pr: linear in ln P with end-cell extrapolation, atmospherics.jl:16-26).
struct RcmColArgs {
    int np, nrad;
    const double* F;         // [2*nrad] spectrally integrated diffuse F+ then F- (all ranks summed)
    const double* beamF;     // [nrad] static stellar-beam part of F- (all ranks summed) or null
    const double* lnPr;      // [nrad]
    const double* lnPe;      // [np]
    const int* ecell;        // [np] cell of ln Pe in lnPr
    const double* dPe;       // [np-1] Pe[i+1] - Pe[i]
    const double* gcp;       // [np-1] g / cp_i
    double cs_surf;
    const double* lnP;       // [np] ln of cell-centre pressures (+ surface)
    const int* rcell;        // [nrad] cell of ln Pr in lnP
    const double* dt;        // device scalar
    double* T;               // [np]
    double* H;               // [np]
    double* R;               // [np]
    double* Fout;            // [3*nrad] F+, F-, Fnet of this step (beam included)
    double* rkT;             // [nrad] 1/(k T(Pr)) for the next step
    double* Tlev;            // [nrad]
};

__device__ __forceinline__ double rcm_lin(const double* x, const double* y, int i, double q)
{
    // LinearInterpolator: (q - x[i]) * (y[i+1] - y[i]) / (x[i+1] - x[i]) + y[i]
    return __dadd_rn(__ddiv_rn(__dmul_rn(q - x[i], y[i + 1] - y[i]), x[i + 1] - x[i]), y[i]);
}

__device__ __forceinline__ void rcm_update_body(const RcmColArgs& a, double* sm)
{
    double* sFnet = sm;                  // [nrad]
    double* sR = sFnet + a.nrad;         // [np]
    double* sT = sR + a.np;              // [np]
    const int np = a.np, nrad = a.nrad;
    if (a.F) {
        for (int r = threadIdx.x; r < nrad; r += blockDim.x) {
            const double up = a.F[r], dn = a.F[nrad + r] + (a.beamF ? a.beamF[r] : 0.0);
            const double net = up - dn;                                  // fluxes.jl:380
            sFnet[r] = net;
            a.Fout[r] = up; a.Fout[nrad + r] = dn; a.Fout[2 * nrad + r] = net;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < np; e += blockDim.x) {
            const double Rn = -rcm_lin(a.lnPr, sFnet, a.ecell[e], a.lnPe[e]);   // R = -fFnet(Pe)  (:124)
            sR[e] = Rn;
            a.R[e] = Rn;
        }
        __syncthreads();
        const double dt = *a.dt;
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            double H;
            if (i < np - 1) H = __ddiv_rn(__dmul_rn(a.gcp[i], sR[i] - sR[i + 1]), a.dPe[i]);   // (g/cp)*dR/dP  (:137)
            else H = __ddiv_rn(sR[np - 1], a.cs_surf);                                          // (:140)
            a.H[i] = H;
            const double Tn = __dadd_rn(a.T[i], __dmul_rn(dt, H));                              // step! (:149)
            a.T[i] = Tn;
            sT[i] = Tn;
        }
    } else {
        for (int i = threadIdx.x; i < np; i += blockDim.x) sT[i] = a.T[i];
    }
    __syncthreads();
    for (int r = threadIdx.x; r < nrad; r += blockDim.x) {
        const double Tl = rcm_lin(a.lnP, sT, a.rcell[r], a.lnPr[r]);
        a.Tlev[r] = Tl;
        a.rkT[r] = 1.0 / (CS_KB * Tl);
    }
}

__global__ void __launch_bounds__(256) rcm_update_kernel(RcmColArgs a)
{
    extern __shared__ double sm[];
    rcm_update_body(a, sm);
}

// ---- nu-sharded step with the collective fused into the step's last kernel.  Every rank owns a mailbox in its own device
// memory, mapped into the other ranks' address spaces (CUDA IPC between processes, peer access inside one):
//   flags  uint64 [2][nranks]          flag[slot][r] = step + 1 once rank r's partial sums of that step have landed
//   data   double [2][nranks][n2]      n2 = 2 nrad partial sums (F+ then F-) of rank r
// The tail kernel (one CTA) does the fixed-order spectral reduction of this rank's per-CTA partials (as flux_reduce_kernel),
// stores the n2 sums straight into EVERY rank's mailbox over NVLink (plain peer stores, then a system fence, then the flag),
// spins on its own mailbox until all nranks flags of this step are there, adds the nranks vectors in RANK ORDER -- the same
// bits on every rank, which the replicated column update needs -- and runs the column update.  Slots alternate with the step
// parity: a rank can only be one step ahead of the slowest one (it needs everybody's flag to finish a step), and it posts step
// s+1 after it has consumed step s, so two slots are enough.  A bounded spin (about two seconds of SM clocks) raises an error
// word instead of hanging the device when a rank never arrives.
constexpr int RCM_MAX_RANKS = 16;
struct RcmPeerArgs {
    const double* part;                  // [nblocks][n2] per-CTA partial sums of rcm_rt_kernel
    int nblocks, n2, rank, nranks;
    unsigned long long* step;            // device word: steps completed so far (the same on every rank)
    int* err;                            // device word: set when a flag did not arrive in time
    size_t data_off;                     // byte offset of the data block inside a mailbox
    char* mail[RCM_MAX_RANKS];           // the nranks mailboxes as visible from this device (mail[rank] = its own)
};

__global__ void __launch_bounds__(256) rcm_tail_kernel(RcmColArgs a, RcmPeerArgs p)
{
    extern __shared__ double sm[];
    double* sF = sm + a.nrad + 2 * a.np;          // [n2], behind the column update's scratch
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n2 = p.n2;
    if (p.part) {
        for (int t = warp; t < n2; t += 8) {
            double v = 0.0;
            for (int b = lane; b < p.nblocks; b += 32) v += p.part[(size_t)b * n2 + t];
            v = warp_sum(v);
            if (lane == 0) sF[t] = v;
        }
    } else {
        for (int t = tid; t < n2; t += 256) sF[t] = a.F[t];      // already reduced by flux_reduce_kernel (one warp per output)
    }
    __syncthreads();
    if (p.nranks > 1) {
        const unsigned long long step = *p.step;
        const int slot = (int)(step & 1ULL);
        for (int q = 0; q < p.nranks; q++) {
            double* dst = reinterpret_cast<double*>(p.mail[q] + p.data_off) + ((size_t)slot * p.nranks + p.rank) * n2;
            for (int t = tid; t < n2; t += 256) dst[t] = sF[t];
        }
        __threadfence_system();
        __syncthreads();
        if (tid < p.nranks) {
            volatile unsigned long long* f = reinterpret_cast<unsigned long long*>(p.mail[tid]) + (size_t)slot * p.nranks + p.rank;
            *f = step + 1ULL;
        }
        if (tid < p.nranks) {
            volatile unsigned long long* f = reinterpret_cast<unsigned long long*>(p.mail[p.rank]) + (size_t)slot * p.nranks + tid;
            const long long t0 = clock64();
            while (*f != step + 1ULL) {
                if (clock64() - t0 > 4000000000LL) { atomicExch(p.err, 1); break; }
            }
        }
        __syncthreads();
        __threadfence_system();
        const double* src = reinterpret_cast<const double*>(p.mail[p.rank] + p.data_off) + (size_t)slot * p.nranks * n2;
        for (int t = tid; t < n2; t += 256) {
            double v = 0.0;
            for (int r = 0; r < p.nranks; r++) v += __ldcv(src + (size_t)r * n2 + t);
            sF[t] = v;
        }
        __syncthreads();
        if (tid == 0) *p.step = step + 1ULL;
    }
    a.F = sF;
    rcm_update_body(a, sm);
}

int findcell_host(const std::vector<double>& x, double q)
{
    const int n = (int)x.size();
    if (q <= x[0]) return 0;
    if (q >= x[(size_t)n - 1]) return n - 2;
    return (int)(std::upper_bound(x.begin(), x.end(), q) - x.begin()) - 1;
}

}  // namespace

struct cs_rcm {
    cs_ctx* ctx = nullptr;
    int64_t nnu = 0;
    int np = 0, nrad = 0, ns = 0;
    int nblocks = 0;
    bool has_beam = false, has_albedo = false;
    // owned device memory (one allocation each, stream-ordered)
    double *tau = nullptr, *tr = nullptr, *B_s = nullptr, *part = nullptr, *beam_surf = nullptr, *fa = nullptr;
    double *nu = nullptr, *w = nullptr;
    char* col = nullptr;      // packed column block (see offsets below)
    size_t col_bytes = 0;
    // pointers into col
    double *F = nullptr, *beamF = nullptr, *lnPr = nullptr, *lnPe = nullptr, *dPe = nullptr, *gcp = nullptr, *lnP = nullptr;
    double *dt = nullptr, *T = nullptr, *H = nullptr, *R = nullptr, *Fout = nullptr, *rkT = nullptr, *Tlev = nullptr, *stream = nullptr;
    int *ecell = nullptr, *rcell = nullptr;
    double cs_surf = 1.0;
    double dt_host = -1.0;    // value currently in *dt (graph replays read it from device memory)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool graph_tried = false;
    // peer exchange (cs_rcm_peer_*): own mailbox (cudaMalloc: exportable), the ranks' mailboxes as seen from this device
    char* mail_own = nullptr;
    size_t mail_bytes = 0, mail_data_off = 0;
    int peer_rank = 0, peer_nranks = 0;
    char* mail[16] = {nullptr};
    unsigned long long* peer_step = nullptr;     // [2 words]: step counter, error flag
};

namespace {

template <int NS> int32_t rcm_launch_rt(cs_rcm* r, cudaStream_t st, const RcmRtArgs& a, size_t smem)
{
    rcm_rt_kernel<NS><<<r->nblocks, RT_THREADS, smem, st>>>(a);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

size_t rcm_rt_smem(const cs_rcm* r) { return sizeof(double) * ((size_t)r->nrad + 2 * r->ns + (size_t)2 * r->nrad * RT_WARPS); }

int32_t rcm_enqueue_rt(cs_rcm* r);

// partial fluxes of this device: rt + spectral reduction into dF (2*nrad doubles, device)
int32_t rcm_enqueue_fluxes(cs_rcm* r, double* dF)
{
    CS_TRY(rcm_enqueue_rt(r));
    flux_reduce_kernel<<<(2 * r->nrad + 3) / 4, 128, 0, r->ctx->stream>>>(r->part, r->nblocks, 2 * r->nrad, dF);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(r->ctx, 1);
    return CS_OK;
}

// the flux kernel alone: per-CTA partial sums in r->part
int32_t rcm_enqueue_rt(cs_rcm* r)
{
    cs_ctx* ctx = r->ctx;
    cudaStream_t st = ctx->stream;
    RcmRtArgs a;
    a.nu = r->nu; a.w = r->w; a.fa = r->has_albedo ? r->fa : nullptr; a.beam_surf = r->has_beam ? r->beam_surf : nullptr;
    a.tau = r->tau; a.tr = r->tr; a.rkT = r->rkT; a.stream = r->stream;
    a.nnu = r->nnu; a.np = r->nrad; a.ns = r->ns; a.B_s = r->B_s; a.part = r->part;
    const size_t smem = rcm_rt_smem(r);
    switch (r->ns) {
    case 3: CS_TRY(rcm_launch_rt<3>(r, st, a, smem)); break;
    case 4: CS_TRY(rcm_launch_rt<4>(r, st, a, smem)); break;
    case 5: CS_TRY(rcm_launch_rt<5>(r, st, a, smem)); break;
    case 6: CS_TRY(rcm_launch_rt<6>(r, st, a, smem)); break;
    case 8: CS_TRY(rcm_launch_rt<8>(r, st, a, smem)); break;
    default: CS_TRY(rcm_launch_rt<0>(r, st, a, smem)); break;
    }
    cs_count_launch(ctx, 1);
    return CS_OK;
}

RcmColArgs rcm_col_args(const cs_rcm* r, const double* dF)
{
    RcmColArgs a;
    a.np = r->np; a.nrad = r->nrad; a.F = dF; a.beamF = r->has_beam ? r->beamF : nullptr;
    a.lnPr = r->lnPr; a.lnPe = r->lnPe; a.ecell = r->ecell; a.dPe = r->dPe; a.gcp = r->gcp; a.cs_surf = r->cs_surf;
    a.lnP = r->lnP; a.rcell = r->rcell; a.dt = r->dt; a.T = r->T; a.H = r->H; a.R = r->R; a.Fout = r->Fout;
    a.rkT = r->rkT; a.Tlev = r->Tlev;
    return a;
}

// flux kernels + the fused tail (peer exchange, column update): no collective call
int32_t rcm_enqueue_step_peer(cs_rcm* r)
{
    // the second-stage spectral reduction keeps its own launch: one warp per output over ~600 per-CTA partials is 2 us on 51
    // CTAs, but 50+ us when a single CTA walks the 2*nrad outputs (measured: the fully fused tail made the step slower)
    CS_TRY(rcm_enqueue_fluxes(r, r->F));
    RcmColArgs a = rcm_col_args(r, r->F);       // this rank's sums; the tail replaces them by the sums over all ranks
    RcmPeerArgs p;
    p.part = nullptr; p.nblocks = r->nblocks; p.n2 = 2 * r->nrad; p.rank = r->peer_rank; p.nranks = r->peer_nranks;
    p.step = r->peer_step; p.err = reinterpret_cast<int*>(r->peer_step + 1); p.data_off = r->mail_data_off;
    for (int q = 0; q < RCM_MAX_RANKS; q++) p.mail[q] = q < r->peer_nranks ? r->mail[q] : nullptr;
    const size_t smem = sizeof(double) * (3 * (size_t)r->nrad + 2 * (size_t)r->np);
    rcm_tail_kernel<<<1, 256, smem, r->ctx->stream>>>(a, p);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(r->ctx, 1);
    return CS_OK;
}

// column update from summed fluxes dF (null: only recompute the level temperatures from T)
int32_t rcm_enqueue_update(cs_rcm* r, const double* dF)
{
    cudaStream_t st = r->ctx->stream;
    RcmColArgs a = rcm_col_args(r, dF);
    const size_t smem = sizeof(double) * ((size_t)r->nrad + 2 * (size_t)r->np);
    rcm_update_kernel<<<1, 256, smem, st>>>(a);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(r->ctx, 1);
    return CS_OK;
}

int32_t rcm_set_dt(cs_rcm* r, double dt)
{
    if (dt == r->dt_host) return CS_OK;
    CS_TRY(cs_stage_h2d(r->ctx, r->dt, &dt, sizeof(double)));
    r->dt_host = dt;
    return CS_OK;
}

}  // namespace

extern "C" int32_t cs_rcm_free(cs_rcm* r)
{
    if (!r) return CS_OK;
    cudaSetDevice(r->ctx->device);
    cudaStream_t st = r->ctx->stream;
    if (r->exec) cudaGraphExecDestroy(r->exec);
    if (r->graph) cudaGraphDestroy(r->graph);
    cs_free(r->tau, st); cs_free(r->tr, st); cs_free(r->B_s, st); cs_free(r->part, st); cs_free(r->beam_surf, st);
    cs_free(r->fa, st); cs_free(r->nu, st); cs_free(r->w, st); cs_free(r->col, st);
    if (r->mail_own || r->peer_step) cudaStreamSynchronize(st);
    if (r->mail_own) cudaFree(r->mail_own);       // the peers must have closed their mappings (cs_ipc_close) before
    if (r->peer_step) cudaFree(r->peer_step);
    delete r;
    return CS_OK;
}

extern "C" int32_t cs_rcm_create(cs_sigma* s, int64_t np, const double* Pe, const double* P, const double* T0,
                                 const double* cp, double c_surf, int64_t nrad, const double* Pr, int32_t nlob,
                                 const double* wlob, const double* mu, double g, const double* fS, const double* fa,
                                 double theta_s, int32_t nstream, const double* m, const double* W,
                                 const double* nu_weights, cs_rcm** out)
{
    CS_REQUIRE(s && Pe && P && T0 && cp && Pr && wlob && mu && m && W && out, CS_ERR_ARG, "null argument");
    *out = nullptr;
    CS_REQUIRE(np >= 2 && nrad >= 2, CS_ERR_ARG, "need at least two cells and two radiative levels");
    CS_TRY(check_profile_args(s, nrad, Pr, nlob));
    CS_REQUIRE(nstream >= 1 && nstream <= CS_MAX_STREAMS, CS_ERR_ARG, "nstream must be in [1,%d]", CS_MAX_STREAMS);
    CS_REQUIRE(theta_s >= 0 && theta_s < CS_PI / 2, CS_ERR_ARG, "azimuth angle theta must be in [0,pi/2)");
    CS_REQUIRE(g > 0 && c_surf > 0, CS_ERR_ARG, "gravity and surface heat capacity must be positive");
    for (int64_t i = 1; i < np; i++) {
        CS_REQUIRE(Pe[i] > Pe[i - 1], CS_ERR_ARG, "cell-edge pressures must ascend (radiative_convective.jl:55-57)");
        CS_REQUIRE(P[i] > P[i - 1], CS_ERR_ARG, "cell-centre pressures must ascend");
    }
    for (int64_t i = 1; i < nrad; i++) CS_REQUIRE(Pr[i] > Pr[i - 1], CS_ERR_ARG, "radiative levels must ascend strictly");
    for (int64_t i = 0; i + 1 < np; i++) CS_REQUIRE(cp[i] > 0, CS_ERR_ARG, "heat capacities must be positive");
    cs_ctx* ctx = s->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t nnu = s->nnu;
    const int L = (int)nrad - 1;
    cs_rcm* r = new cs_rcm();
    r->ctx = ctx; r->nnu = nnu; r->np = (int)np; r->nrad = (int)nrad; r->ns = nstream;
    r->nblocks = (int)((nnu + RT_THREADS - 1) / RT_THREADS);
    r->has_beam = fS != nullptr; r->has_albedo = fa != nullptr;
    r->cs_surf = c_surf;
    const size_t smem_rt = rcm_rt_smem(r);
    if (smem_rt > 200 * 1024) {
        delete r;
        cs_set_error("too many radiative levels for the device-resident RCM (%lld)", (long long)nrad);
        return CS_ERR_ARG;
    }
    // ---- packed column block
    std::vector<double> lnPr((size_t)nrad), lnPe((size_t)np), lnP((size_t)np), dPe((size_t)np - 1), gcp((size_t)np - 1);
    for (int64_t i = 0; i < nrad; i++) lnPr[(size_t)i] = log(Pr[i]);
    for (int64_t i = 0; i < np; i++) { lnPe[(size_t)i] = log(Pe[i]); lnP[(size_t)i] = log(P[i]); }
    for (int64_t i = 0; i + 1 < np; i++) { dPe[(size_t)i] = Pe[i + 1] - Pe[i]; gcp[(size_t)i] = g / cp[i]; }
    std::vector<int> ecell((size_t)np), rcell((size_t)nrad);
    for (int64_t i = 0; i < np; i++) ecell[(size_t)i] = findcell_host(lnPr, lnPe[(size_t)i]);
    for (int64_t i = 0; i < nrad; i++) rcell[(size_t)i] = findcell_host(lnP, lnPr[(size_t)i]);
    std::vector<double> stream;
    stream.insert(stream.end(), W, W + nstream);
    for (int k = 0; k < nstream; k++) stream.push_back(1.0 / m[k]);
    // layout (doubles): F[2nrad] beamF[nrad] lnPr[nrad] lnPe[np] dPe[np-1] gcp[np-1] lnP[np] dt[1] T[np] H[np] R[np]
    //                   Fout[3nrad] rkT[nrad] Tlev[nrad] stream[2ns], then ints ecell[np] rcell[nrad]
    std::vector<double> host;
    auto put = [&](const double* p, size_t n) { size_t o = host.size(); if (p) host.insert(host.end(), p, p + n); else host.insert(host.end(), n, 0.0); return o; };
    const size_t oF = put(nullptr, 2 * (size_t)nrad), obF = put(nullptr, (size_t)nrad), olnPr = put(lnPr.data(), (size_t)nrad);
    const size_t olnPe = put(lnPe.data(), (size_t)np), odPe = put(dPe.data(), (size_t)np - 1), ogcp = put(gcp.data(), (size_t)np - 1);
    const size_t olnP = put(lnP.data(), (size_t)np), odt = put(nullptr, 1), oT = put(T0, (size_t)np), oH = put(nullptr, (size_t)np);
    const size_t oR = put(nullptr, (size_t)np), oFout = put(nullptr, 3 * (size_t)nrad), orkT = put(nullptr, (size_t)nrad);
    const size_t oTlev = put(nullptr, (size_t)nrad), ostream = put(stream.data(), stream.size());
    const size_t nd = host.size();
    r->col_bytes = nd * sizeof(double) + sizeof(int) * ((size_t)np + (size_t)nrad);
    int32_t rc = CS_OK;
    auto fail = [&](int32_t code) { cs_rcm_free(r); return code; };
    if (cs_malloc((void**)&r->col, r->col_bytes, st) != cudaSuccess) { cs_set_error("cudaMalloc(RCM column block) failed"); return fail(CS_ERR_NOMEM); }
    double* cd = (double*)r->col;
    r->F = cd + oF; r->beamF = cd + obF; r->lnPr = cd + olnPr; r->lnPe = cd + olnPe; r->dPe = cd + odPe; r->gcp = cd + ogcp;
    r->lnP = cd + olnP; r->dt = cd + odt; r->T = cd + oT; r->H = cd + oH; r->R = cd + oR; r->Fout = cd + oFout;
    r->rkT = cd + orkT; r->Tlev = cd + oTlev; r->stream = cd + ostream;
    r->ecell = (int*)(cd + nd); r->rcell = r->ecell + np;
    if ((rc = cs_stage_h2d(ctx, cd, host.data(), nd * sizeof(double))) || (rc = cs_stage_h2d(ctx, r->ecell, ecell.data(), sizeof(int) * (size_t)np)) ||
        (rc = cs_stage_h2d(ctx, r->rcell, rcell.data(), sizeof(int) * (size_t)nrad)))
        return fail(rc);
    r->dt_host = 0.0;
    // ---- per-wavenumber arrays
    const size_t bnu = sizeof(double) * (size_t)nnu;
    if (cs_malloc((void**)&r->tau, bnu * L, st) != cudaSuccess || cs_malloc((void**)&r->tr, bnu * L * nstream, st) != cudaSuccess ||
        cs_malloc((void**)&r->B_s, bnu * nrad, st) != cudaSuccess ||
        cs_malloc((void**)&r->part, sizeof(double) * (size_t)r->nblocks * 2 * nrad, st) != cudaSuccess ||
        cs_malloc((void**)&r->nu, bnu, st) != cudaSuccess || cs_malloc((void**)&r->w, bnu, st) != cudaSuccess ||
        (fS && cs_malloc((void**)&r->beam_surf, bnu, st) != cudaSuccess) || (fa && cs_malloc((void**)&r->fa, bnu, st) != cudaSuccess)) {
        cudaGetLastError();
        cs_set_error("cudaMalloc(RCM transmittance tables, %zu bytes) failed", bnu * L * (nstream + 1));
        return fail(CS_ERR_NOMEM);
    }
    CS_CUDA(cudaMemcpyAsync(r->nu, s->nu, bnu, cudaMemcpyDeviceToDevice, st));
    if (nu_weights) CS_CUDA(cudaMemcpyAsync(r->w, nu_weights, bnu, cudaMemcpyHostToDevice, st));
    else CS_CUDA(cudaMemcpyAsync(r->w, s->w, bnu, cudaMemcpyDeviceToDevice, st));
    if (fa) CS_CUDA(cudaMemcpyAsync(r->fa, fa, bnu, cudaMemcpyHostToDevice, st));
    // ---- static pass: optical depths, transmittances, stellar beam
    {
        std::vector<double> small;
        small.insert(small.end(), Pr, Pr + nrad);
        small.insert(small.end(), mu, mu + (size_t)nlob * L);
        small.insert(small.end(), wlob, wlob + nlob);
        small.insert(small.end(), m, m + nstream);
        auto al = [](size_t b) { return ((b + 255) / 256) * 256; };
        const size_t off_fS = al(small.size() * sizeof(double));
        if ((rc = ctx->s_misc.reserve(off_fS + bnu))) return fail(rc);
        char* base = ctx->s_misc.as<char>();
        if ((rc = cs_stage_h2d(ctx, base, small.data(), small.size() * sizeof(double)))) return fail(rc);
        if (fS) CS_CUDA(cudaMemcpyAsync(base + off_fS, fS, bnu, cudaMemcpyHostToDevice, st));
        RcmStaticArgs a;
        a.sig = s->sig; a.w = r->w; a.fS = fS ? (const double*)(base + off_fS) : nullptr; a.small = (const double*)base;
        a.nnu = nnu; a.np = (int)nrad; a.nlob = nlob; a.ns = nstream;
        a.Cg = 1e-4 * CS_NA / g; a.cos_s = cos(theta_s); a.tau_floor = ctx->tau_floor;
        a.tau = r->tau; a.tr = r->tr; a.beam_surf = fS ? r->beam_surf : nullptr; a.part = fS ? r->part : nullptr;
        const size_t smem = sizeof(double) * (small.size() + (size_t)nrad * RT_WARPS);
        if (smem > 48 * 1024) CS_CUDA(cudaFuncSetAttribute(rcm_static_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rcm_static_kernel<<<r->nblocks, RT_THREADS, smem, st>>>(a);
        CS_CUDA(cudaGetLastError());
        cs_count_launch(ctx, 1);
        if (fS) {
            flux_reduce_kernel<<<((int)nrad + 3) / 4, 128, 0, st>>>(r->part, r->nblocks, (int)nrad, r->beamF);
            CS_CUDA(cudaGetLastError());
            cs_count_launch(ctx, 1);
        }
    }
    if (smem_rt > 48 * 1024) {
        CS_CUDA(cudaFuncSetAttribute(rcm_rt_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rt));
        CS_CUDA(cudaFuncSetAttribute(rcm_rt_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rt));
        CS_CUDA(cudaFuncSetAttribute(rcm_rt_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rt));
        CS_CUDA(cudaFuncSetAttribute(rcm_rt_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rt));
        CS_CUDA(cudaFuncSetAttribute(rcm_rt_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rt));
        CS_CUDA(cudaFuncSetAttribute(rcm_rt_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rt));
    }
    // level temperatures of the first step
    if ((rc = rcm_enqueue_update(r, nullptr))) return fail(rc);
    cudaError_t e = cudaStreamSynchronize(st);      // the caller's host arrays are only valid during the call
    if (e != cudaSuccess) { cs_set_error("cs_rcm_create: %s", cudaGetErrorString(e)); return fail(CS_ERR_CUDA); }
    *out = r;
    return CS_OK;
}

extern "C" int32_t cs_rcm_enqueue_fluxes(cs_rcm* r, double* d_F)
{
    CS_REQUIRE(r, CS_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(r->ctx->mtx);
    CS_CUDA(cudaSetDevice(r->ctx->device));
    return rcm_enqueue_fluxes(r, d_F ? d_F : r->F);
}

extern "C" int32_t cs_rcm_enqueue_update(cs_rcm* r, const double* d_F, double dt)
{
    CS_REQUIRE(r, CS_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(r->ctx->mtx);
    CS_CUDA(cudaSetDevice(r->ctx->device));
    CS_TRY(rcm_set_dt(r, dt));
    return rcm_enqueue_update(r, d_F ? d_F : r->F);
}

static size_t rcm_mail_layout(int nranks, int n2, size_t* data_off)
{
    const size_t flags = ((sizeof(unsigned long long) * 2 * (size_t)nranks + 255) / 256) * 256;
    if (data_off) *data_off = flags;
    return flags + sizeof(double) * 2 * (size_t)nranks * (size_t)n2;
}

extern "C" int32_t cs_rcm_peer_mailbox(cs_rcm* r, int32_t nranks, void** d_mailbox, int64_t* nbytes)
{
    CS_REQUIRE(r && d_mailbox, CS_ERR_ARG, "null argument");
    CS_REQUIRE(nranks >= 1 && nranks <= RCM_MAX_RANKS, CS_ERR_ARG, "number of ranks must be in [1,%d]", RCM_MAX_RANKS);
    cs_ctx* ctx = r->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    CS_REQUIRE(!r->mail_own, CS_ERR_ARG, "the mailbox of this column already exists");
    size_t off = 0;
    const size_t bytes = rcm_mail_layout(nranks, 2 * r->nrad, &off);
    // plain cudaMalloc, not the stream-ordered pool: the block is exported to other processes (cudaIpcGetMemHandle)
    CS_CUDA(cudaMalloc((void**)&r->mail_own, bytes));
    CS_CUDA(cudaMemset(r->mail_own, 0, bytes));
    CS_CUDA(cudaMalloc((void**)&r->peer_step, 2 * sizeof(unsigned long long)));
    CS_CUDA(cudaMemset(r->peer_step, 0, 2 * sizeof(unsigned long long)));
    CS_CUDA(cudaDeviceSynchronize());
    r->mail_bytes = bytes; r->mail_data_off = off; r->peer_nranks = nranks;
    *d_mailbox = r->mail_own;
    if (nbytes) *nbytes = (int64_t)bytes;
    return CS_OK;
}

extern "C" int32_t cs_rcm_peer_connect(cs_rcm* r, int32_t rank, int32_t nranks, void* const* d_mailboxes)
{
    CS_REQUIRE(r && d_mailboxes, CS_ERR_ARG, "null argument");
    CS_REQUIRE(r->mail_own && nranks == r->peer_nranks, CS_ERR_ARG, "cs_rcm_peer_mailbox(nranks) comes first, with the same nranks");
    CS_REQUIRE(rank >= 0 && rank < nranks, CS_ERR_ARG, "rank %d outside [0,%d)", rank, nranks);
    std::lock_guard<std::recursive_mutex> lk(r->ctx->mtx);
    for (int q = 0; q < nranks; q++) {
        CS_REQUIRE(d_mailboxes[q] || q == rank, CS_ERR_ARG, "mailbox of rank %d is null", q);
        r->mail[q] = q == rank ? r->mail_own : static_cast<char*>(d_mailboxes[q]);
    }
    r->peer_rank = rank;
    return CS_OK;
}

extern "C" int32_t cs_rcm_enqueue_step_peer(cs_rcm* r, double dt)
{
    CS_REQUIRE(r, CS_ERR_ARG, "null argument");
    CS_REQUIRE(r->mail_own && r->mail[r->peer_rank] == r->mail_own, CS_ERR_ARG, "cs_rcm_peer_connect comes first");
    std::lock_guard<std::recursive_mutex> lk(r->ctx->mtx);
    CS_CUDA(cudaSetDevice(r->ctx->device));
    CS_TRY(rcm_set_dt(r, dt));
    return rcm_enqueue_step_peer(r);
}

extern "C" int32_t cs_rcm_peer_status(cs_rcm* r, int64_t* steps, int32_t* timed_out)
{
    CS_REQUIRE(r && r->peer_step, CS_ERR_ARG, "no peer exchange on this column");
    std::lock_guard<std::recursive_mutex> lk(r->ctx->mtx);
    CS_CUDA(cudaSetDevice(r->ctx->device));
    unsigned long long h[2] = {0, 0};
    CS_CUDA(cudaMemcpyAsync(h, r->peer_step, sizeof(h), cudaMemcpyDeviceToHost, r->ctx->stream));
    CS_CUDA(cudaStreamSynchronize(r->ctx->stream));
    if (steps) *steps = (int64_t)h[0];
    if (timed_out) *timed_out = (int32_t)(h[1] & 0xffffffffULL);
    return CS_OK;
}

extern "C" int32_t cs_ipc_export(void* d_ptr, uint8_t* handle64)
{
    CS_REQUIRE(d_ptr && handle64, CS_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    CS_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle64, &h, 64);
    return CS_OK;
}

extern "C" int32_t cs_ipc_open(cs_ctx* ctx, const uint8_t* handle64, void** d_ptr)
{
    CS_REQUIRE(ctx && handle64 && d_ptr, CS_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CS_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CS_OK;
}

extern "C" int32_t cs_ipc_close(cs_ctx* ctx, void* d_ptr)
{
    CS_REQUIRE(ctx && d_ptr, CS_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    CS_CUDA(cudaStreamSynchronize(ctx->stream));
    CS_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return CS_OK;
}

extern "C" int32_t cs_rcm_step(cs_rcm* r, double dt, int64_t nsteps)
{
    CS_REQUIRE(r && nsteps >= 0, CS_ERR_ARG, "bad arguments");
    cs_ctx* ctx = r->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CS_TRY(rcm_set_dt(r, dt));
    if (!r->graph_tried) {
        // one step = rt + spectral reduction + column update, captured once; dt is read from device memory so the same
        // executable graph serves every time step
        r->graph_tried = true;
        const int64_t l0 = ctx->launches;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            int32_t rc = rcm_enqueue_fluxes(r, r->F);
            if (!rc) rc = rcm_enqueue_update(r, r->F);
            cudaGraph_t gph = nullptr;
            cudaError_t e = cudaStreamEndCapture(st, &gph);
            if (!rc && e == cudaSuccess && gph && cudaGraphInstantiate(&r->exec, gph, 0) == cudaSuccess) {
                r->graph = gph;
            } else {
                if (gph) cudaGraphDestroy(gph);
                r->exec = nullptr;
                cudaGetLastError();
            }
        } else {
            cudaGetLastError();
        }
        ctx->launches = l0;      // capturing launched nothing
    }
    const int sp = cs_span_begin(ctx, CS_T_RT, true);
    for (int64_t k = 0; k < nsteps; k++) {
        if (r->exec) {
            CS_CUDA(cudaGraphLaunch(r->exec, st));
            cs_count_launch(ctx, 3);
        } else {
            CS_TRY(rcm_enqueue_fluxes(r, r->F));
            CS_TRY(rcm_enqueue_update(r, r->F));
        }
    }
    cs_span_end(ctx, sp);
    CS_CUDA(cudaStreamSynchronize(st));
    cs_spans_collect(ctx, false);
    return CS_OK;
}

extern "C" int32_t cs_rcm_set_temperature(cs_rcm* r, const double* T)
{
    CS_REQUIRE(r && T, CS_ERR_ARG, "null argument");
    cs_ctx* ctx = r->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    CS_TRY(cs_stage_h2d(ctx, r->T, T, sizeof(double) * (size_t)r->np));
    CS_TRY(rcm_enqueue_update(r, nullptr));
    return CS_OK;
}

extern "C" int32_t cs_rcm_state(cs_rcm* r, double* T, double* H, double* R, double* Fup, double* Fdn, double* Fnet)
{
    CS_REQUIRE(r, CS_ERR_ARG, "null argument");
    cs_ctx* ctx = r->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t bn = sizeof(double) * (size_t)r->np, br = sizeof(double) * (size_t)r->nrad;
    if (T) CS_CUDA(cudaMemcpyAsync(T, r->T, bn, cudaMemcpyDeviceToHost, st));
    if (H) CS_CUDA(cudaMemcpyAsync(H, r->H, bn, cudaMemcpyDeviceToHost, st));
    if (R) CS_CUDA(cudaMemcpyAsync(R, r->R, bn, cudaMemcpyDeviceToHost, st));
    if (Fup) CS_CUDA(cudaMemcpyAsync(Fup, r->Fout, br, cudaMemcpyDeviceToHost, st));
    if (Fdn) CS_CUDA(cudaMemcpyAsync(Fdn, r->Fout + r->nrad, br, cudaMemcpyDeviceToHost, st));
    if (Fnet) CS_CUDA(cudaMemcpyAsync(Fnet, r->Fout + 2 * r->nrad, br, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}

extern "C" int32_t cs_rcm_ctx(cs_rcm* r, cs_ctx** ctx)
{
    CS_REQUIRE(r && ctx, CS_ERR_ARG, "null argument");
    *ctx = r->ctx;
    return CS_OK;
}

extern "C" int32_t cs_rcm_info(cs_rcm* r, int64_t* np, int64_t* nrad, int64_t* nnu)
{
    CS_REQUIRE(r, CS_ERR_ARG, "null argument");
    if (np) *np = r->np;
    if (nrad) *nrad = r->nrad;
    if (nnu) *nnu = r->nnu;
    return CS_OK;
}

extern "C" int32_t cs_rcm_flux_buffer(cs_rcm* r, double** d_F)
{
    CS_REQUIRE(r && d_F, CS_ERR_ARG, "null argument");
    *d_F = r->F;
    return CS_OK;
}
