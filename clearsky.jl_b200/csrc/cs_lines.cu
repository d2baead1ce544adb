// cs_lines.cu -- K1 (per-line, per-level preparation) and K2 (windowed line summation).
//
// Replaces the bodies of doppler!/lorentz!/voigt!/PHCO2! and surf! of the reference
// (src/absorption/line_shapes.jl:53-87, 200-211, 313-324, 412-424, 527-540) for a whole batch of
// (T, P, Pp) nodes at once.
//
// Data layout in HBM
//   lines (SoA, sorted by wavenumber)             7 x double + int16 per line, resident per cs_lines
//   rec [level][line]  double4 {nul, a, b, thr}    fast-path record, streamed through shared memory by TMA
//   slow[level][line]  double4 {d, y, A, 0}        Voigt near-centre parameters, read straight from L2 (rare)
//   sigma[level][nu]                               output, nu fastest (= Julia sigma[:, node])
//
// K2 design (FP64 CUDA cores; the roofline that bounds it is the FP64 FMA pipe, not HBM):
//   grid = (nu tiles, levels); CTA = 8 consumer warps + 1 producer warp.  A warp owns 32*R contiguous nu
//   points (R per lane, lane-strided so loads/stores coalesce).  Lines are sorted, so the lines that can
//   touch a tile form one contiguous index range; the producer streams that range through a 4-stage
//   shared-memory ring with cp.async.bulk (TMA 1-D bulk copy) completing on mbarriers.  Each consumer warp
//   knows -- from four binary searches done once -- which lines are inside the cut-off for ALL of its
//   points (interior: no predicate), which are inside for SOME (edge: exact inclusive per-point predicate
//   |nu - nul| <= cut, line_shapes.jl:10) and which for none (culled warp-wide, never touched).
//   Far-wing Voigt (|z|^2 >= 1.6e4, >99 % of evaluations) is algebraically the Lorentz profile
//   S*gamma/(pi*(dnu^2+gamma^2)); two lines share one reciprocal:  K1/q1 + K2/q2 = (K1*q2+K2*q1)/(q1*q2).
//   Evaluations that are not safely in the far wing take the general Algorithm-985 routine.
#include "cs_internal.cuh"
#include <algorithm>

namespace {

constexpr int LS_WARPS = 8;                       // consumer warps per CTA
constexpr int LS_THREADS = (LS_WARPS + 1) * 32;   // + 1 producer warp
constexpr int LS_CHUNK = 256;                     // lines per shared-memory stage (8 KB)
constexpr int LS_STAGES = 4;

struct LevelParams {
    double T, P, Pp, scale;
    double B1, B2;   // PHCO2 chi coefficients of this level (line_shapes.jl:472,476)
    double pad0, pad1;
};

// ------------------------------------------------------------------------------------------------
// K1: per (line, level) preparation.  scaleintensity (line_shapes.jl:107-123) with chebyQrefQ (:27-48),
// alpha-doppler (:144), gamma-lorentz (:255-257); then the shape-specific record.
struct PrepArgs {
    const double *nu, *S, *ga, *gs, *Epp, *na, *mu;
    const int16_t* iso;
    const int32_t* ncheb;
    const double* cheb;
    int64_t j0, nl;            // prefiltered line range [j0, j0+nl)
    const LevelParams* lev;
    int nlev;
    double4* rec;
    double4* slow;
    int shape;
};

__global__ void __launch_bounds__(256) prep_kernel(PrepArgs a)
{
    int64_t jj = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (jj >= a.nl) return;
    int64_t j = a.j0 + jj;
    const LevelParams lp = a.lev[k];
    const double T = lp.T;
    const double nul = a.nu[j];
    // --- scaleintensity
    const double c2 = 100.0 * CS_H * CS_C / CS_KB;
    double ea = -c2 * a.Epp[j];
    double eb = -c2 * nul;
    double n = exp(ea / T) * (1 - exp(eb / T));
    double d = exp(ea / CS_TREF) * (1 - exp(eb / CS_TREF));
    int is = a.iso[j] - 1;
    const double* ch = a.cheb + (size_t)is * CS_MAXCHEB;
    int nch = a.ncheb[is];
    double tau = 2 * (T - CS_TMIN) / (CS_TMAX - CS_TMIN) - 1;
    double c1 = 1.0, c2c = tau;
    double y = ch[0] + ch[1] * c2c;
    for (int q = 2; q < nch; q++) {
        double c3 = 2 * tau * c2c - c1;
        y += ch[q] * c3;
        c1 = c2c;
        c2c = c3;
    }
    double QrefQ = 1.0 / y;
    double S = a.S[j] * QrefQ * (n / d);
    // --- widths
    double alpha = (nul / CS_C) * sqrt(2.0 * CS_R * T / a.mu[j]);
    double gamma = (pow(CS_TREF / T, a.na[j])) * (a.ga[j] * (lp.P - lp.Pp) + a.gs[j] * lp.Pp) / CS_ATM;
    const double sqpi = 1.7724538509055160273;       // sqrt(pi)
    const double sqln2 = 0.83255461115769775635;     // sqrt(log(2))
    const double osqpiln2 = 0.46971863934982566689;  // 1/sqrt(pi/log(2))
    double4 r, s;
    s = make_double4(0, 0, 0, 0);
    if (a.shape == CS_DOPPLER) {
        r = make_double4(nul, 1.0 / (alpha * alpha), S / (alpha * sqpi), 0.0);
    } else if (a.shape == CS_LORENTZ) {
        r = make_double4(nul, gamma * gamma, S * gamma / CS_PI, 0.0);
    } else {
        double beta = 1 / alpha;
        double dd = sqln2 * beta;
        // far-wing test |z|^2 >= 1.6e4 in Lorentz variables: dnu^2 + gamma^2 >= 1.6e4/d^2; the margin
        // sends the rounding-ambiguous sliver to the general routine, which decides like the reference
        double thr = (1.6e4 / (dd * dd)) * (1.0 + 1e-9);
        if (a.shape == CS_VOIGT)
            r = make_double4(nul, gamma * gamma, S * gamma / CS_PI, thr);
        else
            r = make_double4(nul, gamma, S / CS_PI, thr);
        s = make_double4(dd, gamma * dd, S * (osqpiln2 * beta), gamma);
    }
    size_t o = (size_t)k * a.nl + jj;
    a.rec[o] = r;
    if (a.slow) a.slow[o] = s;
}

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// first j in [lo,hi) for which pred(nul[j]) is false, given pred is true on a prefix
template <typename Pred> __device__ __forceinline__ int64_t first_false(const double* nul, int64_t lo, int64_t hi, Pred pred)
{
    while (lo < hi) {
        int64_t m = (lo + hi) >> 1;
        if (pred(nul[m])) lo = m + 1; else hi = m;
    }
    return lo;
}

struct LineSumArgs {
    const double* nu;
    int64_t nnu;
    const double* nul;      // prefiltered line positions [nl], ascending
    int64_t nl;
    const double4* rec;     // [nlev][nl]
    const double4* slow;    // [nlev][nl] (Voigt / PHCO2)
    const LevelParams* lev;
    double cut;
    double* out;            // [nlev][nnu]
    int accumulate;         // 0: out = scale*sigma (surf! overwrites), 1: out += scale*sigma
};

// general (any region) Voigt evaluation of one (line, point): A * Re w(dnu*d + i*y)
__device__ __noinline__ double voigt_general(const double4* __restrict__ slow, int64_t j, double dnu)
{
    double4 s = slow[j];
    return s.z * cs_faddeyeva985(dnu * s.x, s.y);
}
// PHCO2: gamma scaled by chi before forming y (line_shapes.jl:496-499)
__device__ __noinline__ double phco2_general(const double4* __restrict__ slow, int64_t j, double dnu, double chi)
{
    double4 s = slow[j];
    return s.z * cs_faddeyeva985(dnu * s.x, (chi * s.w) * s.x);
}

// chi factor of Perrin & Hartmann (line_shapes.jl:467-481), strict '<' at 3, 30, 120
__device__ __forceinline__ double chi_phco2(double adnu, double B1, double B2)
{
    if (adnu < 3.0) return 1.0;
    if (adnu < 30.0) return exp(-B1 * (adnu - 3.0));
    if (adnu < 120.0) return exp(-B1 * 27.0 - B2 * (adnu - 30.0));
    return exp(-B1 * 27.0 - B2 * 90.0 - 0.0232 * (adnu - 120.0));
}

// one (line, point) evaluation, any shape, any region; used on edge lines and for shapes without pairing
template <int SHAPE>
__device__ __forceinline__ double eval_one(const double4 rc, double dnu, const double4* __restrict__ slow,
                                           int64_t j, double B1, double B2)
{
    if (SHAPE == CS_LORENTZ) {
        double q = fma(dnu, dnu, rc.y);
        return rc.z * cs_rcp(q);
    } else if (SHAPE == CS_DOPPLER) {
        double t = dnu * dnu * rc.y;
        return (t < 746.0) ? rc.z * exp(-t) : 0.0;
    } else if (SHAPE == CS_VOIGT) {
        double q = fma(dnu, dnu, rc.y);
        if (__double2hiint(q) > __double2hiint(rc.w)) return rc.z * cs_rcp(q);
        return voigt_general(slow, j, dnu);
    } else {
        double chi = chi_phco2(fabs(dnu), B1, B2);
        double ge = chi * rc.y;
        double q = fma(dnu, dnu, ge * ge);
        if (__double2hiint(q) > __double2hiint(rc.w)) return (rc.z * ge) * cs_rcp(q);
        return phco2_general(slow, j, dnu, chi);
    }
}

template <int SHAPE, int R>
__global__ void __launch_bounds__(LS_THREADS) line_sum_kernel(LineSumArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[LS_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[LS_STAGES];
    __shared__ int64_t s_tile[2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lev = blockIdx.y;
    constexpr int TILE = LS_WARPS * 32 * R;
    const int64_t tile0 = (int64_t)blockIdx.x * TILE;
    const int64_t tile1 = min(tile0 + (int64_t)TILE, a.nnu);   // exclusive
    const double cut = a.cut;

    if (threadIdx.x == 0) {
        for (int s = 0; s < LS_STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], LS_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // lines that can touch this tile: not too far below its lowest point, not too far above its highest
        double tmin = a.nu[tile0], tmax = a.nu[tile1 - 1];
        int64_t lo = first_false(a.nul, 0, a.nl, [=](double x) { return (tmin - x) > cut; });
        int64_t hi = first_false(a.nul, lo, a.nl, [=](double x) { return !((x - tmax) > cut); });
        s_tile[0] = lo;
        s_tile[1] = hi;
    }
    __syncthreads();
    const int64_t lo = s_tile[0], hi = s_tile[1];
    const int nchunk = (int)((hi - lo + LS_CHUNK - 1) / LS_CHUNK);
    const double4* rec_lev = a.rec + (size_t)lev * a.nl;
    const double4* slow_lev = a.slow ? a.slow + (size_t)lev * a.nl : nullptr;

    if (warp == LS_WARPS) {
        // ---------------- producer warp: one elected lane feeds the ring with TMA bulk copies
        if (lane == 0) {
            for (int c = 0; c < nchunk; c++) {
                int s = c % LS_STAGES;
                uint32_t ph = (c / LS_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                int64_t c0 = lo + (int64_t)c * LS_CHUNK;
                uint32_t nrec = (uint32_t)min((int64_t)LS_CHUNK, hi - c0);
                uint32_t bytes = nrec * (uint32_t)sizeof(double4);
                mbar_arrive_expect_tx(&full_bar[s], bytes);
                tma_bulk_g2s(smem_raw + (size_t)s * LS_CHUNK * sizeof(double4), rec_lev + c0, bytes, &full_bar[s]);
            }
        }
        return;
    }

    // ---------------- consumer warps
    const int64_t wbase = tile0 + (int64_t)warp * (32 * R);
    double nup[R], acc[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = wbase + 32 * r + lane;
        valid[r] = i < a.nnu;
        nup[r] = a.nu[valid[r] ? i : (a.nnu - 1)];
        acc[r] = 0.0;
    }
    // warp's line ranges: [wlo, whi) touches some point, [ilo, ihi) is inside the cut-off for all points
    int64_t wlo = hi, whi = hi, ilo = hi, ihi = hi;
    if (wbase < a.nnu) {
        double wmin = a.nu[wbase];
        double wmax = a.nu[min(wbase + 32 * R, a.nnu) - 1];
        int64_t v = 0;
        if (lane == 0) v = first_false(a.nul, lo, hi, [=](double x) { return (wmin - x) > cut; });
        if (lane == 1) v = first_false(a.nul, lo, hi, [=](double x) { return !((x - wmax) > cut); });
        if (lane == 2) v = first_false(a.nul, lo, hi, [=](double x) { return (wmax - x) > cut; });
        if (lane == 3) v = first_false(a.nul, lo, hi, [=](double x) { return !((x - wmin) > cut); });
        wlo = __shfl_sync(0xffffffffu, v, 0);
        whi = __shfl_sync(0xffffffffu, v, 1);
        ilo = __shfl_sync(0xffffffffu, v, 2);
        ihi = __shfl_sync(0xffffffffu, v, 3);
        if (ilo >= ihi) { ilo = whi; ihi = whi; }   // window narrower than the warp's span: all edge
        ilo = min(max(ilo, wlo), whi);
        ihi = min(max(ihi, ilo), whi);
    }
    const double B1 = a.lev[lev].B1, B2 = a.lev[lev].B2;

    for (int c = 0; c < nchunk; c++) {
        int s = c % LS_STAGES;
        uint32_t ph = (c / LS_STAGES) & 1;
        const int64_t c0 = lo + (int64_t)c * LS_CHUNK;
        const int64_t c1 = min(c0 + (int64_t)LS_CHUNK, hi);
        // every warp waits for every chunk (even one it skips) so that its release below can never run
        // ahead of the ring phase
        mbar_wait(&full_bar[s], ph);
        // chunk-local line indices of this warp's three segments
        const int j0 = (int)(max(c0, wlo) - c0), j1 = (int)(min(c1, whi) - c0);
        if (j0 < j1) {
            const double4* st = reinterpret_cast<const double4*>(smem_raw + (size_t)s * LS_CHUNK * sizeof(double4));
            const int ia = (int)(min(max(ilo, c0), c1) - c0), ib = (int)(min(max(ihi, c0), c1) - c0);
            // segment A: edge lines below the interior, exact inclusive per-point predicate
            int e = min(j1, ia);
            for (int j = j0; j < e; j++) {
                double4 rc = st[j];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    double dnu = nup[r] - rc.x;
                    if (!(fabs(dnu) > cut)) acc[r] += eval_one<SHAPE>(rc, dnu, slow_lev, c0 + j, B1, B2);
                }
            }
            // segment B: interior lines, no predicate
            const int b0 = max(j0, ia), b1 = min(j1, ib);
            if (SHAPE == CS_VOIGT || SHAPE == CS_LORENTZ) {
                int j = b0;
                for (; j + 1 < b1; j += 2) {
                    double4 ra = st[j], rb = st[j + 1];
                    double qa[R], qb[R];
                    bool slowp = false;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        double da = nup[r] - ra.x, db = nup[r] - rb.x;
                        qa[r] = fma(da, da, ra.y);
                        qb[r] = fma(db, db, rb.y);
                        if (SHAPE == CS_VOIGT)
                            slowp |= (__double2hiint(qa[r]) <= __double2hiint(ra.w)) |
                                     (__double2hiint(qb[r]) <= __double2hiint(rb.w));
                    }
                    if (!slowp) {
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            double num = ra.z * qb[r];
                            num = fma(rb.z, qa[r], num);
                            double den = qa[r] * qb[r];
                            acc[r] = fma(num, cs_rcp(den), acc[r]);
                        }
                    } else {
#pragma unroll 1
                        for (int r = 0; r < R; r++) {
                            acc[r] += eval_one<SHAPE>(ra, nup[r] - ra.x, slow_lev, c0 + j, B1, B2);
                            acc[r] += eval_one<SHAPE>(rb, nup[r] - rb.x, slow_lev, c0 + j + 1, B1, B2);
                        }
                    }
                }
                if (j < b1) {
                    double4 rc = st[j];
#pragma unroll
                    for (int r = 0; r < R; r++) acc[r] += eval_one<SHAPE>(rc, nup[r] - rc.x, slow_lev, c0 + j, B1, B2);
                }
            } else {
                for (int j = b0; j < b1; j++) {
                    double4 rc = st[j];
#pragma unroll
                    for (int r = 0; r < R; r++) acc[r] += eval_one<SHAPE>(rc, nup[r] - rc.x, slow_lev, c0 + j, B1, B2);
                }
            }
            // segment C: edge lines above the interior
            for (int j = max(j0, ib); j < j1; j++) {
                double4 rc = st[j];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    double dnu = nup[r] - rc.x;
                    if (!(fabs(dnu) > cut)) acc[r] += eval_one<SHAPE>(rc, dnu, slow_lev, c0 + j, B1, B2);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    const double scale = a.lev[lev].scale;
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = wbase + 32 * r + lane;
        if (valid[r]) {
            size_t o = (size_t)lev * a.nnu + i;
            double v = scale * acc[r];
            a.out[o] = a.accumulate ? a.out[o] + v : v;
        }
    }
}

template <int SHAPE, int R> int32_t launch_line_sum(cs_ctx* ctx, const LineSumArgs& a, int nlev)
{
    constexpr int TILE = LS_WARPS * 32 * R;
    size_t smem = (size_t)LS_STAGES * LS_CHUNK * sizeof(double4);
    dim3 grid((unsigned)((a.nnu + TILE - 1) / TILE), (unsigned)nlev);
    line_sum_kernel<SHAPE, R><<<grid, LS_THREADS, smem, ctx->stream>>>(a);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    return CS_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host orchestration shared by cs_xsec, cs_bake and cs_sigma_add_lines
int32_t cs_lines_accumulate(cs_lines* L, int32_t shape, int64_t nnu, const double* d_nu, const double* h_nu,
                            int64_t nlev, const double* h_T, const double* h_P, const double* h_Pp,
                            const double* h_scale, double cut, double* d_out, int accumulate)
{
    cs_ctx* ctx = L->ctx;
    CS_REQUIRE(shape >= CS_DOPPLER && shape <= CS_PHCO2, CS_ERR_ARG, "unknown line shape id %d", shape);
    CS_REQUIRE(nnu > 0 && nlev > 0, CS_ERR_ARG, "empty wavenumber or level list");
    CS_REQUIRE(cut >= 0, CS_ERR_ARG, "negative cut-off");
    for (int64_t k = 0; k < nlev; k++) {
        // chebyQrefQ's assert (line_shapes.jl:29)
        CS_REQUIRE(h_T[k] >= CS_TMIN && h_T[k] <= CS_TMAX, CS_ERR_DOMAIN,
                   "temperature outside of Qref/Q interpolation range [%g, %g]: T = %g", CS_TMIN, CS_TMAX, h_T[k]);
    }
    // includedlines(nu::Vector, ...) strict prefilter (line_shapes.jl:18-22)
    double numin = h_nu[0], numax = h_nu[nnu - 1];
    const std::vector<double>& ln = L->h_nu;
    int64_t j0 = std::upper_bound(ln.begin(), ln.end(), numin - cut) - ln.begin();     // first nul > numin-cut
    int64_t j1 = std::lower_bound(ln.begin(), ln.end(), numax + cut) - ln.begin();     // first nul >= numax+cut
    int64_t nl = j1 > j0 ? j1 - j0 : 0;

    cudaStream_t st = ctx->stream;
    if (nl == 0) {
        if (!accumulate) CS_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)nnu * nlev, st));
        return CS_OK;
    }
    // level batches sized so that the per-level records stay within a fixed HBM budget
    size_t free_b = 0, total_b = 0;
    CS_CUDA(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = std::min<size_t>((size_t)8 << 30, free_b / 4 + ctx->s_rec.cap + ctx->s_slow.cap);
    bool need_slow = (shape == CS_VOIGT || shape == CS_PHCO2);
    size_t per_level = (size_t)nl * sizeof(double4) * (need_slow ? 2 : 1);
    int64_t nb = std::max<int64_t>(1, std::min<int64_t>(nlev, (int64_t)(budget / per_level)));
    nb = std::min<int64_t>(nb, 32768);
    CS_TRY(ctx->s_rec.reserve((size_t)nb * nl * sizeof(double4)));
    if (need_slow) CS_TRY(ctx->s_slow.reserve((size_t)nb * nl * sizeof(double4)));
    CS_TRY(ctx->s_lev.reserve(sizeof(LevelParams) * (size_t)nb));

    std::vector<LevelParams> hl((size_t)nb);
    float ms;
    for (int64_t k0 = 0; k0 < nlev; k0 += nb) {
        int64_t kb = std::min(nb, nlev - k0);
        for (int64_t k = 0; k < kb; k++) {
            LevelParams& lp = hl[(size_t)k];
            lp.T = h_T[k0 + k];
            lp.P = h_P[k0 + k];
            lp.Pp = h_Pp[k0 + k];
            lp.scale = h_scale ? h_scale[k0 + k] : 1.0;
            lp.B1 = 0.0888 - 0.16 * exp(-0.0041 * lp.T);   // line_shapes.jl:472
            lp.B2 = 0.0526 * exp(-0.00152 * lp.T);          // line_shapes.jl:476
            lp.pad0 = lp.pad1 = 0;
        }
        // pageable H2D of a few KB; synchronous with respect to the host buffer
        CS_CUDA(cudaMemcpyAsync(ctx->s_lev.p, hl.data(), sizeof(LevelParams) * (size_t)kb, cudaMemcpyHostToDevice, st));
        CS_CUDA(cudaStreamSynchronize(st));

        PrepArgs pa;
        pa.nu = L->nu; pa.S = L->S; pa.ga = L->ga; pa.gs = L->gs; pa.Epp = L->Epp; pa.na = L->na; pa.mu = L->mu;
        pa.iso = L->iso; pa.ncheb = L->ncheb; pa.cheb = L->cheb;
        pa.j0 = j0; pa.nl = nl; pa.lev = ctx->s_lev.as<LevelParams>(); pa.nlev = (int)kb;
        pa.rec = ctx->s_rec.as<double4>();
        pa.slow = need_slow ? ctx->s_slow.as<double4>() : nullptr;
        pa.shape = shape;
        CS_CUDA(cudaEventRecord(ctx->ev0, st));
        dim3 pg((unsigned)((nl + 255) / 256), (unsigned)kb);
        prep_kernel<<<pg, 256, 0, st>>>(pa);
        CS_CUDA(cudaGetLastError());
        cs_count_launch(ctx);
        CS_CUDA(cudaEventRecord(ctx->ev1, st));

        LineSumArgs la;
        la.nu = d_nu; la.nnu = nnu; la.nul = L->nu + j0; la.nl = nl;
        la.rec = pa.rec; la.slow = pa.slow; la.lev = pa.lev; la.cut = cut;
        la.out = d_out + (size_t)k0 * nnu; la.accumulate = accumulate;
        switch (shape) {
        case CS_DOPPLER: CS_TRY((launch_line_sum<CS_DOPPLER, 4>(ctx, la, (int)kb))); break;
        case CS_LORENTZ: CS_TRY((launch_line_sum<CS_LORENTZ, 4>(ctx, la, (int)kb))); break;
        case CS_VOIGT:   CS_TRY((launch_line_sum<CS_VOIGT, 4>(ctx, la, (int)kb))); break;
        default:         CS_TRY((launch_line_sum<CS_PHCO2, 4>(ctx, la, (int)kb))); break;
        }
        CS_CUDA(cudaEventRecord(ctx->ev2, st));
        CS_CUDA(cudaEventSynchronize(ctx->ev2));
        CS_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->last_kernel_ms[CS_T_PREP] += ms;
        CS_CUDA(cudaEventElapsedTime(&ms, ctx->ev1, ctx->ev2));
        ctx->last_kernel_ms[CS_T_LINESUM] += ms;
    }
    return CS_OK;
}
