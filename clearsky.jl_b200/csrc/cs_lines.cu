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
//   work unit = one warp = (tile of 32*R consecutive nu, one level); one warp per CTA, 16 CTAs per SM.  A lane owns
//   R points strided by 32 (coalesced).  Lines are sorted, so the lines that can touch a tile form one contiguous
//   index range; an elected lane streams that range through the warp's private 2-stage shared-memory ring with
//   cp.async.bulk (TMA 1-D bulk copy, 128 records per stage) completing on per-stage mbarriers.  Binary searches done
//   once per call (tile_ranges_kernel) classify every line of the window per tile: inside the cut-off for ALL points
//   and provably in the far wing (hot loop: no test of any kind), inside for SOME points (edge: exact inclusive
//   predicate |nu - nul| <= cut, line_shapes.jl:10, folded into the numerators), centre close to the tile (near:
//   Faddeyeva region decided per evaluation), outside (never touched).  All per-chunk index arithmetic is 32-bit and
//   relative to the first line of the window.
//   Far-wing Voigt (|z|^2 >= W985_S1 = 3.8e4, the 1-convergent region of Algorithm 985; >98 % of evaluations) is algebraically the Lorentz profile
//   S*gamma/(pi*(dnu^2+gamma^2)); four lines share one reciprocal: n1/d1 + n2/d2 = (n1 d2 + n2 d1)/(d1 d2), twice.
//   Evaluations that need the general Algorithm-985 routine are compacted into a per-warp queue and evaluated with
//   all lanes busy.  PHCO2 factorises chi into per-point and per-line exponentials (see the kernel).
//   Opt-in far-field expansion (cs_ctx_set_farfield): well-separated far-wing lines are summed through local expansions
//   about the tile centre, one line per lane, instead of pair by pair -- a 20-term Taylor series for Voigt / Lorentz
//   ("phase A" in the kernel), power-law series with a chi*gamma correction for the PHCO2 classes >= 30 cm^-1; the
//   direct classes above then only see what is left.
#include "cs_internal.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

#ifndef CS_LS_FOLD
#define CS_LS_FOLD 16
#endif
#ifndef CS_LS_R_EXP         // points per lane (tile = 32 R points) of the Voigt / Lorentz line sum in expansion mode
#define CS_LS_R_EXP 2
#endif
#ifndef CS_LS_NEAR_TRIM     // per-level trimming of the near range (Voigt)
#define CS_LS_NEAR_TRIM 1
#endif
#ifndef CS_LS_BAND          // Voigt: near-centre lines go through the far-wing fold, their band is corrected per point
#define CS_LS_BAND 1
#endif
#ifndef CS_LS_BAND_PREFETCH  // 0: none, 1: prefetch the near range into L2, 2: into L1, before band_near walks it
#define CS_LS_BAND_PREFETCH 0
#endif
#ifndef CS_LS_EDGE_STAGES   // COLD launch: cut-off edges slice by slice (edge lines folded only for the 32-point slices they reach)
#define CS_LS_EDGE_STAGES 1
#endif
#ifndef CS_LS_PHCO2_FOLD    // lines per reciprocal in the factorised chi classes of PHCO2 (0: pairs)
#define CS_LS_PHCO2_FOLD 8
#endif
#ifndef CS_LS_SPLIT         // direct mode, Voigt (band) / Lorentz: cold classes in line_sum_kernel<.., COLD>, far wings in far_fold_kernel
#define CS_LS_SPLIT 1
#endif
#ifndef CS_LS_COLD_MINBLK   // resident one-warp CTAs per SM of line_sum_kernel<.., COLD> (latency-bound: more warps, fewer registers)
#define CS_LS_COLD_MINBLK 16
#endif
#ifndef CS_LS_HOT_MINBLK    // resident one-warp CTAs per SM of far_fold_kernel
#define CS_LS_HOT_MINBLK 16
#endif
#ifndef CS_LS_MINBLK
#define CS_LS_MINBLK 16
#endif
#ifndef CS_LS_WARPS
#define CS_LS_WARPS 1
#endif
constexpr int LS_WARPS = CS_LS_WARPS;                       // warps per CTA; every warp is an independent work unit
constexpr int LS_THREADS = LS_WARPS * 32;
// 128 lines x 2 stages measured 2 % faster than 64 x 4 and 96 x 3 on C2 (same 8 KB per warp: fewer chunk prologues)
#ifndef CS_LS_CHUNK
#define CS_LS_CHUNK 128
#endif
#ifndef CS_LS_STAGES
#define CS_LS_STAGES 2
#endif
constexpr int LS_CHUNK = CS_LS_CHUNK;             // lines per shared-memory stage (32 B each), one ring per warp
constexpr int LS_STAGES = CS_LS_STAGES;
constexpr int LS_QCAP = 256;                      // per-warp queue of evaluations that need the general routine
constexpr int MP_P = 20;                          // order of the far-field expansion
constexpr double MP_THETA = 4.0;                  // separation (in half tile widths) beyond which lines are expanded
// PHCO2 far wings (|dnu| >= 30 cm^-1 for every point of the tile): orders of the three power-law series (see the kernel)
constexpr int PX_P2 = 10, PX_P4 = 6, PX_P6 = 4;     // classes 30-120: |r| <= 1/16
constexpr int PX_Q2 = 7, PX_Q4 = 5, PX_Q6 = 3;      // class >= 120: |r| <= 1/61, the series can stop earlier
constexpr double PX_HMAX = 2.0;                   // half tile widths above this keep the pair-by-pair sum (ratio h/|u| <= 1/16)
constexpr int LS_NSEG = 5;                        // direct-sum segments of the chunk stream

struct LevelParams {
    double T, P, Pp, scale;
    double B1, B2;   // PHCO2 chi coefficients of this level (line_shapes.jl:472,476)
    double cnear;    // lines with |nul - nu| > cnear*nul are safely in the far wing at this level (< 0: no near range)
    double pexp_ok;  // PHCO2: 1 when (chi*gamma/dnu)^2 < 1e-4 for every line with |dnu| >= 30 at this level (expansion allowed)
    double lgtr;     // log(296/T): (296/T)^na = exp(na lgtr) in K1 (one exp instead of a pow per line and level)
};

// ------------------------------------------------------------------------------------------------
// K1: per (line, level) preparation.  scaleintensity (line_shapes.jl:107-123) with chebyQrefQ (:27-48),
// alpha-doppler (:144), gamma-lorentz (:255-257); then the shape-specific record.
struct PrepArgs {
    const double *nu, *S, *ga, *gs, *Epp, *na, *mu;
    const double* dref;        // per line: exp(-c2 E''/296) (1 - exp(-c2 nul/296)), the level-independent denominator of scaleintensity
    const double* qtab;        // [nlev][niso] Qref/Q(T) (chebyQrefQ depends on the isotopologue and the level only)
    int niso;
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
    double d = a.dref[j];                                  // level-independent: computed once per line list (line_static_kernel)
    int is = a.iso[j] - 1;
    double QrefQ = a.qtab[(size_t)k * a.niso + is];        // per (isotopologue, level): qrefq_kernel
    double S = a.S[j] * QrefQ * (n / d);
    // --- widths
    double alpha = (nul / CS_C) * sqrt(2.0 * CS_R * T / a.mu[j]);
    double gamma = exp(a.na[j] * lp.lgtr) * (a.ga[j] * (lp.P - lp.Pp) + a.gs[j] * lp.Pp) / CS_ATM;
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
        // |z|^2 = x^2 + y^2 = d^2 (dnu^2 + gamma^2): the record carries d^2 so that the Faddeyeva region can be
        // decided from Lorentz variables; ambiguous slivers go to the general routine (decides like the reference)
        if (a.shape == CS_VOIGT)
            r = make_double4(nul, gamma * gamma, S * gamma / CS_PI, dd * dd);
        else
            r = make_double4(nul, gamma, S / CS_PI, dd * dd);
        s = make_double4(dd, gamma * dd, S * (osqpiln2 * beta), gamma);
    }
    size_t o = (size_t)k * a.nl + jj;
    a.rec[o] = r;
    if (a.slow) a.slow[o] = s;
}

// level-independent part of scaleintensity (line_shapes.jl:122), once per line list
__global__ void __launch_bounds__(256) line_static_kernel(const double* __restrict__ nu, const double* __restrict__ Epp, int64_t n,
                                                          double* __restrict__ dref)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double c2 = 100.0 * CS_H * CS_C / CS_KB;
    double ea = -c2 * Epp[j], eb = -c2 * nu[j];
    dref[j] = exp(ea / CS_TREF) * (1 - exp(eb / CS_TREF));
}

// chebyQrefQ (line_shapes.jl:27-48) per (level, isotopologue): forward recurrence starting from a[1] + a[2] tau
__global__ void qrefq_kernel(const LevelParams* __restrict__ lev, const int32_t* __restrict__ ncheb, const double* __restrict__ cheb,
                             int niso, double* __restrict__ qtab)
{
    const int is = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (is >= niso) return;
    const double T = lev[k].T;
    const double* ch = cheb + (size_t)is * CS_MAXCHEB;
    const int nch = ncheb[is];
    double tau = 2 * (T - CS_TMIN) / (CS_TMAX - CS_TMIN) - 1;
    double c1 = 1.0, c2c = tau;
    double y = ch[0] + ch[1] * c2c;
    for (int q = 2; q < nch; q++) {
        double c3 = 2 * tau * c2c - c1;
        y += ch[q] * c3;
        c1 = c2c;
        c2c = c3;
    }
    qtab[(size_t)k * niso + is] = 1.0 / y;
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

// read-only global load of one 32-byte record (two 16-byte loads through the non-coherent path)
__device__ __forceinline__ double4 ld_rec(const double4* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 lo = __ldg(q), hi = __ldg(q + 1);
    return make_double4(lo.x, lo.y, hi.x, hi.y);
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
    int accumulate;         // 0: out = scale*sigma (surf! overwrites), 1: out += scale*sigma, 2: out = log(scale*sigma) (bake: the
                            // table fit wants ln sigma, gases.jl:75-81, and one log per output point is free here)
    int64_t ntiles;
    int nr;                 // entries per tile in ranges (6, 8 with the far-field expansion, or LS_NR for PHCO2)
    const int64_t* ranges;  // [ntiles][nr], see tile_ranges_kernel
    double mp_theta;        // > 0: far-field expansion for lines farther than mp_theta half tile widths (Voigt, Lorentz)
    const double* ffc;      // far-field coefficients [nlev][ntiles][MP_P] precomputed by farfield_kernel (or null)
    double nul_lo, nul_hi;  // first / last prefiltered line position (host copy)
    double near_cn;         // near-centre fraction the per-tile ranges were built with (maximum over the level batch)
    int band;               // Voigt: 1 = per-point band correction of the near-centre lines (host: damping parameter bounded below)
    int split;              // 1 = two launches: cold classes (this kernel, COLD) then far_fold_kernel over the all-inside lines
    int fold_g;             // lines per reciprocal in far_fold_kernel (32, 16 or 4: the host bounds the product of G values q)
    const double2* chix;    // PHCO2 expansion: {X, 1/X}, X = exp(0.0232 (nul - chix_ref)) per prefiltered line (or null)
    double chix_ref;
};

// Voigt evaluation of one (line, point) that is not safely in the far wing: decides the region exactly like
// the reference's faddeyeva(x, y) (same s = fma(x,x,y^2)), with the 1- and 2-convergent forms inlined.
__device__ __noinline__ double voigt_near(const double4* __restrict__ slow, int64_t j, double dnu, double chi)
{
    const double osqpi = 0.56418958354775628695;
    double4 s = slow[j];
    double x = dnu * s.x;
    double y = (chi * s.w) * s.x;          // (chi*gamma)*d ; chi == 1 for plain Voigt (line_shapes.jl:373,498)
    double y2 = y * y;
    double sq = fma(x, x, y2);
    if (sq >= W985_S1) return s.z * (y * osqpi / sq);
    if (sq >= W985_S2) {
        // Re[i z/(sqrt(pi)(z^2-1/2))] = y (s+1/2) / (sqrt(pi) ((s-1/2)^2 + 2 y^2))
        double sm = sq - 0.5;
        return s.z * (osqpi * y * (sq + 0.5) / fma(sm, sm, 2.0 * y2));
    }
    return s.z * cs_faddeyeva985(x, y);
}

// chi factor of Perrin & Hartmann (line_shapes.jl:467-481), strict '<' at 3, 30, 120
__device__ __forceinline__ double chi_phco2(double adnu, double B1, double B2)
{
    if (adnu < 3.0) return 1.0;
    if (adnu < 30.0) return exp(-B1 * (adnu - 3.0));
    if (adnu < 120.0) return exp(-B1 * 27.0 - B2 * (adnu - 30.0));
    return exp(-B1 * 27.0 - B2 * 90.0 - 0.0232 * (adnu - 120.0));
}

// hi-word thresholds of |z|^2 with a 1e-6 guard band (the hi word of a double resolves 2^-20 ~ 1e-6):
//   hi(s) >  S1_HI              =>  s > S1 (1+1e-6)   : 1 convergent, certainly          (S1, S2 = W985_S1, W985_S2)
//   S2_HI < hi(s) < S1_LO       =>  S2 (1+1e-6) < s < S1 (1-1e-6) : 2 convergents, certainly
// anything else is decided by the general routine with the reference's own s = fma(x,x,y^2).
#if CS_W985_MAP
#define CS_S1_HI 0x40E28E03   /* hi word of 3.8e4*(1+1.0e-6) rounded up   (3.8e4 = 0x40E28E00 00000000) */
#define CS_S1_LO 0x40E28DFC   /* hi word of 3.8e4*(1-1.0e-6) rounded down */
#define CS_S2_HI 0x40700002   /* hi word of 256*(1+1.0e-6) rounded up     (256 = 0x40700000 00000000) */
#else
#define CS_S1_HI 0x40CF4004   /* hi word of 1.6e4*(1+1.0e-6) rounded up   (1.6e4 = 0x40CF4000 00000000) */
#define CS_S1_LO 0x40CF3FFB   /* hi word of 1.6e4*(1-1.0e-6) rounded down */
#define CS_S2_HI 0x40640002   /* hi word of 160*(1+1.0e-6) rounded up     (160 = 0x40640000 00000000) */
#endif

// branch-free choice between the 1- and 2-convergent forms, both written in Lorentz variables:
//   1: K/q          2: K d^2 (s+1/2) / ((s-1/2)^2 + 2 y^2),   s = d^2 q,  y^2 = d^2 gamma^2
// need = true when neither form is certain (guard bands, |z|^2 < W985_S2): the caller must use the general routine
__device__ __forceinline__ double voigt_12(const double4 rc, double dnu, bool& need)
{
    double q = fma(dnu, dnu, rc.y);
    double s = rc.w * q;
    int hs = __double2hiint(s);
    bool one = hs > CS_S1_HI;
    bool two = (hs < CS_S1_LO) & (hs > CS_S2_HI);
    double sm = s - 0.5;
    double den2 = fma(sm, sm, 2.0 * (rc.w * rc.y));
    double num2 = (rc.z * rc.w) * (s + 0.5);
    need = !(one | two);
    return (one ? rc.z : num2) * cs_rcp(one ? q : den2);
}

// one (line, point) evaluation, any shape, any Faddeyeva region (edge and near-centre lines)
template <int SHAPE>
__device__ __forceinline__ double eval_checked(const double4 rc, double dnu, const double4* __restrict__ slow,
                                               int64_t j, double B1, double B2)
{
    if (SHAPE == CS_LORENTZ) {
        double q = fma(dnu, dnu, rc.y);
        return rc.z * cs_rcp(q);
    } else if (SHAPE == CS_DOPPLER) {
        double t = dnu * dnu * rc.y;
        return (t < 746.0) ? rc.z * exp(-t) : 0.0;   // exp(-t) is exactly 0 beyond (as in the reference)
    } else if (SHAPE == CS_VOIGT) {
        bool need;
        double v = voigt_12(rc, dnu, need);
        if (need) v = voigt_near(slow, j, dnu, 1.0);
        return v;
    } else {
        double chi = chi_phco2(fabs(dnu), B1, B2);
        double ge = chi * rc.y;
        double q = fma(dnu, dnu, ge * ge);
        if (__double2hiint(rc.w * q) > CS_S1_HI) return (rc.z * ge) * cs_rcp(q);
        return voigt_near(slow, j, dnu, chi);
    }
}

// PHCO2 expansion: the chi factor of the >= 120 cm^-1 class, exp(-+0.0232 (nu0 - nul)), split into a per-tile scalar and a
// per-line factor that depends on neither the tile nor the level: one exp per line per call instead of one per line per
// (tile, level).  ref keeps the arguments small (|nul - ref| <= half the span of the line list).
__global__ void chix_kernel(const double* __restrict__ nul, int64_t nl, double ref, double2* __restrict__ chix)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nl) return;
    double x = exp(0.0232 * (nul[j] - ref));
    chix[j] = make_double2(x, 1.0 / x);
}

// per-tile line classification, computed once per call (it does not depend on the level):
// ranges[tile][0..5] = wlo, whi, ilo, ihi, nlo, nhi over the prefiltered sorted line positions
//   [wlo,whi): lines within the cut-off of SOME point of the tile (exact FP64 rule of line_shapes.jl:10)
//   [ilo,ihi): lines within the cut-off of ALL points (no per-point predicate needed)
//   [nlo,nhi): lines whose centre is within cn*nul of the tile: the far-wing form may not apply there
constexpr int LS_NR = 18;   // entries per tile (6 general + 12 PHCO2 chi-class boundaries)
constexpr int LS_NR_SPLIT = 12;   // split direct mode with 128-point tiles: 6 general + 2 x 3 slice borders of the cut-off edges
__global__ void tile_ranges_kernel(const double* __restrict__ nu, int64_t nnu, const double* __restrict__ nul,
                                   int64_t nl, double cut, double cn, int tile_pts, int64_t ntiles, int nr,
                                   double mp_theta, int64_t* __restrict__ ranges)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles * nr) return;
    int64_t tile = t / nr;
    int k = (int)(t % nr);
    int64_t i0 = tile * tile_pts, i1 = min(i0 + (int64_t)tile_pts, nnu);
    const double tmin = nu[i0], tmax = nu[i1 - 1];
    int64_t v;
    if (nr == 8 && k >= 6) {
        // far-field expansion borders: lines at least mp_theta half-widths away from the tile centre
        const double cen = 0.5 * (tmin + tmax), D = mp_theta * (0.5 * (tmax - tmin));
        if (k == 6) v = first_false(nul, 0, nl, [=](double x) { return x <= cen - D; });
        else        v = first_false(nul, 0, nl, [=](double x) { return x < cen + D; });
        ranges[t] = v;
        return;
    }
    if (nr == LS_NR_SPLIT && k >= 6) {
        // split direct mode: cut-off borders of the 32-point slices of the tile (the COLD launch sums the edge lines slice by
        // slice).  Entries 6.. : b_r, r = 1..S-1 = first line partially inside slice r (b_0 = entry 0); then d_r, r = 0..S-2 =
        // first line beyond slice r (d_{S-1} = entry 1); the same FP64 predicates as entries 0 and 1 on the slice's end points
        const int S = tile_pts / 32;
        if (k < 6 + (S - 1)) {
            const int r = k - 5;
            const double smin = nu[min(i0 + 32 * (int64_t)r, i1 - 1)];
            v = first_false(nul, 0, nl, [=](double x) { return (smin - x) > cut; });
        } else {
            const int r = k - (6 + (S - 1));
            const double smax = nu[min(i0 + 32 * (int64_t)r + 31, i1 - 1)];
            v = first_false(nul, 0, nl, [=](double x) { return !((x - smax) > cut); });
        }
        ranges[t] = v;
        return;
    }
    switch (k) {
    case 0: v = first_false(nul, 0, nl, [=](double x) { return (tmin - x) > cut; }); break;
    case 1: v = first_false(nul, 0, nl, [=](double x) { return !((x - tmax) > cut); }); break;
    case 2: v = first_false(nul, 0, nl, [=](double x) { return (tmax - x) > cut; }); break;
    case 3: v = first_false(nul, 0, nl, [=](double x) { return !((x - tmin) > cut); }); break;
    case 4: v = first_false(nul, 0, nl, [=](double x) { return x * (1.0 + cn) < tmin; }); break;
    case 5: v = first_false(nul, 0, nl, [=](double x) { return !(x * (1.0 - cn) > tmax); }); break;
    // PHCO2 chi classes (line_shapes.jl:467-481, strict '<' at 3, 30, 120), decided for ALL points of the tile with
    // the same FP subtraction chi itself uses.  Lines below the tile: |dnu| = nu_p - x in [tmin - x, tmax - x].
    case 6: v = first_false(nul, 0, nl, [=](double x) { return (tmin - x) >= 120.0; }); break;      // end of ">=120 for all"
    case 7: v = first_false(nul, 0, nl, [=](double x) { return !((tmax - x) < 120.0); }); break;    // start of "<120 for all"
    case 8: v = first_false(nul, 0, nl, [=](double x) { return (tmin - x) >= 30.0; }); break;
    case 9: v = first_false(nul, 0, nl, [=](double x) { return !((tmax - x) < 30.0); }); break;
    case 10: v = first_false(nul, 0, nl, [=](double x) { return (tmin - x) >= 3.0; }); break;
    case 11: v = first_false(nul, 0, nl, [=](double x) { return !((tmax - x) < 3.0); }); break;      // start of "<3 for all"
    // lines above the tile: |dnu| = x - nu_p in [x - tmax, x - tmin]
    case 12: v = first_false(nul, 0, nl, [=](double x) { return (x - tmin) < 3.0; }); break;         // end of "<3 for all"
    case 13: v = first_false(nul, 0, nl, [=](double x) { return !((x - tmax) >= 3.0); }); break;     // start of ">=3 for all"
    case 14: v = first_false(nul, 0, nl, [=](double x) { return (x - tmin) < 30.0; }); break;
    case 15: v = first_false(nul, 0, nl, [=](double x) { return !((x - tmax) >= 30.0); }); break;
    case 16: v = first_false(nul, 0, nl, [=](double x) { return (x - tmin) < 120.0; }); break;
    default: v = first_false(nul, 0, nl, [=](double x) { return !((x - tmax) >= 120.0); }); break;
    }
    ranges[t] = v;
}

// ---- cold paths of K2 (edge lines, near-centre lines, deferred general evaluations).  They are kept out of line
// and accumulate into the warp's shared-memory accumulators (each lane only touches its own R slots, except for
// the deferred pass which uses shared atomics), so that the register allocation and instruction schedule of the
// hot far-wing loop are not polluted by them.
// per-warp shared memory behind the ring: nutile[TILE], cacc[TILE], queue[LS_QCAP]; PHCO2 adds the chi tables
// Etab[6][TILE] (3 classes x 2 sides), the per-chunk line factors Fp[LS_CHUNK] and the 18 segment borders
template <int SHAPE, int R> __host__ __device__ constexpr size_t ls_extra_bytes()
{
    return (size_t)2 * 32 * R * sizeof(double) + LS_QCAP * sizeof(uint32_t) +
           (SHAPE == CS_PHCO2 ? (size_t)6 * 32 * R * sizeof(double) + LS_CHUNK * sizeof(double) + 18 * sizeof(int64_t) : 0);
}

struct WarpCold {
    const double* nutile;      // [32*R] wavenumbers of this warp's tile (shared memory)
    double* cacc;              // [32*R] accumulators of the cold paths (shared memory)
    uint32_t* queue;           // [LS_QCAP] deferred (line, point) pairs
    const double4* slow_lev;   // near-centre parameters of this level, from the first line of the window (global)
    double cut, B1, B2;
    int lane;
    bool edge_is_far;
};

template <int SHAPE, int R>
__device__ __noinline__ void cold_flush(const WarpCold& w, const double4* st, int c0, int qn)
{
    for (int e = w.lane; e < qn; e += 32) {
        uint32_t en = w.queue[e];
        int j = (int)(en >> 8), p = (int)(en & 255u);
        double4 rc = st[j];
        atomicAdd(&w.cacc[p], voigt_near(w.slow_lev, c0 + j, w.nutile[p] - rc.x, 1.0));
    }
    __syncwarp();
}

// ---- the far-wing fold.  G lines per reciprocal, folded left to right: (n, d) <- (n q + K d, d q) costs 3 FP64 operations
// per line (like a pairwise merge of fractions) but keeps ONE running fraction per point, so the group can be long:
// 2 G + 3 (G - 1) + 4 operations per G evaluations = 5.06 per evaluation at G = 16 (5.25 for a 2 x 2 tree), with one MUFU
// seed per 16 evaluations instead of one per 4.  d is a product of G values q >= gamma^2 (and >= W985_S1 alpha^2 / ln 2 for
// Voigt): far from underflow for any physical list.
template <int R, int G>
__device__ __forceinline__ void fold_run(const double4* __restrict__ st, int& j, int f1, const double (&nup)[R], double (&acc)[R])
{
    for (; j + G - 1 < f1; j += G) {
        double fn[R], fd[R];
        {
            const double4 ra = st[j];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const double da = nup[r] - ra.x;
                fd[r] = fma(da, da, ra.y);
                fn[r] = ra.z;
            }
        }
#pragma unroll
        for (int g = 1; g < G; g++) {
            const double4 rb = st[j + g];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const double db = nup[r] - rb.x;
                const double qb = fma(db, db, rb.y);
                fn[r] = fma(rb.z, fd[r], fn[r] * qb);
                fd[r] *= qb;
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = fma(fn[r], cs_rcp(fd[r]), acc[r]);
    }
}
// the last, shorter group of a range: same fold, rolled (5 FP64 operations per evaluation + loop)
template <int R>
__device__ __forceinline__ void fold_tail(const double4* __restrict__ st, int& j, int f1, const double (&nup)[R], double (&acc)[R])
{
    if (j >= f1) return;
    double fn[R], fd[R];
    {
        const double4 ra = st[j];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const double da = nup[r] - ra.x;
            fd[r] = fma(da, da, ra.y);
            fn[r] = ra.z;
        }
    }
#pragma unroll 1
    for (j++; j < f1; j++) {
        const double4 rb = st[j];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const double db = nup[r] - rb.x;
            const double qb = fma(db, db, rb.y);
            fn[r] = fma(rb.z, fd[r], fn[r] * qb);
            fd[r] *= qb;
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = fma(fn[r], cs_rcp(fd[r]), acc[r]);
}

// edge lines: exact inclusive per-point predicate (line_shapes.jl:10)
template <int SHAPE, int R>
__device__ __noinline__ void cold_edge(const WarpCold& w, const double4* st, int c0, int e0, int e1)
{
    const int lane = w.lane;
    const double cut = w.cut;
    double nup[R], acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) { nup[r] = w.nutile[32 * r + lane]; acc[r] = 0.0; }
    int j = e0;
    if ((SHAPE == CS_VOIGT || SHAPE == CS_LORENTZ) && w.edge_is_far) {
        // far-wing form with the inclusive predicate folded into the numerators (K -> 0 outside the window): no vote,
        // no branch, so the R chains of a lane interleave exactly like in the hot loop
        for (; j + 3 < e1; j += 4) {
            double4 ra = st[j], rb = st[j + 1], rc = st[j + 2], rd = st[j + 3];
#pragma unroll
            for (int r = 0; r < R; r++) {
                double da = nup[r] - ra.x, db = nup[r] - rb.x, dc = nup[r] - rc.x, dd = nup[r] - rd.x;
                double ka = (fabs(da) > cut) ? 0.0 : ra.z, kb = (fabs(db) > cut) ? 0.0 : rb.z;
                double kc = (fabs(dc) > cut) ? 0.0 : rc.z, kd = (fabs(dd) > cut) ? 0.0 : rd.z;
                double qa = fma(da, da, ra.y), qb = fma(db, db, rb.y);
                double qc = fma(dc, dc, rc.y), qd = fma(dd, dd, rd.y);
                double n1 = fma(kb, qa, ka * qb), d1 = qa * qb;
                double n2 = fma(kd, qc, kc * qd), d2 = qc * qd;
                acc[r] = fma(fma(n2, d1, n1 * d2), cs_rcp(d1 * d2), acc[r]);
            }
        }
        for (; j < e1; j++) {
            double4 rc = st[j];
#pragma unroll
            for (int r = 0; r < R; r++) {
                double dnu = nup[r] - rc.x;
                acc[r] = fma((fabs(dnu) > cut) ? 0.0 : rc.z, cs_rcp(fma(dnu, dnu, rc.y)), acc[r]);
            }
        }
    } else {
        for (; j < e1; j++) {
            double4 rc = st[j];
#pragma unroll
            for (int r = 0; r < R; r++) {
                double dnu = nup[r] - rc.x;
                if (!(fabs(dnu) > cut)) acc[r] += eval_checked<SHAPE>(rc, dnu, w.slow_lev, c0 + j, w.B1, w.B2);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) w.cacc[32 * r + lane] += acc[r];
}

// near lines: inside the cut-off for every point, Faddeyeva region decided per evaluation.  Returns the new
// queue length.
template <int SHAPE, int R>
__device__ __noinline__ int cold_near(const WarpCold& w, const double4* st, int c0, int n0, int n1, int qn)
{
    const int lane = w.lane;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t qaddr = smem_u32(w.queue);
    double nup[R], acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) { nup[r] = w.nutile[32 * r + lane]; acc[r] = 0.0; }
    for (int j = n0; j < n1; j++) {
        double4 rc = st[j];
        if (SHAPE == CS_VOIGT) {
            if (qn > LS_QCAP - 32 * R) { cold_flush<SHAPE, R>(w, st, c0, qn); qn = 0; }
            // all R evaluations first (independent chains), then the deferral bookkeeping without divergent branches:
            // (line, point) pairs that need the general routine are compacted in (line, slice, lane) order
            bool need[R];
            unsigned m[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                double v = voigt_12(rc, nup[r] - rc.x, need[r]);
                acc[r] += need[r] ? 0.0 : v;
            }
#pragma unroll
            for (int r = 0; r < R; r++) m[r] = __ballot_sync(0xffffffffu, need[r]);
            // most lines defer nothing in most 32-point slices: a slice with an empty ballot skips its bookkeeping
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (m[r]) {
                    if (need[r]) {
                        const uint32_t en = ((uint32_t)j << 8) | (uint32_t)(32 * r + lane);
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + 4u * (uint32_t)(qn + __popc(m[r] & lt_mask))), "r"(en)
                                     : "memory");
                    }
                    qn += __popc(m[r]);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] += eval_checked<SHAPE>(rc, nup[r] - rc.x, w.slow_lev, c0 + j, w.B1, w.B2);
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) w.cacc[32 * r + lane] += acc[r];
    __syncwarp();
    return qn;
}

// ---- Voigt near band, per point.  The per-tile near range [nlo,nhi) holds every line whose Doppler zone (|z|^2 < W985_S1)
// touches SOME point of the tile; for one point only the lines with nul (1 - cn) <= nu <= nul (1 + cn) can be inside their
// zone -- ~1/4 of the (line, point) pairs of the range on the C2 grid.  So the whole range goes through the far-wing fold
// like any other line (K/q for every pair, no test), and each lane walks the band of ITS points (two binary searches per
// point, records read straight from L1/L2: the lanes read different lines) adding the difference to the true value:
//   2 convergents:  K d^2 (s+1/2)/D - K/q = K (3/2 s - 1/4 - 2 y^2)/(q D),  D = (s-1/2)^2 + 2 y^2   (no cancellation)
//   general routine (|z|^2 < W985_S2 and the guard slivers): w985 value - K/q, deferred to the queue as before.
// K/q exceeds the true value by at most 1/(sqrt(pi) y) (at the line centre), so the subtraction costs log10 of that in
// digits: the host only selects this path when y >= 1e-5 for every line of every level of the batch.
template <int R>
__device__ __noinline__ void band_flush(const WarpCold& w, const double4* __restrict__ rec_near,
                                        const double4* __restrict__ slow_near, int qn)
{
    for (int e = w.lane; e < qn; e += 32) {
        const uint32_t en = w.queue[e];
        const int j = (int)(en >> 8), p = (int)(en & 255u);
        const double4 rc = ld_rec(rec_near + j);
        const double dnu = w.nutile[p] - rc.x;
        atomicAdd(&w.cacc[p], voigt_near(slow_near, j, dnu, 1.0) - rc.z * cs_rcp(fma(dnu, dnu, rc.y)));
    }
    __syncwarp();
}

template <int R>
__device__ __noinline__ void band_near(const WarpCold& w, const double4* __restrict__ rec_near,
                                       const double4* __restrict__ slow_near, const double* __restrict__ nul_near, int nn, double cn)
{
    const int lane = w.lane;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t qaddr = smem_u32(w.queue);
    const double cp = 1.0 + cn, cm = 1.0 - cn;
#if CS_LS_BAND_PREFETCH
    // the records were written by K1 gigabytes ago: pull the near range (positions, records, near-centre parameters) towards the
    // SM before the searches and the walk touch it line by line
    for (int i = lane; 4 * i < nn; i += 32) {
#if CS_LS_BAND_PREFETCH == 2
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rec_near + 4 * i));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(slow_near + 4 * i));
#else
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rec_near + 4 * i));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(slow_near + 4 * i));
#endif
    }
    for (int i = lane; 16 * i < nn; i += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(nul_near + 16 * i));
#endif
    double nup[R], acc[R];
    int jlo[R], len[R];
    int maxlen = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        nup[r] = w.nutile[32 * r + lane];
        acc[r] = 0.0;
    }
    // band of each point: [first line with !(nul (1+cn) < nu), first line with nul (1-cn) > nu) -- the same two predicates the
    // per-tile range was built with (tile_ranges_kernel, entries 4 and 5), so a pair outside the band is certainly far wing.
    // Both are true on a prefix of the sorted positions: two branch-free binary searches per point (counts of the prefixes),
    // all 2 R loads of a step in flight together
    {
        int c1[R], c2[R];
#pragma unroll
        for (int r = 0; r < R; r++) { c1[r] = 0; c2[r] = 0; }
        int step = 1;
        while (2 * step <= nn) step *= 2;
        for (; step > 0; step >>= 1) {
            double x1[R], x2[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                x1[r] = __ldg(nul_near + min(c1[r] + step, nn) - 1);
                x2[r] = __ldg(nul_near + min(c2[r] + step, nn) - 1);
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (c1[r] + step <= nn && x1[r] * cp < nup[r]) c1[r] += step;
                if (c2[r] + step <= nn && !(x2[r] * cm > nup[r])) c2[r] += step;
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) { jlo[r] = c1[r]; len[r] = max(c2[r] - c1[r], 0); maxlen = max(maxlen, len[r]); }
    }
    maxlen = __reduce_max_sync(0xffffffffu, maxlen);
    int qn = 0;
    for (int k = 0; k < maxlen; k++) {
        if (qn > LS_QCAP - 32 * R) { band_flush<R>(w, rec_near, slow_near, qn); qn = 0; }
        bool need[R];
        int jj[R];
        unsigned m[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const bool valid = k < len[r];
            jj[r] = min(jlo[r] + k, nn - 1);
            const double4 rc = ld_rec(rec_near + jj[r]);
            const double dnu = nup[r] - rc.x;
            const double q = fma(dnu, dnu, rc.y);
            const double s = rc.w * q;
            const int hs = __double2hiint(s);
            const bool one = hs > CS_S1_HI;
            const bool two = (hs < CS_S1_LO) & (hs > CS_S2_HI);
            const double y2 = rc.w * rc.y;
            const double sm = s - 0.5;
            const double D = fma(sm, sm, 2.0 * y2);
            const double num = rc.z * fma(-2.0, y2, fma(1.5, s, -0.25));
            const double v = num * cs_rcp(q * D);
            acc[r] += (valid & two) ? v : 0.0;
            need[r] = valid & !(one | two);
        }
#pragma unroll
        for (int r = 0; r < R; r++) m[r] = __ballot_sync(0xffffffffu, need[r]);
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (m[r]) {
                if (need[r]) {
                    const uint32_t en = ((uint32_t)jj[r] << 8) | (uint32_t)(32 * r + lane);
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + 4u * (uint32_t)(qn + __popc(m[r] & lt_mask))), "r"(en)
                                 : "memory");
                }
                qn += __popc(m[r]);
            }
        }
    }
    __syncwarp();
    if (qn > 0) band_flush<R>(w, rec_near, slow_near, qn);
#pragma unroll
    for (int r = 0; r < R; r++) w.cacc[32 * r + lane] += acc[r];
    __syncwarp();
}

// ---- cut-off edges, slice by slice (COLD launch, far-wing form).  A line at the lower edge of the window is inside the cut-off
// for the lowest 32-point slices of the tile only, so the edge lines [b_s, b_{s+1}) are folded for slices 0..s alone (the exact
// inclusive predicate of line_shapes.jl:10 folded into the numerators), and likewise [d_s, d_{s+1}) for slices s..R-1 at the upper
// edge: 10/16 of the (line, slice) pairs of the whole-tile edge loop, one record load per line for all active slices.
// NA slices [R0, R0+NA) of the tile, lines [j0, j1) of the stage; one reciprocal per 16 lines
template <int R, int R0, int NA>
__device__ __forceinline__ void edge_fold(const double4* __restrict__ st, int j0, int j1, double cut, const double (&nup)[R],
                                          double (&acc)[R])
{
    for (int j = j0; j < j1; j += 16) {
        const int je = min(j + 16, j1);
        double fn[NA], fd[NA];
#pragma unroll
        for (int r = 0; r < NA; r++) { fn[r] = 0.0; fd[r] = 1.0; }
#pragma unroll 2
        for (int g = j; g < je; g++) {
            const double4 rb = st[g];
#pragma unroll
            for (int r = 0; r < NA; r++) {
                const double db = nup[R0 + r] - rb.x;
                const double kb = (fabs(db) > cut) ? 0.0 : rb.z;
                const double qb = fma(db, db, rb.y);
                fn[r] = fma(kb, fd[r], fn[r] * qb);
                fd[r] *= qb;
            }
        }
#pragma unroll
        for (int r = 0; r < NA; r++) acc[R0 + r] = fma(fn[r], cs_rcp(fd[r]), acc[R0 + r]);
    }
}

template <int R, int S> struct EdgeStages {
    // lower edge: stage S = lines [b[S], b[S+1]) for slices 0..S; upper edge: stage S = lines [d[S], d[S+1]) for slices S..R-1
    static __device__ __forceinline__ void lower(const double4* st, const int* b, int c0, int n, double cut, const double (&nup)[R],
                                                 double (&acc)[R])
    {
        const int x0 = min(max(b[S] - c0, 0), n), x1 = min(max(b[S + 1] - c0, 0), n);
        if (x0 < x1) edge_fold<R, 0, S + 1>(st, x0, x1, cut, nup, acc);
        if constexpr (S + 1 < R) EdgeStages<R, S + 1>::lower(st, b, c0, n, cut, nup, acc);
    }
    static __device__ __forceinline__ void upper(const double4* st, const int* d, int c0, int n, double cut, const double (&nup)[R],
                                                 double (&acc)[R])
    {
        const int x0 = min(max(d[S] - c0, 0), n), x1 = min(max(d[S + 1] - c0, 0), n);
        if (x0 < x1) edge_fold<R, S, R - S>(st, x0, x1, cut, nup, acc);
        if constexpr (S + 1 < R) EdgeStages<R, S + 1>::upper(st, d, c0, n, cut, nup, acc);
    }
};

// PHCO2, factorised chi classes: the same left-to-right fold as fold_run with gamma_eff = E(nu) F(line) per pair and one
// reciprocal per G lines (two lines per reciprocal cost 9.25 FP64 operations per evaluation).  q >= 9 (|dnu| >= 3 for every point of these classes), so a product of G = 8 of them is far from overflow.
template <int R, int G>
__device__ __forceinline__ void phco2_fold_run(const double4* __restrict__ st, const double* __restrict__ Fp, int& j, int x1,
                                               const double (&nup)[R], const double (&E)[R], double (&acc)[R])
{
    // gamma_eff^2 = E^2 F^2 and K gamma_eff = (K F) E: the per-point squares and the per-line products are formed once, which
    // leaves 7 operations per evaluation (dnu, E^2 F^2, q, (K F) E, three for the fold)
    double E2[R];
#pragma unroll
    for (int r = 0; r < R; r++) E2[r] = E[r] * E[r];
    for (; j + G - 1 < x1; j += G) {
        double fn[R], fd[R];
        {
            const double4 ra = st[j];
            const double fa = Fp[j];
            const double f2 = fa * fa, kf = ra.z * fa;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const double da = nup[r] - ra.x;
                fd[r] = fma(da, da, E2[r] * f2);
                fn[r] = kf * E[r];
            }
        }
#pragma unroll
        for (int g = 1; g < G; g++) {
            const double4 rb = st[j + g];
            const double fb = Fp[j + g];
            const double f2 = fb * fb, kf = rb.z * fb;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const double db = nup[r] - rb.x;
                const double qb = fma(db, db, E2[r] * f2);
                fn[r] = fma(kf * E[r], fd[r], fn[r] * qb);
                fd[r] *= qb;
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = fma(fn[r], cs_rcp(fd[r]), acc[r]);
    }
}

// PHCO2 lines that straddle a chi-class border for this tile: far-wing form with chi evaluated per point
template <int R>
__device__ __noinline__ void cold_phco2_generic(const WarpCold& w, const double4* st, int g0, int g1)
{
    const int lane = w.lane;
    double nup[R], acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) { nup[r] = w.nutile[32 * r + lane]; acc[r] = 0.0; }
    for (int j = g0; j < g1; j++) {
        double4 rc = st[j];
#pragma unroll
        for (int r = 0; r < R; r++) {
            double dnu = nup[r] - rc.x;
            double ge = chi_phco2(fabs(dnu), w.B1, w.B2) * rc.y;
            acc[r] = fma(rc.z * ge, cs_rcp(fma(dnu, dnu, ge * ge)), acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) w.cacc[32 * r + lane] += acc[r];
}

// ------------------------------------------------------------------------------------------------
// Far field through cluster moments (expansion mode, Voigt far wing and Lorentz).  The per-line expansion of K2 costs
// ~70 FP64 operations per eligible line per (tile, level); most of those lines are many tile widths away, where whole
// clusters of 32 consecutive lines can be translated at once:
//   P2M (moments_kernel, once per level):  K/((nu-nul)^2+g^2) = Im[s/(nu-z)], z = nul + i g, s = K/g, so about a real
//        cluster centre zc:  sum_j = sum_{m>=1} mu_m/(nu-zc)^(m+1),  mu_m = Im sum_j s_j (z_j-zc)^m = sum_j K_j C_m,
//        C_1 = 1, A_1 = u, C_{m+1} = u C_m + A_m, A_{m+1} = u A_m - g^2 C_m   (u = nul_j - zc)
//   M2L (farfield_kernel, per tile and level):  with R = zc - cen, t = (nu-cen)/h:
//        a_k -= (h/R)^k/R * S_k,  S_k = sum_m C(m+k,k) beta_m,  beta_m = mu_m (-1/R)^m, and S_k is element 0 of the
//        (k+1)-th suffix-sum pass over (0, beta_1, .., beta_p): no binomials, p^2/2 additions per cluster.
// A cluster is translated when |R| >= FF_THETA (h + rho), rho = its radius including the largest half width; with
// FF_THETA = 5 and FF_P = 18 terms the truncation is below 19 * 5^-18 = 5e-12 of the cluster's own contribution (measured:
// the same 1.2e-12 agreement with the direct sum on C2 as the per-line expansion).  Clusters that are too close,
// and the lines of the eligible ranges that do not fill a cluster, take the per-line expansion here as well, so K2 only
// has to evaluate the resulting polynomial (MP_P coefficients per (tile, level)).
constexpr int FF_P = 18;
constexpr double FF_THETA = 5.0;
constexpr int FF_REC = 20;      // doubles per (level, cluster): mu_1..mu_18, zc, rho

__global__ void __launch_bounds__(128) moments_kernel(const double4* __restrict__ rec, int64_t nl, int64_t ncl,
                                                      double* __restrict__ mom)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lev = blockIdx.y;
    if (c >= ncl) return;
    const double4* r = rec + (size_t)lev * nl + c * 32;
    const double x0 = r[0].x, x1 = r[31].x;
    const double zc = 0.5 * (x0 + x1);
    double mu[FF_P];
#pragma unroll
    for (int m = 0; m < FF_P; m++) mu[m] = 0.0;
    double g2max = 0.0;
    for (int j = 0; j < 32; j++) {
        const double4 rc = r[j];
        const double u = rc.x - zc, g2 = rc.y, K = rc.z;
        g2max = fmax(g2max, g2);
        double C = 1.0, A = u;
#pragma unroll
        for (int m = 0; m < FF_P; m++) {
            mu[m] = fma(K, C, mu[m]);
            const double Cn = fma(u, C, A);
            A = fma(u, A, -(g2 * C));
            C = Cn;
        }
    }
    double* o = mom + ((size_t)lev * ncl + c) * FF_REC;
#pragma unroll
    for (int m = 0; m < FF_P; m++) o[m] = mu[m];
    o[FF_P] = zc;
    const double ext = 0.5 * (x1 - x0);
    o[FF_P + 1] = sqrt(fma(ext, ext, g2max));
}

struct FarFieldArgs {
    const double* nu;
    int64_t nnu, nl, ncl, ntiles;
    const double4* rec;       // [nlev][nl]
    const double* mom;        // [nlev][ncl][FF_REC]
    const int64_t* ranges;    // [ntiles][8]
    double* ffc;              // [nlev][ntiles][MP_P]
    int tile_pts, lorentz;
};

__global__ void __launch_bounds__(128) farfield_kernel(FarFieldArgs a)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * 4 + warp;
    const int lev = blockIdx.y;
    if (tile >= a.ntiles) return;
    // the same line classes as line_sum_kernel (indices relative to the prefiltered list)
    const int64_t* rg = a.ranges + tile * 8;
    const int64_t wlo = rg[0], whi = rg[1];
    int64_t ilo = min(max(rg[2], wlo), whi), ihi = min(max(rg[3], wlo), whi);
    if (ilo >= ihi) { ilo = whi; ihi = whi; }
    ihi = max(ihi, ilo);
    int64_t mlo = min(max(rg[6], ilo), ihi), mhi = min(max(rg[7], mlo), ihi);
    int64_t nlo = a.lorentz ? mlo : min(max(rg[4], wlo), whi), nhi = a.lorentz ? mlo : min(max(rg[5], wlo), whi);
    nlo = min(max(nlo, ilo), ihi);
    nhi = min(max(nhi, nlo), ihi);
    mlo = min(mlo, nlo);
    mhi = max(mhi, nhi);
    const int64_t i0 = tile * a.tile_pts, i1 = min(i0 + (int64_t)a.tile_pts, a.nnu);
    const double tmin = a.nu[i0], tmax = a.nu[i1 - 1];
    const double cen = 0.5 * (tmin + tmax), h = 0.5 * (tmax - tmin);
    const double h2 = h * h, twoh = 2.0 * h;
    const double4* rec = a.rec + (size_t)lev * a.nl;
    const double* mom = a.mom + (size_t)lev * a.ncl * FF_REC;
    double am[MP_P];
#pragma unroll
    for (int k = 0; k < MP_P; k++) am[k] = 0.0;
    // per-line expansion of lines [j0,j1), one line per lane (same series as phase A of line_sum_kernel)
    auto p2l = [&](int64_t j0, int64_t j1) {
        for (int64_t j = j0 + lane; j < j1; j += 32) {
            const double4 rc = ld_rec(rec + j);
            const double u = rc.x - cen;
            const double r = cs_rcp(fma(u, u, rc.y));
            const double al = (u * r) * twoh, be = r * h2;
            double ck2 = rc.z * r;
            double ck1 = al * ck2;
            am[0] += ck2;
            am[1] += ck1;
#pragma unroll
            for (int k = 2; k < MP_P; k++) {
                double ck = fma(al, ck1, -(be * ck2));
                am[k] += ck;
                ck2 = ck1;
                ck1 = ck;
            }
        }
    };
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        const int64_t ja = side ? mhi : ilo, jb = side ? ihi : mlo;
        if (ja >= jb) continue;
        const int64_t cf = (ja + 31) >> 5, cl = min(jb >> 5, a.ncl);
        if (cf >= cl) { p2l(ja, jb); continue; }
        p2l(ja, cf << 5);
        p2l(cl << 5, jb);
        for (int64_t cb = cf; cb < cl; cb += 32) {
            const int64_t c = cb + lane;
            bool near_ = false;
            if (c < cl) {
                const double* mm = mom + (size_t)c * FF_REC;
                const double R = mm[FF_P] - cen, rho = mm[FF_P + 1];
                if (fabs(R) >= FF_THETA * (h + rho)) {
                    // beta_m = mu_m (-1/R)^m, then FF_P suffix-sum passes over the live triangle (m + k <= FF_P)
                    const double ir = 1.0 / R, mir = -ir;
                    double cc[FF_P + 1];
                    cc[0] = 0.0;
                    double pw = mir;
#pragma unroll
                    for (int m = 1; m <= FF_P; m++) { cc[m] = mm[m - 1] * pw; pw *= mir; }
                    double sc = -ir;              // -(h/R)^k / R
                    const double hr = h * ir;
#pragma unroll
                    for (int k = 0; k < FF_P; k++) {
#pragma unroll
                        for (int m = FF_P - k - 1; m >= 0; m--) cc[m] += cc[m + 1];
                        am[k] = fma(sc, cc[0], am[k]);
                        sc *= hr;
                    }
                } else {
                    near_ = true;
                }
            }
            // clusters that are too close for the translation: their 32 lines take the per-line expansion
            unsigned todo = __ballot_sync(0xffffffffu, near_);
            while (todo) {
                const int b = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t cc0 = (cb + b) << 5;
                p2l(cc0, cc0 + 32);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < MP_P; k++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) am[k] += __shfl_xor_sync(0xffffffffu, am[k], off);
    }
    if (lane == 0) {
        double* o = a.ffc + ((size_t)lev * a.ntiles + tile) * MP_P;
#pragma unroll
        for (int k = 0; k < MP_P; k++) o[k] = am[k];
    }
}

// K2.  Work unit = one warp = (tile of 32*R consecutive wavenumbers, level).  Each warp streams the records of
// ITS OWN line window through a private shared-memory ring fed by TMA bulk copies (one elected lane issues
// cp.async.bulk, completion on a per-stage mbarrier), so warps never wait for each other.  The 8 warps of a CTA
// take 8 adjacent tiles of the same level: their windows overlap by ~98 %, so the copies hit L2.
// MP = the far-field expansion is compiled in (a direct-mode launch takes the MP = false instantiation: none of the expansion's
// code, segment tables or registers)
// BAND = Voigt near-centre lines through the fold + per-point band correction (band_near) instead of cold_near per tile
// COLD = only the cold classes (cut-off edges, band / deferred evaluations): the lines inside the cut-off for every point are
// left to far_fold_kernel, launched after this one.  The partial sum goes to `out` in the form far_fold_kernel completes:
//   accumulate 0: out = scale*cold   1: out += scale*cold   2: out = cold (far_fold_kernel writes log(scale*(out + far)))
template <int SHAPE, int R, bool MP, bool BAND, bool COLD = false>
__global__ void __launch_bounds__(LS_THREADS, (COLD ? CS_LS_COLD_MINBLK : CS_LS_MINBLK) / LS_WARPS) line_sum_kernel(LineSumArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[LS_WARPS][LS_STAGES];
    __shared__ int seg_tab[LS_WARPS][3 * LS_NSEG];   // per warp: segment starts, ends, chunk counts (window-relative)
    __shared__ int edge_tab[LS_WARPS][2 * (R + 1)];  // COLD: stage borders of the cut-off edges (window-relative)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lev = blockIdx.y;
    constexpr int TILE = 32 * R;
    const int64_t tile = (int64_t)blockIdx.x * LS_WARPS + warp;
    if (tile >= a.ntiles) return;
    const int64_t tile0 = tile * TILE;
    const LevelParams lp = a.lev[lev];
    double4* ring = reinterpret_cast<double4*>(smem_raw) + (size_t)warp * LS_STAGES * LS_CHUNK;
    // per-warp extras behind the rings: copy of the tile's wavenumbers, cold-path accumulators, deferred queue
    constexpr size_t EXTRA = ls_extra_bytes<SHAPE, R>();
    unsigned char* xb = smem_raw + (size_t)LS_WARPS * LS_STAGES * LS_CHUNK * sizeof(double4) + (size_t)warp * EXTRA;
    WarpCold w;
    w.nutile = reinterpret_cast<double*>(xb);
    w.cacc = reinterpret_cast<double*>(xb) + TILE;
    w.queue = reinterpret_cast<uint32_t*>(w.cacc + TILE);
    w.slow_lev = a.slow ? a.slow + (size_t)lev * a.nl : nullptr;
    w.cut = a.cut; w.B1 = lp.B1; w.B2 = lp.B2; w.lane = lane;

    if (lane == 0) {
        for (int s = 0; s < LS_STAGES; s++) mbar_init(&full_bar[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    // all line indices below are 32-bit and relative to the first line of the tile's window (wlo64)
    const double near_cn = a.near_cn;
    const int64_t* rg = a.ranges + tile * a.nr;
    const int64_t wlo64 = rg[0];
    const int whi = (int)(rg[1] - wlo64);
    auto rel = [&](int k) { return (int)max(min(rg[k] - wlo64, (int64_t)whi), (int64_t)0); };
    int ilo = rel(2), ihi = rel(3), nlo = rel(4), nhi = rel(5);
    if (ilo >= ihi) { ilo = whi; ihi = whi; }        // cut-off window narrower than the tile: every line is an edge line
    ihi = max(ihi, ilo);
    // edge lines are normally ~cut-off away from every point, i.e. far wing; only when the near-centre range
    // reaches into the edge classes (tiny cut-offs, very coarse grids) do they need the per-evaluation region test
    w.edge_is_far = (SHAPE == CS_LORENTZ) ||
                    (SHAPE == CS_VOIGT && rg[4] - wlo64 >= ilo && rg[5] - wlo64 <= ihi && ilo < ihi);
    // far-field expansion (Voigt far wing and Lorentz only): lines [ilo,mlo) and [mhi,ihi) are summed through a local
    // Taylor expansion about the tile centre instead of point by point
    const bool mp = MP && (SHAPE == CS_VOIGT || SHAPE == CS_LORENTZ) && a.mp_theta > 0.0 && a.nr >= 8;
    int mlo = ilo, mhi = ihi;
    if (mp) { mlo = min(max(rel(6), ilo), ihi); mhi = min(max(rel(7), mlo), ihi); }
    if (SHAPE == CS_LORENTZ) { nlo = mlo; nhi = mlo; }   // no near-centre branch (empty range at the start of the direct lines)
    nlo = min(max(nlo, ilo), ihi);
    nhi = min(max(nhi, nlo), ihi);
    if (mp) { mlo = min(mlo, nlo); mhi = max(mhi, nhi); }   // never expand lines of the near-centre range
    // ---- PHCO2 chi-class borders (window-relative), needed before the chunk stream is laid out
    //   E | F4- | G | F3- | G | F2- | G | plain | near | plain | G | F2+ | G | F3+ | G | F4+ | E
    constexpr size_t EXTRA_ = ls_extra_bytes<SHAPE, R>();
    int* bnd = reinterpret_cast<int*>(xb + EXTRA_) - 36;      // last 18 int64 slots of the warp's extras hold 18 ints
    if (SHAPE == CS_PHCO2) {
        if (lane == 0) {
            int b[18];
            b[0] = 0; b[1] = ilo;
            for (int k = 0; k < 6; k++) b[2 + k] = min(max(rel(6 + k), ilo), nlo);    // never intrude into the near range
            b[8] = nlo; b[9] = nhi;
            for (int k = 0; k < 6; k++) b[10 + k] = max(min(rel(12 + k), ihi), nhi);
            b[16] = ihi; b[17] = whi;
            for (int k = 1; k < 18; k++) b[k] = max(b[k], b[k - 1]);
            for (int k = 0; k < 18; k++) bnd[k] = b[k];
        }
        __syncwarp();
    }
    // PHCO2 far-wing expansion: classes F3 and F4 (|dnu| >= 30 for every point of the tile) of a narrow enough tile at a
    // level where the chi*gamma correction series converges fast (flag set by the host)
    const double tile_h = 0.5 * (a.nu[min(tile0 + TILE, a.nnu) - 1] - a.nu[tile0]);
    const bool px = MP && SHAPE == CS_PHCO2 && a.mp_theta > 0.0 && lp.pexp_ok > 0.0 && tile_h <= PX_HMAX && a.chix != nullptr;
    // segments streamed through the ring, in this order; a chunk never spans two segments:
    //   Voigt/Lorentz with the expansion: expanded [ilo,mlo) [mhi,ihi) (phase A), then direct [0,ilo) [mlo,mhi) [ihi,whi)
    //   PHCO2 with the expansion:         the five direct ranges between and around F4-, F3-, F3+, F4+; the expanded lines
    //                                     are read straight from global memory, one line per lane (their 1000+ chunks per
    //                                     tile would cost more in chunk prologues than the ring saves)
    //   no expansion:                     [0,whi)
    // (kept in shared memory: indexed dynamically by the chunk number, warp-uniform)
    const bool multi = mp || px || COLD;
    int* slo = seg_tab[warp];
    int* shi = seg_tab[warp] + LS_NSEG;
    int* sch = seg_tab[warp] + 2 * LS_NSEG;
    if (lane == 0) {
        for (int q = 0; q < LS_NSEG; q++) { slo[q] = 0; shi[q] = 0; }
        if (mp) {
            // expanded lines (phase A) are streamed first -- unless farfield_kernel already summed them
            slo[0] = ilo; shi[0] = a.ffc ? ilo : mlo;
            slo[1] = mhi; shi[1] = a.ffc ? mhi : ihi;
            slo[2] = 0;   shi[2] = ilo;
            slo[3] = mlo; shi[3] = mhi;
            slo[4] = ihi; shi[4] = whi;
        } else if (px) {
            slo[0] = 0;       shi[0] = bnd[1];
            slo[1] = bnd[2];  shi[1] = bnd[3];
            slo[2] = bnd[4];  shi[2] = bnd[13];
            slo[3] = bnd[14]; shi[3] = bnd[15];
            slo[4] = bnd[16]; shi[4] = whi;
        } else if (COLD) {
            slo[0] = 0;   shi[0] = ilo;
            slo[1] = ihi; shi[1] = whi;
            if (CS_LS_EDGE_STAGES && a.nr == LS_NR_SPLIT) {
                // lower stages [b_s, b_{s+1}) with b_0 = 0, b_R = ilo; upper stages [d_s, d_{s+1}) with d_0 = ihi, d_R = whi
                int* eb = edge_tab[warp];
                int* ed = edge_tab[warp] + (R + 1);
                eb[0] = 0; eb[R] = ilo; ed[0] = ihi; ed[R] = whi;
                for (int r = 1; r < R; r++) { eb[r] = min(rel(5 + r), ilo); ed[r] = max(rel(6 + (R - 1) + (r - 1)), ihi); }
                for (int r = 1; r <= R; r++) { eb[r] = max(eb[r], eb[r - 1]); ed[r] = max(ed[r], ed[r - 1]); }
            }
        } else {
            shi[0] = whi;
        }
        for (int q = 0; q < LS_NSEG; q++) sch[q] = (shi[q] - slo[q] + LS_CHUNK - 1) / LS_CHUNK;
    }
    __syncwarp();
    int nchunk = 0;
#pragma unroll
    for (int q = 0; q < LS_NSEG; q++) nchunk += sch[q];
    const int nchunkA = mp ? sch[0] + sch[1] : 0;
    const double4* rec_lev = a.rec + (size_t)lev * a.nl + wlo64;          // records of the window
    if (w.slow_lev) w.slow_lev += wlo64;
    // chunk number -> line range [c0,c1)
    auto chunk_bounds = [&](int c, int& c0, int& c1) {
        if (!multi) {            // one segment
            c0 = c * LS_CHUNK;
            c1 = min(c0 + LS_CHUNK, whi);
        } else {
            int q = 0;
            while (q < LS_NSEG - 1 && c >= sch[q]) { c -= sch[q]; q++; }
            c0 = slo[q] + c * LS_CHUNK;
            c1 = min(c0 + LS_CHUNK, shi[q]);
        }
    };

    auto issue = [&](int c) {   // lane 0 only
        int s = c % LS_STAGES;
        int c0, c1;
        chunk_bounds(c, c0, c1);
        uint32_t bytes = (uint32_t)(c1 - c0) * (uint32_t)sizeof(double4);
        mbar_arrive_expect_tx(&full_bar[warp][s], bytes);
        tma_bulk_g2s(ring + (size_t)s * LS_CHUNK, rec_lev + c0, bytes, &full_bar[warp][s]);
    };
    if (lane == 0)
        for (int c = 0; c < LS_STAGES && c < nchunk; c++) issue(c);

    double nup[R], acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = tile0 + 32 * r + lane;
        nup[r] = a.nu[i < a.nnu ? i : (a.nnu - 1)];
        acc[r] = 0.0;
        const_cast<double*>(w.nutile)[32 * r + lane] = nup[r];
        w.cacc[32 * r + lane] = 0.0;
    }
    __syncwarp();
    int qn = 0;   // entries in the deferred queue (warp-uniform)
    // ---- PHCO2: chi(|dnu|) = exp(-c0 - B (|dnu| - a)) factorises, for a line whose chi class is the same for all
    // points of the tile, into a per-point factor E = exp(-+B (nu - nu0)) and a per-line factor
    // F = exp(-c0 - B (+-(nu0 - nul) - a)): one multiplication per evaluation instead of one exp.
    double* Etab = reinterpret_cast<double*>(w.queue + LS_QCAP);
    double* Fp = Etab + 6 * TILE;
    const double nu0 = w.nutile[0];
    if (SHAPE == CS_PHCO2) {
        const double Bc[3] = {lp.B1, lp.B2, 0.0232};
#pragma unroll
        for (int r = 0; r < R; r++) {
            double dx = nup[r] - nu0;
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
                Etab[(2 * cc + 0) * TILE + 32 * r + lane] = exp(-Bc[cc] * dx);   // line below the point
                Etab[(2 * cc + 1) * TILE + 32 * r + lane] = exp(Bc[cc] * dx);    // line above the point
            }
        }
        __syncwarp();
    }

    // ---- phase A: far-field expansion.  For a line at u = nul - cen (|u| >= theta*h) and t = (nu - cen)/h in [-1,1]
    //   K/((h t - u)^2 + g^2) = sum_k c_k t^k,  c_0 = K/q0,  c_1 = al c_0,  c_k = al c_{k-1} - be c_{k-2},
    //   q0 = u^2 + g^2, al = 2 u h/q0, be = h^2/q0   (generating function of the Chebyshev polynomials U_k);
    // |c_k| <= (k+1) theta^-k c_0, so MP_P = 20 terms at theta = 4 truncate below 3e-11 of each line's own value
    // (all terms of the sum are positive, so that also bounds the relative error of the sum).  A lane sums the
    // coefficients of every 32nd line (3 FP64 ops per coefficient), a butterfly reduces them over the warp, and every
    // point then costs MP_P FMAs (Horner) instead of 5.25 FP64 ops per line.
    if (mp && a.ffc) {
        // far field precomputed per (tile, level) by farfield_kernel: evaluate its polynomial at the lane's points
        const double cen = 0.5 * (w.nutile[0] + a.nu[min(tile0 + TILE, a.nnu) - 1]);
        const double h = 0.5 * (a.nu[min(tile0 + TILE, a.nnu) - 1] - w.nutile[0]);
        const double ih = h > 0.0 ? 1.0 / h : 0.0;
        const double* fc = a.ffc + ((size_t)lev * a.ntiles + tile) * MP_P;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const double t = (nup[r] - cen) * ih;
            double v = fc[MP_P - 1];
#pragma unroll
            for (int k = MP_P - 2; k >= 0; k--) v = fma(v, t, fc[k]);
            acc[r] = v;
        }
    }
    if (mp && nchunkA > 0) {
        const double cen = 0.5 * (w.nutile[0] + a.nu[min(tile0 + TILE, a.nnu) - 1]);
        const double h = 0.5 * (a.nu[min(tile0 + TILE, a.nnu) - 1] - w.nutile[0]);
        const double h2 = h * h, twoh = 2.0 * h;
        double am[MP_P];
#pragma unroll
        for (int k = 0; k < MP_P; k++) am[k] = 0.0;
        for (int c = 0; c < nchunkA; c++) {
            const int s = c % LS_STAGES;
            const uint32_t ph = (c / LS_STAGES) & 1;
            int c0, c1;
            chunk_bounds(c, c0, c1);
            mbar_wait(&full_bar[warp][s], ph);
            const double4* st = ring + (size_t)s * LS_CHUNK;
            const int n = c1 - c0;
            for (int jj = lane; jj < n; jj += 32) {
                const double4 rc = st[jj];
                const double u = rc.x - cen;
                const double r = cs_rcp(fma(u, u, rc.y));
                const double al = (u * r) * twoh, be = r * h2;
                double ck2 = rc.z * r;
                double ck1 = al * ck2;
                am[0] += ck2;
                am[1] += ck1;
#pragma unroll
                for (int k = 2; k < MP_P; k++) {
                    double ck = fma(al, ck1, -(be * ck2));
                    am[k] += ck;
                    ck2 = ck1;
                    ck1 = ck;
                }
            }
            __syncwarp();
            if (lane == 0 && c + LS_STAGES < nchunk) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(c + LS_STAGES);
            }
        }
#pragma unroll
        for (int k = 0; k < MP_P; k++) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) am[k] += __shfl_xor_sync(0xffffffffu, am[k], off);
        }
        const double ih = h > 0.0 ? 1.0 / h : 0.0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const double t = (nup[r] - cen) * ih;
            double v = am[MP_P - 1];
#pragma unroll
            for (int k = MP_P - 2; k >= 0; k--) v = fma(v, t, am[k]);
            acc[r] = v;
        }
    }

    // ---- phase A, PHCO2: far wings of class F3 / F4 lines (|dnu| >= 30 cm^-1 for every point of the tile).  With
    // ge = chi*gamma = E(nu) Fg (per-point factor E, per-line factor Fg as in the direct path) and eps = (ge/dnu)^2,
    //   K ge/(dnu^2 + ge^2) = K ge/dnu^2 (1 - eps + eps^2 - ...) = E [K Fg/dnu^2] - E^3 [K Fg^3/dnu^4] + E^5 [K Fg^5/dnu^6] - ...
    // (eps < 1e-4 at the levels the host enables, so the first neglected term is below 1e-12), and each bracket, summed over
    // the lines, is a power law expanded about the tile centre: with u = nul - cen, r = h/u, t = (nu - cen)/h,
    //   1/(h t - u)^(2m) = u^(-2m) sum_k C(k+2m-1, k) r^k t^k,   |r| <= 1/16 because |u| >= 30 + h and h <= 2,
    // so a lane only accumulates the power sums G_m[k] = sum_j w_j r_j^k (2 FP64 ops per term); 10 / 6 / 4 terms keep the
    // truncation below 1e-11 of the line's own value.  One exp per line (its chi factor) instead of one per evaluation
    // or one multiplication per evaluation in the factorised direct path.
    if (px) {
        const double cen = 0.5 * (w.nutile[0] + a.nu[min(tile0 + TILE, a.nnu) - 1]);
        const double h = tile_h;
        const double ih = h > 0.0 ? 1.0 / h : 0.0;
        constexpr int NG = PX_P2 + PX_P4 + PX_P6;
        double G[NG];
        auto flush = [&](int q) {
            // q: 0 F4-, 1 F3-, 2 F3+, 3 F4+  ->  chi class 2,1,1,2 and side below/below/above/above
            const int tab = q == 0 ? 4 : (q == 1 ? 2 : (q == 2 ? 3 : 5));
#pragma unroll
            for (int k = 0; k < NG; k++) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) G[k] += __shfl_xor_sync(0xffffffffu, G[k], off);
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                const double t = (nup[r] - cen) * ih;
                const double E = Etab[tab * TILE + 32 * r + lane], E2 = E * E;
                double p2 = (double)PX_P2 * G[PX_P2 - 1];
#pragma unroll
                for (int k = PX_P2 - 2; k >= 0; k--) p2 = fma(p2, t, (double)(k + 1) * G[k]);
                double p4 = (double)((PX_P4 + 2) * (PX_P4 + 1) * PX_P4 / 6) * G[PX_P2 + PX_P4 - 1];
#pragma unroll
                for (int k = PX_P4 - 2; k >= 0; k--) p4 = fma(p4, t, (double)((k + 3) * (k + 2) * (k + 1) / 6) * G[PX_P2 + k]);
                double p6 = (double)((PX_P6 + 4) * (PX_P6 + 3) * (PX_P6 + 2) * (PX_P6 + 1) * PX_P6 / 120) * G[NG - 1];
#pragma unroll
                for (int k = PX_P6 - 2; k >= 0; k--)
                    p6 = fma(p6, t, (double)((k + 5) * (k + 4) * (k + 3) * (k + 2) * (k + 1) / 120) * G[PX_P2 + PX_P4 + k]);
                acc[r] = fma(E, fma(-E2, fma(-E2, p6, p4), p2), acc[r]);
            }
        };
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            const int lo = bnd[q == 0 ? 1 : (q == 1 ? 3 : (q == 2 ? 13 : 15))];
            const int hi = bnd[q == 0 ? 2 : (q == 1 ? 4 : (q == 2 ? 14 : 16))];
            if (lo >= hi) continue;
#pragma unroll
            for (int k = 0; k < NG; k++) G[k] = 0.0;
            const bool up = q >= 2;
            const bool cls2 = (q == 0) || (q == 3);
            if (cls2) {
                // chi line factor = Cq * X^(-+1): Cq carries everything that does not depend on the line
                const double c0c = lp.B1 * 27.0 + lp.B2 * 90.0;
                const double Cq = up ? exp(-c0c + 0.0232 * (120.0 + (nu0 - a.chix_ref)))
                                     : exp(-c0c - 0.0232 * ((nu0 - a.chix_ref) - 120.0));
                const double2* cx = a.chix + wlo64;
                for (int j = lo + lane; j < hi; j += 32) {
                    const double4 rc = ld_rec(rec_lev + j);
                    const double2 xx = __ldg(cx + j);
                    const double Fg = (Cq * (up ? xx.y : xx.x)) * rc.y;         // chi's line factor times gamma
                    const double u = rc.x - cen;
                    const double iu = copysign(cs_rcp(fabs(u)), u);
                    const double r = h * iu, iu2 = iu * iu, Fg2 = Fg * Fg;
                    double w2 = (rc.z * Fg) * iu2;
                    double w4 = (w2 * Fg2) * iu2;
                    double w6 = (w4 * Fg2) * iu2;
#pragma unroll
                    for (int k = 0; k < PX_Q2; k++) { G[k] += w2; w2 *= r; }
#pragma unroll
                    for (int k = 0; k < PX_Q4; k++) { G[PX_P2 + k] += w4; w4 *= r; }
#pragma unroll
                    for (int k = 0; k < PX_Q6; k++) { G[PX_P2 + PX_P4 + k] += w6; w6 *= r; }
                }
            } else {
                const double B = lp.B2, c0c = lp.B1 * 27.0, aa = 30.0;
                for (int j = lo + lane; j < hi; j += 32) {
                    const double4 rc = ld_rec(rec_lev + j);
                    const double d = up ? (rc.x - nu0) - aa : (nu0 - rc.x) - aa;
                    const double Fg = exp(-c0c - B * d) * rc.y;                  // chi's line factor times gamma
                    const double u = rc.x - cen;
                    const double iu = copysign(cs_rcp(fabs(u)), u);
                    const double r = h * iu, iu2 = iu * iu, Fg2 = Fg * Fg;
                    double w2 = (rc.z * Fg) * iu2;
                    double w4 = (w2 * Fg2) * iu2;
                    double w6 = (w4 * Fg2) * iu2;
#pragma unroll
                    for (int k = 0; k < PX_P2; k++) { G[k] += w2; w2 *= r; }
#pragma unroll
                    for (int k = 0; k < PX_P4; k++) { G[PX_P2 + k] += w4; w4 *= r; }
#pragma unroll
                    for (int k = 0; k < PX_P6; k++) { G[PX_P2 + PX_P4 + k] += w6; w6 *= r; }
                }
            }
            flush(q);
        }
    }

    if (BAND && nlo < nhi)
        band_near<R>(w, rec_lev + nlo, w.slow_lev + nlo, a.nul + wlo64 + nlo, nhi - nlo, lp.cnear);

    int sgc = 0;   // PHCO2: cursor into the 17 chi-class segments
    for (int c = nchunkA; c < nchunk; c++) {
        const int s = c % LS_STAGES;
        const uint32_t ph = (c / LS_STAGES) & 1;
        int c0, c1;
        chunk_bounds(c, c0, c1);
        mbar_wait(&full_bar[warp][s], ph);
        const double4* st = ring + (size_t)s * LS_CHUNK;
        // chunk-local boundaries of the five classes: [0,xa) edge | [xa,xb) far | [xb,xc) near | [xc,xd) far | [xd,n) edge
        const int n = c1 - c0;
        const int xa = min(max(ilo - c0, 0), n), xd = min(max(ihi - c0, 0), n);
        int xb_ = min(max(nlo - c0, 0), n), xc = min(max(nhi - c0, 0), n);
        if (BAND) { xb_ = xd; xc = xd; }      // the near range is folded like any far range; band_near corrected it per point
        if (!BAND && SHAPE == CS_VOIGT && CS_LS_NEAR_TRIM && xb_ < xc && lp.cnear < near_cn) {
            // the per-tile near range was sized with the widest Doppler zone of the level batch (its warmest level); at THIS
            // level the lines at both ends of it are still safely in the far wing: hand them back to the far ranges
            const double tlo = w.nutile[0], thi = w.nutile[TILE - 1];
            const double cp = 1.0 + lp.cnear, cm = 1.0 - lp.cnear;
            int nb = 0, na = 0;
            for (int base = xb_; base < xc; base += 32) {
                const int j = base + lane;
                const bool ok = j < xc;
                const double x = st[ok ? j : xc - 1].x;
                nb += __popc(__ballot_sync(0xffffffffu, ok && x * cp < tlo));
                na += __popc(__ballot_sync(0xffffffffu, ok && x * cm > thi));
            }
            xb_ += nb;
            xc = max(xb_, xc - na);
        }
        if (SHAPE == CS_PHCO2) {
            // per-line chi factors of this chunk (lines in one of the six factorised segments)
            for (int jj = lane; jj < n; jj += 32) {
                const int jg = c0 + jj;
                int cls = -1, up = 0;
                if (jg >= bnd[1] && jg < bnd[2]) cls = 2;
                else if (jg >= bnd[3] && jg < bnd[4]) cls = 1;
                else if (jg >= bnd[5] && jg < bnd[6]) cls = 0;
                else if (jg >= bnd[11] && jg < bnd[12]) { cls = 0; up = 1; }
                else if (jg >= bnd[13] && jg < bnd[14]) { cls = 1; up = 1; }
                else if (jg >= bnd[15] && jg < bnd[16]) { cls = 2; up = 1; }
                if (cls >= 0) {
                    double4 rc = st[jj];
                    double B = cls == 0 ? lp.B1 : (cls == 1 ? lp.B2 : 0.0232);
                    double c0c = cls == 0 ? 0.0 : (cls == 1 ? lp.B1 * 27.0 : lp.B1 * 27.0 + lp.B2 * 90.0);
                    double aa = cls == 0 ? 3.0 : (cls == 1 ? 30.0 : 120.0);
                    double d = up ? (rc.x - nu0) - aa : (nu0 - rc.x) - aa;
                    Fp[jj] = exp(-c0c - B * d) * rc.y;      // chi's line factor times gamma
                }
            }
            __syncwarp();
            // chunks arrive in ascending line order: a cursor skips the segments that end before this chunk, and the loop
            // stops at the first segment that starts after it (1-2 iterations instead of 17)
            while (sgc < 16 && bnd[sgc + 1] <= c0) sgc++;
#pragma unroll 1
            for (int sgm = sgc; sgm < 17 && bnd[sgm] < c1; sgm++) {
                const int x0 = min(max(bnd[sgm] - c0, 0), n), x1 = min(max(bnd[sgm + 1] - c0, 0), n);
                if (x0 >= x1) continue;
                if (sgm == 0 || sgm == 16) { cold_edge<SHAPE, R>(w, st, c0, x0, x1); continue; }
                if (sgm == 8) { qn = cold_near<SHAPE, R>(w, st, c0, x0, x1, qn); continue; }
                if (sgm == 7 || sgm == 9) {
                    // |dnu| < 3 for every point: chi = 1, plain far-wing Voigt with gamma from the record
                    int j = x0;
                    for (; j + 1 < x1; j += 2) {
                        double4 ra = st[j], rb = st[j + 1];
                        double ga2 = ra.y * ra.y, gb2 = rb.y * rb.y, ka = ra.z * ra.y, kb = rb.z * rb.y;
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            double da = nup[r] - ra.x, db = nup[r] - rb.x;
                            double qa = fma(da, da, ga2), qb = fma(db, db, gb2);
                            acc[r] = fma(fma(kb, qa, ka * qb), cs_rcp(qa * qb), acc[r]);
                        }
                    }
                    for (; j < x1; j++) {
                        double4 rc = st[j];
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            double dnu = nup[r] - rc.x;
                            acc[r] = fma(rc.z * rc.y, cs_rcp(fma(dnu, dnu, rc.y * rc.y)), acc[r]);
                        }
                    }
                    continue;
                }
                if ((sgm & 1) == 0) { cold_phco2_generic<R>(w, st, x0, x1); continue; }   // straddles a chi border
                // factorised chi: sgm 1,3,5 = classes 2,1,0 below; 11,13,15 = classes 0,1,2 above
                const int tab = sgm < 8 ? 2 * ((5 - sgm) / 2) : 2 * ((sgm - 11) / 2) + 1;
                double E[R];
#pragma unroll
                for (int r = 0; r < R; r++) E[r] = Etab[tab * TILE + 32 * r + lane];
                int j = x0;
#if CS_LS_PHCO2_FOLD
                phco2_fold_run<R, CS_LS_PHCO2_FOLD>(st, Fp, j, x1, nup, E, acc);
#endif
                for (; j + 1 < x1; j += 2) {
                    double4 ra = st[j], rb = st[j + 1];
                    double fa = Fp[j], fb = Fp[j + 1];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        double da = nup[r] - ra.x, db = nup[r] - rb.x;
                        double gea = E[r] * fa, geb = E[r] * fb;
                        double qa = fma(da, da, gea * gea), qb = fma(db, db, geb * geb);
                        double num = fma(rb.z * geb, qa, (ra.z * gea) * qb);
                        acc[r] = fma(num, cs_rcp(qa * qb), acc[r]);
                    }
                }
                for (; j < x1; j++) {
                    double4 rc = st[j];
                    double f = Fp[j];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        double dnu = nup[r] - rc.x;
                        double ge = E[r] * f;
                        acc[r] = fma(rc.z * ge, cs_rcp(fma(dnu, dnu, ge * ge)), acc[r]);
                    }
                }
            }
        } else {
        const bool staged = COLD && CS_LS_EDGE_STAGES && a.nr == LS_NR_SPLIT && w.edge_is_far;
        if (xa > 0) {
            if (staged) EdgeStages<R, 0>::lower(st, edge_tab[warp], c0, n, w.cut, nup, acc);
            else cold_edge<SHAPE, R>(w, st, c0, 0, xa);
        }
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            // far lines: inside the cut-off for every point and safely in the far wing -> no test of any kind
            int j = pass ? xc : xa;
            const int f1 = pass ? xd : xb_;
            if (SHAPE != CS_DOPPLER) {      // Doppler: exp(-(dnu/alpha)^2) underflows to exactly 0 out here
                if (SHAPE == CS_VOIGT || SHAPE == CS_LORENTZ) {
                    // direct mode (long ranges: whole chunks) in groups of CS_LS_FOLD lines per reciprocal, the short direct
                    // remainder of the expansion mode in groups of 8: a group is one dependent chain per point, and a short
                    // range has nothing else to overlap it with (measured on C2: 16 / 8 is the best pair; the 2 x 2 tree of
                    // round 1 -- four lines per reciprocal, 5.25 FP64 ops per evaluation -- was 5 % slower)
                    fold_run<R, (MP ? 8 : CS_LS_FOLD)>(st, j, f1, nup, acc);
                    fold_tail<R>(st, j, f1, nup, acc);
                }
                for (; j < f1; j++) {
                    double4 rc = st[j];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        double dnu = nup[r] - rc.x;
                        if (SHAPE == CS_PHCO2) {
                            double ge = chi_phco2(fabs(dnu), w.B1, w.B2) * rc.y;
                            acc[r] = fma(rc.z * ge, cs_rcp(fma(dnu, dnu, ge * ge)), acc[r]);
                        } else {
                            acc[r] = fma(rc.z, cs_rcp(fma(dnu, dnu, rc.y)), acc[r]);
                        }
                    }
                }
            }
            if (pass == 0 && xb_ < xc) qn = cold_near<SHAPE, R>(w, st, c0, xb_, xc, qn);
        }
        if (xd < n) {
            if (staged) EdgeStages<R, 0>::upper(st, edge_tab[warp] + (R + 1), c0, n, w.cut, nup, acc);
            else cold_edge<SHAPE, R>(w, st, c0, xd, n);
        }
        if (SHAPE == CS_VOIGT && qn > 0) { cold_flush<SHAPE, R>(w, st, c0, qn); qn = 0; }
        }
        // stage s is free again: refill it with chunk c + LS_STAGES
        __syncwarp();
        if (lane == 0 && c + LS_STAGES < nchunk) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(c + LS_STAGES);
        }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = tile0 + 32 * r + lane;
        if (i < a.nnu) {
            size_t o = (size_t)lev * a.nnu + i;
            if (COLD) {
                const double c = acc[r] + w.cacc[32 * r + lane];
                a.out[o] = a.accumulate == 1 ? a.out[o] + lp.scale * c : (a.accumulate == 2 ? c : lp.scale * c);
            } else {
                double v = lp.scale * (acc[r] + w.cacc[32 * r + lane]);
                a.out[o] = a.accumulate == 1 ? a.out[o] + v : (a.accumulate == 2 ? log(v) : v);
            }
        }
    }
}

// K2, far wings only (direct mode, Voigt with the band correction and Lorentz): the lines inside the cut-off for EVERY point of
// the tile -- 95 % of the window on C2 -- need no test of any kind, so they get a kernel that is nothing but the fold: same work
// unit (one warp = tile x level, private TMA ring), a fraction of the registers and of the instruction footprint of
// line_sum_kernel, hence more resident warps to keep the FP64 pipe fed.  Runs after line_sum_kernel<.., COLD> and completes `out`.
constexpr int HF_STAGES = 2;
template <int R, int G>
__global__ void __launch_bounds__(32, CS_LS_HOT_MINBLK) far_fold_kernel(LineSumArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[HF_STAGES];
    constexpr int TILE = 32 * R;
    const int lane = threadIdx.x;
    const int lev = blockIdx.y;
    const int64_t tile = blockIdx.x;
    const int64_t tile0 = tile * TILE;
    double4* ring = reinterpret_cast<double4*>(smem_raw);
    if (lane == 0) {
        for (int s = 0; s < HF_STAGES; s++) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    // the same all-inside range as line_sum_kernel computes (window-relative, clamped)
    const int64_t* rg = a.ranges + tile * a.nr;
    const int64_t wlo64 = rg[0];
    const int whi = (int)(rg[1] - wlo64);
    int ilo = (int)max(min(rg[2] - wlo64, (int64_t)whi), (int64_t)0), ihi = (int)max(min(rg[3] - wlo64, (int64_t)whi), (int64_t)0);
    if (ilo >= ihi) { ilo = whi; ihi = whi; }
    const int nfar = ihi - ilo;
    const int nchunk = (nfar + LS_CHUNK - 1) / LS_CHUNK;
    const double4* rec_far = a.rec + (size_t)lev * a.nl + wlo64 + ilo;
    auto issue = [&](int c) {   // lane 0 only
        const int s = c % HF_STAGES;
        const int c0 = c * LS_CHUNK;
        const uint32_t bytes = (uint32_t)(min(LS_CHUNK, nfar - c0)) * (uint32_t)sizeof(double4);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        tma_bulk_g2s(ring + (size_t)s * LS_CHUNK, rec_far + c0, bytes, &full_bar[s]);
    };
    if (lane == 0)
        for (int c = 0; c < HF_STAGES && c < nchunk; c++) issue(c);
    double nup[R], acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t i = tile0 + 32 * r + lane;
        nup[r] = a.nu[i < a.nnu ? i : (a.nnu - 1)];
        acc[r] = 0.0;
    }
    for (int c = 0; c < nchunk; c++) {
        const int s = c % HF_STAGES;
        mbar_wait(&full_bar[s], (uint32_t)((c / HF_STAGES) & 1));
        const double4* st = ring + (size_t)s * LS_CHUNK;
        const int n = min(LS_CHUNK, nfar - c * LS_CHUNK);
        int j = 0;
        fold_run<R, G>(st, j, n, nup, acc);
        fold_tail<R>(st, j, n, nup, acc);
        __syncwarp();
        if (lane == 0 && c + HF_STAGES < nchunk) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(c + HF_STAGES);
        }
    }
    const double scale = a.lev[lev].scale;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t i = tile0 + 32 * r + lane;
        if (i < a.nnu) {
            const size_t o = (size_t)lev * a.nnu + i;
            const double prev = a.out[o];
            a.out[o] = a.accumulate == 2 ? log(scale * (prev + acc[r])) : fma(scale, acc[r], prev);
        }
    }
}

template <int SHAPE, int R> int32_t launch_line_sum(cs_ctx* ctx, LineSumArgs a, int nlev, double cn)
{
    constexpr int TILE = 32 * R;
    cudaStream_t st = ctx->stream;
    a.ntiles = (a.nnu + TILE - 1) / TILE;
    a.near_cn = cn;
    if (SHAPE == CS_DOPPLER) a.mp_theta = 0.0;
    constexpr bool CAN_BAND = CS_LS_BAND && SHAPE == CS_VOIGT;
    // the band correction pays in direct mode (the near lines ride the long far-wing fold); in expansion mode the direct fold is
    // short and the per-tile cold_near measured faster (26.6 against 27.9 ms on C2)
    const bool band = CAN_BAND && a.band && !(a.mp_theta > 0.0);
    constexpr bool CAN_SPLIT = CS_LS_SPLIT && LS_WARPS == 1 && ((SHAPE == CS_VOIGT && CAN_BAND) || SHAPE == CS_LORENTZ);
    const bool split = CAN_SPLIT && a.split && !(a.mp_theta > 0.0) && (band || SHAPE == CS_LORENTZ);
    a.nr = (SHAPE == CS_PHCO2) ? LS_NR : (a.mp_theta > 0.0 ? 8 : ((split && R == 4) ? LS_NR_SPLIT : 6));
    // scratch: the per-tile ranges, then (PHCO2 expansion) the per-line chi factors of the >= 120 cm^-1 class, which are
    // level-independent (skipped when the span of the line list would overflow exp: the expansion is then off and every
    // pair is summed directly), or (Voigt / Lorentz expansion) the cluster moments and the far-field coefficients
    auto al256 = [](size_t b) { return ((b + 255) / 256) * 256; };
    const size_t off = al256(sizeof(int64_t) * a.nr * (size_t)a.ntiles);
    const bool want_chix = SHAPE == CS_PHCO2 && a.mp_theta > 0.0 && a.nl > 0 && 0.0232 * 0.5 * (a.nul_hi - a.nul_lo) < 600.0;
    const bool want_ff = (SHAPE == CS_VOIGT || SHAPE == CS_LORENTZ) && a.mp_theta > 0.0 && a.nl > 0 && !ctx->ff_no_moments;
    const int64_t ncl = a.nl / 32;
    const size_t off_ffc = off + al256(sizeof(double) * FF_REC * (size_t)ncl * nlev);
    const size_t total = want_chix ? off + sizeof(double2) * (size_t)a.nl
                                   : (want_ff ? off_ffc + sizeof(double) * MP_P * (size_t)a.ntiles * nlev : off);
    CS_TRY(ctx->s_w.reserve(total));
    a.ranges = ctx->s_w.as<int64_t>();
    tile_ranges_kernel<<<(unsigned)((a.ntiles * a.nr + 127) / 128), 128, 0, st>>>(a.nu, a.nnu, a.nul, a.nl, a.cut, cn, TILE,
                                                                                  a.ntiles, a.nr, a.mp_theta,
                                                                                  ctx->s_w.as<int64_t>());
    CS_CUDA(cudaGetLastError());
    a.chix = nullptr;
    a.chix_ref = 0.0;
    a.ffc = nullptr;
    if (want_chix) {
        double2* cx = reinterpret_cast<double2*>(ctx->s_w.as<char>() + off);
        a.chix_ref = 0.5 * (a.nul_lo + a.nul_hi);
        chix_kernel<<<(unsigned)((a.nl + 255) / 256), 256, 0, st>>>(a.nul, a.nl, a.chix_ref, cx);
        CS_CUDA(cudaGetLastError());
        a.chix = cx;
    }
    if (want_ff) {
        FarFieldArgs fa;
        fa.nu = a.nu; fa.nnu = a.nnu; fa.nl = a.nl; fa.ncl = ncl; fa.ntiles = a.ntiles;
        fa.rec = a.rec;
        fa.mom = reinterpret_cast<const double*>(ctx->s_w.as<char>() + off);
        fa.ranges = a.ranges;
        fa.ffc = reinterpret_cast<double*>(ctx->s_w.as<char>() + off_ffc);
        fa.tile_pts = TILE;
        fa.lorentz = SHAPE == CS_LORENTZ;
        if (ncl > 0) {
            moments_kernel<<<dim3((unsigned)((ncl + 127) / 128), (unsigned)nlev), 128, 0, st>>>(
                a.rec, a.nl, ncl, reinterpret_cast<double*>(ctx->s_w.as<char>() + off));
            CS_CUDA(cudaGetLastError());
        }
        farfield_kernel<<<dim3((unsigned)((a.ntiles + 3) / 4), (unsigned)nlev), 128, 0, st>>>(fa);
        CS_CUDA(cudaGetLastError());
        cs_count_launch(ctx, 2);
        a.ffc = fa.ffc;
    }
    size_t smem = (size_t)LS_WARPS * (LS_STAGES * LS_CHUNK * sizeof(double4) + ls_extra_bytes<SHAPE, R>());
    if (smem > 48 * 1024)   // per device: not cached, several contexts may live on different GPUs
    {
        CS_CUDA(cudaFuncSetAttribute(line_sum_kernel<SHAPE, R, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CS_CUDA(cudaFuncSetAttribute(line_sum_kernel<SHAPE, R, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (CAN_BAND) {
            CS_CUDA(cudaFuncSetAttribute(line_sum_kernel<SHAPE, R, false, CAN_BAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    }
    dim3 grid((unsigned)((a.ntiles + LS_WARPS - 1) / LS_WARPS), (unsigned)nlev);
    if (split) {
        constexpr bool B = SHAPE == CS_VOIGT;
        if (smem > 48 * 1024)
            CS_CUDA(cudaFuncSetAttribute(line_sum_kernel<SHAPE, R, false, B, CAN_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        line_sum_kernel<SHAPE, R, false, B, CAN_SPLIT><<<grid, LS_THREADS, smem, st>>>(a);
        CS_CUDA(cudaGetLastError());
        const dim3 hot_grid((unsigned)a.ntiles, (unsigned)nlev);
        const size_t hot_smem = HF_STAGES * LS_CHUNK * sizeof(double4);
        if (a.fold_g >= 32) far_fold_kernel<R, 32><<<hot_grid, 32, hot_smem, st>>>(a);
        else if (a.fold_g >= 16) far_fold_kernel<R, 16><<<hot_grid, 32, hot_smem, st>>>(a);
        else far_fold_kernel<R, 4><<<hot_grid, 32, hot_smem, st>>>(a);
        CS_CUDA(cudaGetLastError());
        cs_count_launch(ctx, 3);
        return CS_OK;
    }
    if (band) {
        line_sum_kernel<SHAPE, R, false, CAN_BAND><<<grid, LS_THREADS, smem, st>>>(a);      // direct mode only (see above)
    } else {
        if (a.mp_theta > 0.0) line_sum_kernel<SHAPE, R, true, false><<<grid, LS_THREADS, smem, st>>>(a);
        else line_sum_kernel<SHAPE, R, false, false><<<grid, LS_THREADS, smem, st>>>(a);
    }
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx, 2);
    return CS_OK;
}

// per-line S(T), alpha(T), gamma(T,P,Pp) in the reference's operation order (same code path as prep_kernel)
__global__ void __launch_bounds__(256) line_params_kernel(PrepArgs a, double* S_T, double* alpha, double* gamma)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.nl) return;
    const LevelParams lp = a.lev[0];
    const double T = lp.T, nul = a.nu[j];
    const double c2 = 100.0 * CS_H * CS_C / CS_KB;
    double ea = -c2 * a.Epp[j], eb = -c2 * nul;
    double n = exp(ea / T) * (1 - exp(eb / T));
    double d = exp(ea / CS_TREF) * (1 - exp(eb / CS_TREF));
    int is = a.iso[j] - 1;
    const double* ch = a.cheb + (size_t)is * CS_MAXCHEB;
    double tau = 2 * (T - CS_TMIN) / (CS_TMAX - CS_TMIN) - 1;
    double c1 = 1.0, c2c = tau, y = ch[0] + ch[1] * c2c;
    for (int q = 2; q < a.ncheb[is]; q++) {
        double c3 = 2 * tau * c2c - c1;
        y += ch[q] * c3;
        c1 = c2c;
        c2c = c3;
    }
    if (S_T) S_T[j] = a.S[j] * (1.0 / y) * (n / d);
    if (alpha) alpha[j] = (nul / CS_C) * sqrt(2.0 * CS_R * T / a.mu[j]);
    if (gamma) gamma[j] = (pow(CS_TREF / T, a.na[j])) * (a.ga[j] * (lp.P - lp.Pp) + a.gs[j] * lp.Pp) / CS_ATM;
}

__global__ void fill_kernel(double* p, size_t n, double v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

// called by cs_lines_upload once the line arrays are on the device
int32_t cs_lines_static(cs_lines* L, cudaStream_t st)
{
    line_static_kernel<<<(unsigned)((L->n + 255) / 256), 256, 0, st>>>(L->nu, L->Epp, L->n, L->dref);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(L->ctx);
    return CS_OK;
}

extern "C" int32_t cs_line_params(cs_lines* L, double T, double P, double Pp, double* S_T, double* alpha, double* gamma)
{
    CS_REQUIRE(L, CS_ERR_ARG, "null argument");
    CS_REQUIRE(T >= CS_TMIN && T <= CS_TMAX, CS_ERR_DOMAIN, "temperature outside of Qref/Q interpolation range [%g, %g]: T = %g",
               CS_TMIN, CS_TMAX, T);
    cs_ctx* ctx = L->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    LevelParams lp = {T, P, Pp, 1.0, 0, 0, 0, 0, 0};
    CS_TRY(ctx->s_lev.reserve(sizeof(LevelParams)));
    CS_TRY(ctx->s_misc.reserve(sizeof(double) * 3 * (size_t)L->n));
    CS_CUDA(cudaMemcpyAsync(ctx->s_lev.p, &lp, sizeof(lp), cudaMemcpyHostToDevice, st));
    PrepArgs pa;
    pa.nu = L->nu; pa.S = L->S; pa.ga = L->ga; pa.gs = L->gs; pa.Epp = L->Epp; pa.na = L->na; pa.mu = L->mu;
    pa.iso = L->iso; pa.ncheb = L->ncheb; pa.cheb = L->cheb; pa.j0 = 0; pa.nl = L->n;
    pa.dref = L->dref; pa.qtab = nullptr; pa.niso = L->niso;
    pa.lev = ctx->s_lev.as<LevelParams>(); pa.nlev = 1; pa.rec = nullptr; pa.slow = nullptr; pa.shape = CS_VOIGT;
    double* d = ctx->s_misc.as<double>();
    line_params_kernel<<<(unsigned)((L->n + 255) / 256), 256, 0, st>>>(pa, d, d + L->n, d + 2 * L->n);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    if (S_T) CS_CUDA(cudaMemcpyAsync(S_T, d, sizeof(double) * (size_t)L->n, cudaMemcpyDeviceToHost, st));
    if (alpha) CS_CUDA(cudaMemcpyAsync(alpha, d + L->n, sizeof(double) * (size_t)L->n, cudaMemcpyDeviceToHost, st));
    if (gamma) CS_CUDA(cudaMemcpyAsync(gamma, d + 2 * L->n, sizeof(double) * (size_t)L->n, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}

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
    // includedlines(nu::Vector, ...) strict prefilter (line_shapes.jl:18-22) on the grid the reference would see: the
    // caller's own grid, or the global one a nu slice belongs to (cs_lines_set_grid_range) -- interior slice edges are
    // then decided by the inclusive per-point rule alone, exactly like in the unsharded run
    double numin = L->has_range ? L->rng_lo : h_nu[0], numax = L->has_range ? L->rng_hi : h_nu[nnu - 1];
    const std::vector<double>& ln = L->h_nu;
    int64_t j0 = std::upper_bound(ln.begin(), ln.end(), numin - cut) - ln.begin();     // first nul > numin-cut
    int64_t j1 = std::lower_bound(ln.begin(), ln.end(), numax + cut) - ln.begin();     // first nul >= numax+cut
    int64_t nl = j1 > j0 ? j1 - j0 : 0;

    cudaStream_t st = ctx->stream;
    if (nl == 0) {
        if (accumulate == 0) CS_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)nnu * nlev, st));
        if (accumulate == 2) {      // log(0): every byte 0xFF would be a NaN, so fill -inf explicitly
            fill_kernel<<<1024, 256, 0, st>>>(d_out, (size_t)nnu * nlev, -INFINITY);
            CS_CUDA(cudaGetLastError());
        }
        return CS_OK;
    }
    // level batches sized so that the per-level records stay within a fixed HBM budget
    bool need_slow = (shape == CS_VOIGT || shape == CS_PHCO2);
    size_t per_level = (size_t)nl * sizeof(double4) * (need_slow ? 2 : 1);
    size_t have = need_slow ? 2 * std::min(ctx->s_rec.cap, ctx->s_slow.cap) : ctx->s_rec.cap;
    int64_t nb = nlev;
    if (per_level * (size_t)nlev > have) {   // scratch must grow: size the level batch against free HBM
        size_t free_b = 0, total_b = 0;
        CS_CUDA(cudaMemGetInfo(&free_b, &total_b));
        size_t budget = std::min<size_t>((size_t)8 << 30, free_b / 4 + ctx->s_rec.cap + ctx->s_slow.cap);
        nb = std::max<int64_t>(1, std::min<int64_t>(nlev, (int64_t)(budget / per_level)));
    }
    nb = std::min<int64_t>(nb, 32768);
    CS_TRY(ctx->s_rec.reserve((size_t)nb * nl * sizeof(double4)));
    if (need_slow) CS_TRY(ctx->s_slow.reserve((size_t)nb * nl * sizeof(double4)));
    CS_TRY(ctx->s_lev.reserve(sizeof(LevelParams) * (size_t)nb));
    CS_TRY(ctx->s_q.reserve(sizeof(double) * (size_t)nb * (size_t)L->niso));

    std::vector<LevelParams> hl((size_t)nb);
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
            lp.lgtr = log(CS_TREF / lp.T);
            // near-centre half width as a fraction of the line position: sqrt(thr_j) = nul_j * f(T, mu_j)
            double vth = sqrt(2.0 * CS_R * lp.T / L->mu_min) / CS_C;
            if (shape == CS_VOIGT || shape == CS_PHCO2)
                lp.cnear = sqrt(W985_S1 * (1.0 + 1e-9)) / 0.83255461115769775635 * vth * (1.0 + 1e-6);
            else if (shape == CS_DOPPLER)
                lp.cnear = sqrt(746.0) * vth * (1.0 + 1e-6);
            else
                lp.cnear = -1.0;
            // PHCO2 far-wing expansion: allowed where eps = (chi*gamma/dnu)^2 < 1e-4 for every line with |dnu| >= 30:
            // gamma <= (296/T)^na * max(gamma_a, gamma_s) * P/atm, chi <= exp(-27 B1) there (B2 and 0.0232 are positive)
            {
                double tr = CS_TREF / lp.T;
                double gb = std::max(pow(tr, L->na_min), pow(tr, L->na_max)) * L->g_max * (lp.P / CS_ATM);
                double chimax = exp(-27.0 * lp.B1);
                double e = chimax * gb / 30.0;
                lp.pexp_ok = (e * e < 1e-4) ? 1.0 : 0.0;
            }
        }
        // a few KB through the pinned staging ring: no host synchronisation
        CS_TRY(cs_stage_h2d(ctx, ctx->s_lev.p, hl.data(), sizeof(LevelParams) * (size_t)kb));

        PrepArgs pa;
        pa.nu = L->nu; pa.S = L->S; pa.ga = L->ga; pa.gs = L->gs; pa.Epp = L->Epp; pa.na = L->na; pa.mu = L->mu;
        pa.iso = L->iso; pa.ncheb = L->ncheb; pa.cheb = L->cheb;
        pa.dref = L->dref; pa.qtab = ctx->s_q.as<double>(); pa.niso = L->niso;
        pa.j0 = j0; pa.nl = nl; pa.lev = ctx->s_lev.as<LevelParams>(); pa.nlev = (int)kb;
        pa.rec = ctx->s_rec.as<double4>();
        pa.slow = need_slow ? ctx->s_slow.as<double4>() : nullptr;
        pa.shape = shape;
        const int sp_prep = cs_span_begin(ctx, CS_T_PREP, false);
        dim3 pg((unsigned)((nl + 255) / 256), (unsigned)kb);
        qrefq_kernel<<<dim3((unsigned)((L->niso + 63) / 64), (unsigned)kb), 64, 0, st>>>(pa.lev, L->ncheb, L->cheb, L->niso,
                                                                                      ctx->s_q.as<double>());
        CS_CUDA(cudaGetLastError());
        prep_kernel<<<pg, 256, 0, st>>>(pa);
        CS_CUDA(cudaGetLastError());
        cs_count_launch(ctx, 2);
        cs_span_end(ctx, sp_prep);
        const int sp_sum = cs_span_begin(ctx, CS_T_LINESUM, false);

        LineSumArgs la;
        la.nu = d_nu; la.nnu = nnu; la.nul = L->nu + j0; la.nl = nl;
        la.rec = pa.rec; la.slow = pa.slow; la.lev = pa.lev; la.cut = cut;
        la.out = d_out + (size_t)k0 * nnu; la.accumulate = accumulate;
        la.mp_theta = ctx->farfield == CS_FARFIELD_EXPANSION ? MP_THETA : 0.0;
        la.nul_lo = ln[(size_t)j0]; la.nul_hi = ln[(size_t)j1 - 1];
        double cn = 0.0;
        for (int64_t k = 0; k < kb; k++) cn = std::max(cn, hl[(size_t)k].cnear);
        // Voigt band correction (band_near): only where the damping parameter y = gamma d is bounded below for every line of
        // every level of the batch -- gamma >= min((296/T)^na) (ga_min (P - Pp) + gs_min Pp)/atm, d >= sqrt(ln 2)/(nul_max vth)
        la.band = 0;
        if (shape == CS_VOIGT) {
            double ymin = 1e300;
            for (int64_t k = 0; k < kb; k++) {
                const LevelParams& lp = hl[(size_t)k];
                const double tr = CS_TREF / lp.T;
                const double gmin = std::min(pow(tr, L->na_min), pow(tr, L->na_max)) *
                                    (L->ga_min * (lp.P - lp.Pp) + L->gs_min * lp.Pp) / CS_ATM;
                const double vth = sqrt(2.0 * CS_R * lp.T / L->mu_min) / CS_C;
                const double dmin = 0.83255461115769775635 / (std::max(la.nul_hi, 1e-300) * vth);
                ymin = std::min(ymin, gmin * dmin);
            }
            la.band = (ymin >= 1e-5) ? 1 : 0;      // also false for NaN / negative bounds
            if (ctx->ls_no_band) la.band = 0;
        }
        la.split = ctx->ls_no_split ? 0 : 1;
        // The fold keeps a running denominator d = product of G values q = dnu^2 + gamma^2.  Overflow: q <= dmax^2 + gamma_max^2 with
        // dmax the largest distance the window allows.  Underflow (band / Lorentz: lines folded at their own centres): fewer than k
        // lines can lie within s_k / 2 of a point (s_k = narrowest span of k consecutive lines), the others have q >= s_k^2 / 4, so
        // d >= (gamma_min^2)^(k-1) (s_k^2/4)^(G-k+1) for every k.  G = 32 where that is safe, else 16, else 4.
        la.fold_g = 4;
        if (shape == CS_VOIGT || shape == CS_LORENTZ) {
            double gmin = 1e300, gmax = 0.0;
            for (int64_t k = 0; k < kb; k++) {
                const LevelParams& lp = hl[(size_t)k];
                const double tr = CS_TREF / lp.T;
                const double plo = std::min(pow(tr, L->na_min), pow(tr, L->na_max)), phi = std::max(pow(tr, L->na_min), pow(tr, L->na_max));
                gmin = std::min(gmin, plo * (L->ga_min * (lp.P - lp.Pp) + L->gs_min * lp.Pp) / CS_ATM);
                gmax = std::max(gmax, phi * L->g_max * lp.P / CS_ATM);
            }
            const double dmax = std::min(cut, std::max(fabs(h_nu[nnu - 1] - la.nul_lo), fabs(la.nul_hi - h_nu[0])));
            const double lq_hi = log10(dmax * dmax + gmax * gmax + 1e-300);
            const double lg2 = gmin > 0.0 ? 2.0 * log10(gmin) : -1e300;
            auto safe = [&](int G) {
                if (G * std::max(lq_hi, 0.0) > 280.0) return false;
                const int ks[3] = {4, 8, 16};
                for (int q = 0; q < 3; q++) {
                    if (ks[q] > G) continue;
                    const double sk = L->span_k[q];
                    const double lsk = sk > 0.0 ? std::min(2.0 * log10(0.5 * sk), 0.0) : -1e300;
                    if ((ks[q] - 1) * std::min(lg2, 0.0) + (G - ks[q] + 1) * lsk > -280.0) return true;
                }
                return false;
            };
            la.fold_g = safe(32) ? 32 : (safe(16) ? 16 : 4);
            if (shape == CS_VOIGT && la.fold_g < 16) la.band = 0;      // the single launch folds far lines only (q >= (c nul)^2)
            if (ctx->ls_fold_g > 0) la.fold_g = std::min(la.fold_g, ctx->ls_fold_g);
        }
        switch (shape) {
        case CS_DOPPLER: CS_TRY((launch_line_sum<CS_DOPPLER, 4>(ctx, la, (int)kb, cn))); break;
        // expansion mode: what is left for the pair-by-pair sum scales with the tile width (cut-off edges, lines within 4 half
        // widths, the near band), the far field does not -- 64-point tiles there, 128-point tiles for the direct sum
        case CS_LORENTZ:
            if (la.mp_theta > 0.0 && CS_LS_R_EXP != 4) CS_TRY((launch_line_sum<CS_LORENTZ, CS_LS_R_EXP>(ctx, la, (int)kb, cn)));
            else CS_TRY((launch_line_sum<CS_LORENTZ, 4>(ctx, la, (int)kb, cn)));
            break;
        case CS_VOIGT:
            if (la.mp_theta > 0.0 && CS_LS_R_EXP != 4) CS_TRY((launch_line_sum<CS_VOIGT, CS_LS_R_EXP>(ctx, la, (int)kb, cn)));
            else CS_TRY((launch_line_sum<CS_VOIGT, 4>(ctx, la, (int)kb, cn)));
            break;
        default:         CS_TRY((launch_line_sum<CS_PHCO2, 4>(ctx, la, (int)kb, cn))); break;
        }
        cs_span_end(ctx, sp_sum);
    }
    cs_spans_collect(ctx, false);
    return CS_OK;
}
