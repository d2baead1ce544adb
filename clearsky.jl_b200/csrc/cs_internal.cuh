// cs_internal.cuh -- shared internals of libclearsky_b200 (sm_100a only; FP64 CUDA cores, no tensor cores).
//
// The public boundary is include/clearsky_b200.h (plain C ABI).  Everything here is private.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/clearsky_b200.h"

// ------------------------------------------------------------------------------------------------
// constants: bit-exact copies of the reference's values (src/constants.jl:1-27,
// src/absorption/line_shapes.jl:2-5, src/hitran/molparam.jl:1-2)
#define CS_C    299792458.0
#define CS_H    6.62607015e-34
#define CS_KB   1.38064852e-23      // 2014 CODATA value, as in the reference (constants.jl:6)
#define CS_R    8.31446262
#define CS_ATM  101325.0
#define CS_NA   6.02214076e23
#define CS_LO2  7.21879268e38
#define CS_TREF 296.0
#define CS_T0   273.15
#define CS_TMIN 25.0
#define CS_TMAX 1000.0
#define CS_PI   3.14159265358979323846

// ------------------------------------------------------------------------------------------------
// error plumbing: status codes + thread-local message, no exceptions across the ABI
void cs_set_error(const char* fmt, ...);

#define CS_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            cs_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call,               \
                         cudaGetErrorString(e__));                                           \
            return CS_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define CS_REQUIRE(cond, code, ...)                                                          \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            cs_set_error(__VA_ARGS__);                                                       \
            return (code);                                                                   \
        }                                                                                    \
    } while (0)

#define CS_TRY(expr)                                                                         \
    do {                                                                                     \
        int32_t rc__ = (expr);                                                               \
        if (rc__ != CS_OK) return rc__;                                                      \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device buffer that only grows (scratch reused across calls: no cudaMalloc in the RCM loop)
// Stream-ordered (cudaMallocAsync / cudaFreeAsync on the context stream): growing a buffer neither synchronises the
// device nor races with kernels already enqueued on the old block.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaStream_t st = nullptr;     // the owning context's stream (set at context creation)
    int32_t reserve(size_t bytes)
    {
        if (bytes <= cap) return CS_OK;
        if (p) cudaFreeAsync(p, st);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocAsync(&p, want, st);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cs_set_error("cudaMallocAsync(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            p = nullptr;
            return CS_ERR_NOMEM;
        }
        cap = want;
        return CS_OK;
    }
    void release()
    {
        if (p) cudaFreeAsync(p, st);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// one timed span on the context stream: two recorded events whose elapsed time is read LAZILY (cs_spans_collect), so that
// no compute entry point has to block the host just to fill a timer
struct TimerSpan {
    cudaEvent_t a, b;
    int id;
    bool assign;   // true: last_kernel_ms[id] = ms ("the last call"), false: += ms
    bool closed;   // end event recorded
};

// pinned host staging slot for small host->device parameter blocks (level parameters, per-level tables): the copy is
// asynchronous with respect to the host, the slot is reused only after its copy has executed
struct StageSlot {
    void* p = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    bool used = false;
};
constexpr int CS_NSTAGE = 16;
constexpr size_t CS_STAGE_MAX = (size_t)1 << 20;   // larger blocks are copied straight from the caller's buffer

struct cs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaStream_t copy_stream = nullptr;    // uploads of line lists: a blocking upload then waits for its own copies only, not for the
                                           // line sum of the previous gas still running on `stream`
    int sm_count = 148;
    std::recursive_mutex mtx;
    // scratch
    DevBuf s_nu, s_lev, s_rec, s_slow, s_sigma, s_misc, s_part, s_tau, s_planck, s_out0, s_out1, s_out2;
    DevBuf s_w, s_q;
    // kernel timing (ms), measured with CUDA events on ctx->stream and read lazily
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    std::vector<TimerSpan> spans;          // recorded, not yet read
    std::vector<cudaEvent_t> ev_free;      // recycled events
    double last_kernel_ms[CS_NTIMERS] = {0};    // timers of the most recent call (cs_ctx_timers)
    double total_kernel_ms[CS_NTIMERS] = {0};   // never reset (cs_ctx_timers_total)
    StageSlot stage[CS_NSTAGE];
    int stage_next = 0;
    bool ff_no_moments = false;            // CS_FARFIELD_NO_MOMENTS, read once at context creation
    bool ls_no_split = false;              // CS_LINESUM_NO_SPLIT: one launch for the cold classes and the far wings
    int ls_fold_g = 0;                     // CS_LINESUM_FOLD: cap on the lines per reciprocal of far_fold_kernel (diagnostic)
    bool ls_no_band = false;               // CS_LINESUM_NO_BAND: Voigt near-centre lines per tile (cold_near) instead of the per-point band
    bool table_no_mma = false;             // CS_TABLE_EVAL_NO_MMA, likewise
    bool table_no_fused = false;           // CS_TABLE_FIT_NO_FUSED: two-pass GEMM fit instead of the fused single sweep
    int64_t launches = 0;   // number of kernels of this library launched on this context
    int32_t farfield = CS_FARFIELD_DIRECT;   // K2 far-wing treatment (cs_ctx_set_farfield)
    double tau_floor = 1e-6;                 // K6 floor on the vertical layer depth (cs_ctx_set_tau_floor)
};

struct cs_lines {
    cs_ctx* ctx;
    int64_t n;
    // device SoA (sorted by wavenumber, as SpectralLines guarantees: par.jl:267)
    double *nu, *S, *ga, *gs, *Epp, *na, *mu;
    double* dref = nullptr;   // per line: level-independent denominator of scaleintensity (line_static_kernel)
    int16_t* iso;
    int32_t niso;
    int32_t* ncheb;   // [niso]
    double* cheb;     // [niso][CS_MAXCHEB]
    std::vector<double> h_nu;  // host copy of line positions (window searches, eval counting)
    double mu_min;             // lightest isotopologue present (bounds the Doppler width)
    double g_max, na_min, na_max;   // max(gamma_a, gamma_s) and the range of the temperature exponent (bound gamma per level)
    double ga_min = 0.0, gs_min = 0.0;   // smallest air- / self-broadening coefficients (bound the Voigt damping parameter from below)
    double span_k[3] = {0.0, 0.0, 0.0};  // narrowest span of 4 / 8 / 16 consecutive lines (bounds how many lines can sit on one point)
    // nu-sharded runs: first/last point of the GLOBAL grid, so that the strict includedlines prefilter
    // (line_shapes.jl:18-22) is applied to the grid the reference would see, not to a slice of it
    bool has_range = false;
    double rng_lo = 0.0, rng_hi = 0.0;
};

struct cs_sigma {
    cs_ctx* ctx;
    int64_t nnu, nnode;
    double* nu;      // device [nnu]
    double* sig;     // device [nnode][nnu], nu fastest
    double* w;       // device [nnu] trapezoid weights of nu (util.jl:26-33 rewritten per point)
    std::vector<double> h_nu;
    bool own_nu;
};

struct cs_table {
    cs_ctx* ctx;
    int64_t nnu;
    int32_t nT, nP;
    double Ta, Tb, lnPa, lnPb;   // interpolator bounds = first/last grid coordinates
    double* coef;                // device [nT*nP][nnu], coefficient-major, nu fastest
    double* sigma_block;         // optional device copy of the baked block [nT*nP][nnu] (NULL if dropped)
    int64_t nzeroed;
};

struct cs_cia {
    cs_ctx* ctx;
    int32_t ngrid, nsingle, extrapolate, singles;
    std::vector<int64_t> g_nnu, g_nT, g_off_nu, g_off_T, g_off_k, s_n, s_off;
    std::vector<double> h_T;     // host copy of grid temperatures (cell search per node on the host)
    double *d_nu, *d_T, *d_lnk, *d_snu, *d_slnk;
    int64_t *d_desc;             // packed descriptors
};

struct cs_accel {
    cs_ctx* ctx;
    int64_t nnu, nlev;
    std::vector<double> h_lnP;
    double* lnsig;               // device [nlev][nnu]
};

// ------------------------------------------------------------------------------------------------
// stream-ordered allocation of the long-lived device objects (lines, tables, workspaces): the default memory pool
// keeps freed blocks (release threshold = max), so re-creating objects of the same size every call costs no
// cudaMalloc/cudaFree and no device synchronisation
static inline cudaError_t cs_malloc(void** p, size_t bytes, cudaStream_t st) { return cudaMallocAsync(p, bytes, st); }
static inline void cs_free(void* p, cudaStream_t st)
{
    if (p) cudaFreeAsync(p, st);
}

// ------------------------------------------------------------------------------------------------
// launch bookkeeping
static inline void cs_count_launch(cs_ctx* c, int64_t n = 1) { c->launches += n; }

// fast reciprocal for normal, finite, positive arguments: MUFU.RCP64H seed (measured max relative error 9.9e-7 =
// 2^-19.9 on B200, tools/micro/rcp_accuracy.cu) + ONE cubic Newton step r(1 + e + e^2), e = 1 - a r, which leaves
// e^3 ~ 1e-18 < 2^-53: the result is within 1 ulp of 1/a (measured max 2.2e-16).  nvcc's own 1.0/x adds a second,
// quadratic step and an exponent-range fix-up for correct rounding; a 1e-9 parity budget does not need them.
__device__ __forceinline__ double cs_rcp(double a)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    e = fma(e, e, e);
    return fma(r, e, r);
}

// ------------------------------------------------------------------------------------------------
// Device Re w(x+iy): Algorithm 985 (Zaghloul 2017), the arithmetic behind Faddeyeva985.faddeyeva(x,y)
// (reference call site src/absorption/line_shapes.jl:375).  Written from the published algorithm, real
// arithmetic only; region borders documented in DESIGN.md.  s must be fma(x,x,y*y).
struct cplx {
    double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b)
{
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
// |b|^2 is positive, normal and finite wherever the routine divides (polynomials of bounded arguments): one 1-ulp reciprocal
// serves both components instead of two IEEE divisions (~30 instructions each)
__device__ __forceinline__ cplx cdiv(cplx a, cplx b)
{
    double rden = cs_rcp(b.re * b.re + b.im * b.im);
    return {(a.re * b.re + a.im * b.im) * rden, (a.im * b.re - a.re * b.im) * rden};
}
__device__ __forceinline__ cplx cadd(cplx a, double r) { return {a.re + r, a.im}; }
// c - u*acc  (Horner step of the Humlicek polynomials)
__device__ __forceinline__ cplx hstep(double c, cplx u, cplx acc)
{
    cplx p = cmul(u, acc);
    return {c - p.re, -p.im};
}

// Region borders of Algorithm 985 (|z|^2 thresholds of the 1 / 2 / 3 / 4 Laplace convergents, the y^2 floor of the 4-convergent
// form, the Humlicek w4 region).  Two self-consistent sets exist, each constant sitting exactly where the cheaper form reaches the
// set's accuracy against an accurate w(z) (DESIGN.md section 5 has the derivation):
//   CS_W985_MAP 1 (default)  3.8e4 / 256 / 62 / 30, y^2 >= 1e-13 / 2.5, y^2 < 0.072   max rel. error 4e-5 -- the accuracy the
//                            Algorithm 985 paper states for both parts of w, and SURVEY.md's necessary condition for a restatement
//   CS_W985_MAP 0            1.6e4 / 160 / 107 / 28.5, y^2 >= 6e-14 / 3.5, y^2 < 0.026   max rel. error 1e-4 (SURVEY.md 8c's recollection)
// tools/julia_golden.jl samples the real package across the borders of both; tests/test_reference_golden.py says which one it is.
#ifndef CS_W985_MAP
#define CS_W985_MAP 1
#endif
#if CS_W985_MAP
#define W985_S1 3.8e4
#define W985_S2 256.0
#define W985_S3 62.0
#define W985_S4 30.0
#define W985_Y4 1e-13
#define W985_S5 2.5
#define W985_Y5 0.072
#else
#define W985_S1 1.6e4
#define W985_S2 160.0
#define W985_S3 107.0
#define W985_S4 28.5
#define W985_Y4 6e-14
#define W985_S5 3.5
#define W985_Y5 0.026
#endif

static __device__ __noinline__ double cs_faddeyeva985(double x, double y)
{
    const double osqpi = 0.56418958354775628695;  // 1/sqrt(pi)
    double y2 = y * y;
    double s = fma(x, x, y2);
    if (s >= W985_S1) return y * osqpi / s;                     // 1 convergent
    cplx z = {x, y};
    if (s >= W985_S4 && (s >= W985_S3 || y2 >= W985_Y4)) {
        cplx zz = cmul(z, z);
        cplx num, den;
        if (s >= W985_S2) {                                     // 2 convergents: i z / (z^2 - 1/2)
            num = z;
            den = cadd(zz, -0.5);
        } else if (s >= W985_S3) {                              // 3: i (z^2 - 1) / (z (z^2 - 3/2))
            num = cadd(zz, -1.0);
            den = cmul(z, cadd(zz, -1.5));
        } else {                                                // 4: i z (z^2 - 5/2) / (z^2 (z^2 - 3) + 3/4)
            num = cmul(z, cadd(zz, -2.5));
            den = cadd(cmul(zz, cadd(zz, -3.0)), 0.75);
        }
        cplx q = cdiv(num, den);
        return -q.im * osqpi;                                   // Re(i q / sqrt(pi))
    }
    cplx t = {y, -x};
    if (s >= W985_S5 && y2 < W985_Y5) {                         // Humlicek w4, region IV
        cplx u = cmul(t, t);
        cplx P = {0.56419, 0.0};
        P = hstep(1.320522, u, P);
        P = hstep(35.76683, u, P);
        P = hstep(219.0313, u, P);
        P = hstep(1540.787, u, P);
        P = hstep(3321.9905, u, P);
        P = hstep(36183.31, u, P);
        cplx Q = {1.0, 0.0};
        Q = hstep(1.841439, u, Q);
        Q = hstep(61.57037, u, Q);
        Q = hstep(364.2191, u, Q);
        Q = hstep(2186.181, u, Q);
        Q = hstep(9022.228, u, Q);
        Q = hstep(24322.84, u, Q);
        Q = hstep(32066.6, u, Q);
        cplx r = cmul(t, cdiv(P, Q));
        double sn, cs;
        sincos(u.im, &sn, &cs);
        return exp(u.re) * cs - r.re;
    }
    // Hui, Armstrong & Wray p = 6
    cplx num = {0.564189583562615, 0.0};
    num = cadd(cmul(num, t), 5.912626209773153);
    num = cadd(cmul(num, t), 30.180142196210589);
    num = cadd(cmul(num, t), 93.155580458138441);
    num = cadd(cmul(num, t), 181.928533092181549);
    num = cadd(cmul(num, t), 214.382388694706425);
    num = cadd(cmul(num, t), 122.607931777104326);
    cplx den = {1.0, 0.0};
    den = cadd(cmul(den, t), 10.479857114260399);
    den = cadd(cmul(den, t), 53.992906912940207);
    den = cadd(cmul(den, t), 170.354001821091472);
    den = cadd(cmul(den, t), 348.703917719495792);
    den = cadd(cmul(den, t), 457.334478783897737);
    den = cadd(cmul(den, t), 352.730625110963558);
    den = cadd(cmul(den, t), 122.607931773875350);
    return cdiv(num, den).re;
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}


// internal entry points implemented across the .cu files
void cs_reset_timers(cs_ctx* c);
// lazy timers: cs_span_begin records the start event and returns the span index, cs_span_end the end event;
// cs_spans_collect folds finished spans into the timer arrays (block = wait for all of them)
int cs_span_begin(cs_ctx* c, int id, bool assign);
void cs_span_end(cs_ctx* c, int span);
void cs_spans_collect(cs_ctx* c, bool block);
// asynchronous host->device copy of a small parameter block through the context's pinned staging ring
int32_t cs_stage_h2d(cs_ctx* c, void* dst, const void* src, size_t bytes);
int32_t cs_lines_static(cs_lines* L, cudaStream_t st);
int32_t cs_lines_accumulate(cs_lines* L, int32_t shape, int64_t nnu, const double* d_nu,
                            const double* h_nu, int64_t nlev, const double* h_T, const double* h_P,
                            const double* h_Pp, const double* h_scale, double cut, double* d_out,
                            int accumulate);
