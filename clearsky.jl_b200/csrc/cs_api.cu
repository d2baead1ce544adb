// cs_api.cu -- C ABI entry points: context, lines, cross-sections, sigma workspace.
#include "cs_internal.cuh"
#include <cstdlib>
#include <cstring>
#include <stdarg.h>
#include <algorithm>
#include <limits>

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

void cs_set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

extern "C" const char* cs_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t cs_version(void) { return 100; }

extern "C" int32_t cs_device_count(int32_t* n)
{
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cs_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        *n = 0;
        return CS_ERR_CUDA;
    }
    *n = c;
    return CS_OK;
}

// ------------------------------------------------------------------------------------------------
static int32_t ctx_create(int32_t device, void* stream, bool own, cs_ctx** out)
{
    CS_REQUIRE(out != nullptr, CS_ERR_ARG, "null output pointer");
    *out = nullptr;
    int ndev = 0;
    CS_CUDA(cudaGetDeviceCount(&ndev));
    CS_REQUIRE(ndev > 0, CS_ERR_CUDA, "no CUDA device: libclearsky_b200 has no CPU fallback");
    CS_REQUIRE(device >= 0 && device < ndev, CS_ERR_ARG, "device %d out of range [0,%d)", device, ndev);
    CS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CS_CUDA(cudaGetDeviceProperties(&prop, device));
    CS_REQUIRE(prop.major >= 10, CS_ERR_CUDA,
               "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    cs_ctx* c = new cs_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (own) {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            cs_set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
            return CS_ERR_CUDA;
        }
    } else {
        c->stream = (cudaStream_t)stream;
    }
    c->own_stream = own;
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        c->copy_stream = nullptr;
    }
    {
        DevBuf* bufs[] = {&c->s_nu, &c->s_lev, &c->s_rec, &c->s_slow, &c->s_sigma, &c->s_misc, &c->s_part,
                          &c->s_tau, &c->s_planck, &c->s_out0, &c->s_out1, &c->s_out2, &c->s_w, &c->s_q};
        for (DevBuf* b : bufs) b->st = c->stream;
    }
    // environment switches are read once here, never in a launch path
    if (const char* ff = getenv("CS_FARFIELD"))
        c->farfield = (strcmp(ff, "expansion") == 0) ? CS_FARFIELD_EXPANSION : CS_FARFIELD_DIRECT;
    c->ff_no_moments = getenv("CS_FARFIELD_NO_MOMENTS") != nullptr;
    c->ls_no_band = getenv("CS_LINESUM_NO_BAND") != nullptr;
    c->ls_no_split = getenv("CS_LINESUM_NO_SPLIT") != nullptr;
    if (const char* fg = getenv("CS_LINESUM_FOLD")) c->ls_fold_g = atoi(fg);
    c->table_no_mma = getenv("CS_TABLE_EVAL_NO_MMA") != nullptr;
    c->table_no_fused = getenv("CS_TABLE_FIT_NO_FUSED") != nullptr;
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long thr = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    cudaEventCreate(&c->ev2);
    *out = c;
    return CS_OK;
}

extern "C" int32_t cs_ctx_create(int32_t device, cs_ctx** out) { return ctx_create(device, nullptr, true, out); }
extern "C" int32_t cs_ctx_create_on_stream(int32_t device, void* s, cs_ctx** out) { return ctx_create(device, s, false, out); }

extern "C" int32_t cs_ctx_free(cs_ctx* c)
{
    if (!c) return CS_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->s_nu, &c->s_lev, &c->s_rec, &c->s_slow, &c->s_sigma, &c->s_misc, &c->s_part,
                      &c->s_tau, &c->s_planck, &c->s_out0, &c->s_out1, &c->s_out2, &c->s_w, &c->s_q};
    for (DevBuf* b : bufs) b->release();
    for (TimerSpan& sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (cudaEvent_t e : c->ev_free) cudaEventDestroy(e);
    for (StageSlot& sl : c->stage) {
        if (sl.p) cudaFreeHost(sl.p);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaEventDestroy(c->ev2);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return CS_OK;
}

extern "C" int32_t cs_ctx_synchronize(cs_ctx* c)
{
    CS_REQUIRE(c, CS_ERR_ARG, "null context");
    CS_CUDA(cudaSetDevice(c->device));
    CS_CUDA(cudaStreamSynchronize(c->stream));
    return CS_OK;
}

extern "C" int32_t cs_ctx_timers(cs_ctx* c, double* t)
{
    CS_REQUIRE(c && t, CS_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    CS_CUDA(cudaSetDevice(c->device));
    cs_spans_collect(c, true);
    for (int i = 0; i < CS_NTIMERS; i++) t[i] = c->last_kernel_ms[i];
    return CS_OK;
}

extern "C" int32_t cs_ctx_timers_total(cs_ctx* c, double* t)
{
    CS_REQUIRE(c && t, CS_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    CS_CUDA(cudaSetDevice(c->device));
    cs_spans_collect(c, true);
    for (int i = 0; i < CS_NTIMERS; i++) t[i] = c->total_kernel_ms[i];
    return CS_OK;
}

extern "C" int32_t cs_ctx_launches(cs_ctx* c, int64_t* n)
{
    CS_REQUIRE(c && n, CS_ERR_ARG, "null argument");
    *n = c->launches;
    return CS_OK;
}

extern "C" int32_t cs_ctx_set_farfield(cs_ctx* c, int32_t mode)
{
    CS_REQUIRE(c, CS_ERR_ARG, "null argument");
    CS_REQUIRE(mode == CS_FARFIELD_DIRECT || mode == CS_FARFIELD_EXPANSION, CS_ERR_ARG, "unknown far-field mode %d", mode);
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    c->farfield = mode;
    return CS_OK;
}

extern "C" int32_t cs_ctx_get_farfield(cs_ctx* c, int32_t* mode)
{
    CS_REQUIRE(c && mode, CS_ERR_ARG, "null argument");
    *mode = c->farfield;
    return CS_OK;
}

extern "C" int32_t cs_ctx_set_tau_floor(cs_ctx* c, double tau_min)
{
    CS_REQUIRE(c, CS_ERR_ARG, "null argument");
    CS_REQUIRE(tau_min > 0 && tau_min <= 1.0, CS_ERR_ARG, "tau floor must be in (0, 1] (got %g)", tau_min);
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    c->tau_floor = tau_min;
    return CS_OK;
}

void cs_reset_timers(cs_ctx* c)
{
    cs_spans_collect(c, false);
    // spans still in flight belong to earlier calls: they keep counting towards the totals only
    for (TimerSpan& sp : c->spans)
        if (sp.id >= 0) sp.id = -1 - sp.id;
    for (int i = 0; i < CS_NTIMERS; i++) c->last_kernel_ms[i] = 0.0;
}

static cudaEvent_t ev_acquire(cs_ctx* c)
{
    cudaEvent_t e = nullptr;
    if (!c->ev_free.empty()) {
        e = c->ev_free.back();
        c->ev_free.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}

int cs_span_begin(cs_ctx* c, int id, bool assign)
{
    if (c->spans.size() >= 1024) cs_spans_collect(c, true);   // nobody reads the timers: keep the list bounded
    TimerSpan sp;
    sp.a = ev_acquire(c);
    sp.b = ev_acquire(c);
    sp.id = id;
    sp.assign = assign;
    sp.closed = false;
    cudaEventRecord(sp.a, c->stream);
    c->spans.push_back(sp);
    return (int)c->spans.size() - 1;
}

void cs_span_end(cs_ctx* c, int span)
{
    TimerSpan& sp = c->spans[(size_t)span];
    cudaEventRecord(sp.b, c->stream);
    sp.closed = true;
}

void cs_spans_collect(cs_ctx* c, bool block)
{
    size_t done = 0;
    for (; done < c->spans.size(); done++) {
        TimerSpan& sp = c->spans[done];
        if (!sp.closed) break;
        if (block) {
            if (cudaEventSynchronize(sp.b) != cudaSuccess) { cudaGetLastError(); }
        } else if (cudaEventQuery(sp.b) != cudaSuccess) {
            cudaGetLastError();   // cudaErrorNotReady is not sticky, but clear it
            break;
        }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
        const bool stale = sp.id < 0;               // recorded before the last cs_reset_timers
        const int id = stale ? -1 - sp.id : sp.id;
        c->total_kernel_ms[id] += ms;
        if (!stale) {
            if (sp.assign) c->last_kernel_ms[id] = ms;
            else c->last_kernel_ms[id] += ms;
        }
        c->ev_free.push_back(sp.a);
        c->ev_free.push_back(sp.b);
    }
    c->spans.erase(c->spans.begin(), c->spans.begin() + (std::ptrdiff_t)done);
}

int32_t cs_stage_h2d(cs_ctx* c, void* dst, const void* src, size_t bytes)
{
    if (bytes == 0) return CS_OK;
    if (bytes > CS_STAGE_MAX) {
        CS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
        return CS_OK;
    }
    StageSlot& sl = c->stage[c->stage_next];
    c->stage_next = (c->stage_next + 1) % CS_NSTAGE;
    if (sl.used) CS_CUDA(cudaEventSynchronize(sl.done));    // its previous copy has long executed in the steady state
    if (sl.cap < bytes) {
        if (sl.p) cudaFreeHost(sl.p);
        sl.p = nullptr;
        sl.cap = 0;
        size_t want = std::max<size_t>(bytes + bytes / 2, 4096);
        CS_CUDA(cudaMallocHost(&sl.p, want));
        sl.cap = want;
    }
    if (!sl.done) CS_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    memcpy(sl.p, src, bytes);
    CS_CUDA(cudaMemcpyAsync(dst, sl.p, bytes, cudaMemcpyHostToDevice, c->stream));
    CS_CUDA(cudaEventRecord(sl.done, c->stream));
    sl.used = true;
    return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// FP64 FMA throughput microbenchmark (the "% FP64 peak" denominator is not in MEASURED_PEAKS.json):
// 8 independent register-resident DFMA chains per thread, enough CTAs to fill every SM.
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

extern "C" int32_t cs_fp64_peak(cs_ctx* c, int32_t iters, double* flops)
{
    CS_REQUIRE(c && flops, CS_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    CS_CUDA(cudaSetDevice(c->device));
    CS_TRY(c->s_misc.reserve(64));
    int blocks = c->sm_count * 8;
    if (iters <= 0) iters = 20000;
    dfma_kernel<<<blocks, 256, 0, c->stream>>>(c->s_misc.as<double>(), 100, 1.0);   // warm-up
    CS_CUDA(cudaEventRecord(c->ev0, c->stream));
    dfma_kernel<<<blocks, 256, 0, c->stream>>>(c->s_misc.as<double>(), iters, 1.0);
    CS_CUDA(cudaEventRecord(c->ev1, c->stream));
    CS_CUDA(cudaEventSynchronize(c->ev1));
    CS_CUDA(cudaGetLastError());
    cs_count_launch(c, 2);
    float ms = 0;
    CS_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    double nfma = (double)blocks * 256.0 * (double)iters * 64.0;
    *flops = 2.0 * nfma / (ms * 1e-3);
    return CS_OK;
}

// ------------------------------------------------------------------------------------------------
template <typename T> static int32_t upload(T** dst, const T* src, size_t n, cudaStream_t st)
{
    *dst = nullptr;
    if (n == 0) return CS_OK;
    cudaError_t e = cs_malloc((void**)dst, sizeof(T) * n, st);
    if (e != cudaSuccess) {
        cs_set_error("cudaMalloc(%zu bytes) failed: %s", sizeof(T) * n, cudaGetErrorString(e));
        return CS_ERR_NOMEM;
    }
    CS_CUDA(cudaMemcpyAsync(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, st));
    return CS_OK;
}

extern "C" int32_t cs_lines_upload(cs_ctx* ctx, int64_t n, const double* nu, const double* S, const double* ga,
                                   const double* gs, const double* Epp, const double* na, const double* mu,
                                   const int16_t* iso, int32_t niso, const int32_t* ncheb, const double* cheb,
                                   const uint8_t* hascheb, cs_lines** out)
{
    CS_REQUIRE(ctx && out, CS_ERR_ARG, "null argument");
    *out = nullptr;
    CS_REQUIRE(n > 0, CS_ERR_ARG, "no lines");
    CS_REQUIRE(nu && S && ga && gs && Epp && na && mu && iso, CS_ERR_ARG, "null line-parameter array");
    CS_REQUIRE(niso > 0 && ncheb && cheb && hascheb, CS_ERR_ARG, "missing Qref/Q Chebyshev tables");
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cs_lines* L = new cs_lines();
    L->ctx = ctx;
    L->n = n;
    L->niso = niso;
    // the copies run on the context's copy stream: the call returns when ITS copies are done (the caller's arrays are only
    // valid during the call), without waiting for kernels of an earlier gas still running on the compute stream -- so the
    // upload of the next gas overlaps the line sum of the previous one.  Everything enqueued on the compute stream after this
    // call finds the data in place (the copy stream is synchronised below).
    cudaStream_t st = ctx->copy_stream ? ctx->copy_stream : ctx->stream;
    int32_t rc = CS_OK;
    if ((rc = upload(&L->nu, nu, n, st)) || (rc = upload(&L->S, S, n, st)) || (rc = upload(&L->ga, ga, n, st)) ||
        (rc = upload(&L->gs, gs, n, st)) || (rc = upload(&L->Epp, Epp, n, st)) || (rc = upload(&L->na, na, n, st)) ||
        (rc = upload(&L->mu, mu, n, st)) || (rc = upload(&L->iso, iso, n, st)) ||
        (rc = upload(&L->ncheb, ncheb, (size_t)niso, st)) ||
        (rc = upload(&L->cheb, cheb, (size_t)niso * CS_MAXCHEB, st))) {
        cudaStreamSynchronize(st);
        cs_lines_free(L);
        return rc;
    }
    // host-side validation and statistics run while the copies are in flight
    int32_t bad = CS_OK;
    for (int64_t j = 1; j < n && !bad; j++)
        if (!(nu[j] >= nu[j - 1])) {
            cs_set_error("line wavenumbers must be sorted ascending (line %lld)", (long long)j);
            bad = CS_ERR_ARG;
        }
    for (int64_t j = 0; j < n && !bad; j++) {
        const int is = iso[j];
        if (!(is >= 1 && is <= niso)) {
            cs_set_error("isotopologue number %d out of range [1,%d]", is, niso);
            bad = CS_ERR_ARG;
        } else if (!hascheb[is - 1]) {
            // scaleintensity throws when no interpolating polynomial exists (line_shapes.jl:115-119)
            cs_set_error("no interpolating polynomial available to compute Qref/Q for isotopologue %d", is);
            bad = CS_ERR_ARG;
        } else if (!(ncheb[is - 1] >= 2 && ncheb[is - 1] <= CS_MAXCHEB)) {
            cs_set_error("bad Chebyshev length %d", ncheb[is - 1]);
            bad = CS_ERR_ARG;
        }
    }
    if (bad) {
        cudaStreamSynchronize(st);
        cs_lines_free(L);
        return bad;
    }
    L->h_nu.assign(nu, nu + n);
    L->mu_min = mu[0];
    L->g_max = 0.0; L->na_min = na[0]; L->na_max = na[0];
    L->ga_min = ga[0]; L->gs_min = gs[0];
    for (int64_t j = 0; j < n; j++) {
        L->mu_min = std::min(L->mu_min, mu[j]);
        L->g_max = std::max(L->g_max, std::max(ga[j], gs[j]));
        L->ga_min = std::min(L->ga_min, ga[j]);
        L->gs_min = std::min(L->gs_min, gs[j]);
        L->na_min = std::min(L->na_min, na[j]);
        L->na_max = std::max(L->na_max, na[j]);
    }
    {
        const int ks[3] = {4, 8, 16};
        for (int q = 0; q < 3; q++) {
            double sp = 1e300;
            for (int64_t j = 0; j + ks[q] - 1 < n; j++) sp = std::min(sp, nu[j + ks[q] - 1] - nu[j]);
            L->span_k[q] = sp;
        }
    }
    if (cudaMallocAsync(&L->dref, sizeof(double) * (size_t)n, st) != cudaSuccess) {
        cudaGetLastError();
        cs_set_error("cs_lines_upload: out of device memory");
        L->dref = nullptr;
        cudaStreamSynchronize(st);
        cs_lines_free(L);
        return CS_ERR_NOMEM;
    }
    if ((rc = cs_lines_static(L, st))) {
        cudaStreamSynchronize(st);
        cs_lines_free(L);
        return rc;
    }
    // the caller's arrays are only valid for the duration of the call: wait for the copies
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cs_set_error("cs_lines_upload: %s", cudaGetErrorString(e));
        cs_lines_free(L);
        return CS_ERR_CUDA;
    }
    *out = L;
    return CS_OK;
}

extern "C" int32_t cs_lines_set_grid_range(cs_lines* L, double numin, double numax)
{
    CS_REQUIRE(L, CS_ERR_ARG, "null argument");
    CS_REQUIRE(numax >= numin, CS_ERR_ARG, "empty wavenumber range");
    std::lock_guard<std::recursive_mutex> lk(L->ctx->mtx);
    L->has_range = true;
    L->rng_lo = numin;
    L->rng_hi = numax;
    return CS_OK;
}

extern "C" int32_t cs_lines_free(cs_lines* L)
{
    if (!L) return CS_OK;
    cudaSetDevice(L->ctx->device);
    cudaStream_t st = L->ctx->stream;
    cs_free(L->nu, st); cs_free(L->S, st); cs_free(L->ga, st); cs_free(L->gs, st); cs_free(L->Epp, st); cs_free(L->na, st);
    cs_free(L->mu, st); cs_free(L->dref, st); cs_free(L->iso, st); cs_free(L->ncheb, st); cs_free(L->cheb, st);
    delete L;
    return CS_OK;
}

static int32_t check_nu(int64_t nnu, const double* nu)
{
    CS_REQUIRE(nnu > 0 && nu, CS_ERR_ARG, "empty wavenumber vector");
    // surf!'s assert (line_shapes.jl:59) / checknu (gases.jl:90-95)
    for (int64_t i = 1; i < nnu; i++)
        CS_REQUIRE(nu[i] > nu[i - 1], CS_ERR_ARG, "wavenumber vectors must be sorted in ascending order (index %lld)", (long long)i);
    CS_REQUIRE(nu[0] >= 0, CS_ERR_ARG, "wavenumbers must be positive");
    return CS_OK;
}

extern "C" int32_t cs_xsec(cs_lines* L, int32_t shape, int64_t nnu, const double* nu, int64_t nlev, const double* T,
                           const double* P, const double* Pp, double cut, double* sigma)
{
    CS_REQUIRE(L && nu && T && P && Pp && sigma, CS_ERR_ARG, "null argument");
    CS_TRY(check_nu(nnu, nu));
    cs_ctx* ctx = L->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cs_reset_timers(ctx);
    CS_TRY(ctx->s_nu.reserve(sizeof(double) * (size_t)nnu));
    CS_TRY(ctx->s_sigma.reserve(sizeof(double) * (size_t)nnu * nlev));
    CS_CUDA(cudaMemcpyAsync(ctx->s_nu.p, nu, sizeof(double) * (size_t)nnu, cudaMemcpyHostToDevice, ctx->stream));
    CS_TRY(cs_lines_accumulate(L, shape, nnu, ctx->s_nu.as<double>(), nu, nlev, T, P, Pp, nullptr, cut,
                               ctx->s_sigma.as<double>(), 0));
    CS_CUDA(cudaMemcpyAsync(sigma, ctx->s_sigma.p, sizeof(double) * (size_t)nnu * nlev, cudaMemcpyDeviceToHost, ctx->stream));
    CS_CUDA(cudaStreamSynchronize(ctx->stream));
    return CS_OK;
}

extern "C" int32_t cs_count_evals(cs_lines* L, int64_t nnu, const double* nu, double cut, int64_t* evals)
{
    CS_REQUIRE(L && nu && evals, CS_ERR_ARG, "null argument");
    CS_TRY(check_nu(nnu, nu));
    const std::vector<double>& ln = L->h_nu;
    double numin = L->has_range ? L->rng_lo : nu[0], numax = L->has_range ? L->rng_hi : nu[nnu - 1];
    int64_t j0 = std::upper_bound(ln.begin(), ln.end(), numin - cut) - ln.begin();
    int64_t j1 = std::lower_bound(ln.begin(), ln.end(), numax + cut) - ln.begin();
    int64_t total = 0, lo = j0, hi = j0;
    for (int64_t i = 0; i < nnu; i++) {
        while (lo < j1 && (nu[i] - ln[lo]) > cut) lo++;
        if (hi < lo) hi = lo;
        while (hi < j1 && !(fabs(nu[i] - ln[hi]) > cut)) hi++;
        total += hi - lo;
    }
    *evals = total;
    return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// sigma workspace
__global__ void trapz_weights_kernel(const double* __restrict__ nu, int64_t nnu, double* __restrict__ w)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnu) return;
    const double dl = j > 0 ? nu[j] - nu[j - 1] : 0.0;
    const double dr = j + 1 < nnu ? nu[j + 1] - nu[j] : 0.0;
    w[j] = (dl + dr) / 2;
}

extern "C" int32_t cs_sigma_create(cs_ctx* ctx, int64_t nnu, const double* nu, int64_t nnode, cs_sigma** out)
{
    CS_REQUIRE(ctx && nu && out, CS_ERR_ARG, "null argument");
    *out = nullptr;
    CS_TRY(check_nu(nnu, nu));
    CS_REQUIRE(nnode > 0, CS_ERR_ARG, "no nodes");
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cs_sigma* s = new cs_sigma();
    s->ctx = ctx;
    s->nnu = nnu;
    s->nnode = nnode;
    s->own_nu = true;
    s->h_nu.assign(nu, nu + nnu);
    s->nu = nullptr;
    s->sig = nullptr;
    s->w = nullptr;
    int32_t rc = upload(&s->nu, nu, (size_t)nnu, ctx->stream);
    if (rc) { delete s; return rc; }
    cudaError_t e = cs_malloc((void**)&s->w, sizeof(double) * (size_t)nnu, ctx->stream);
    if (e == cudaSuccess) e = cs_malloc((void**)&s->sig, sizeof(double) * (size_t)nnu * nnode, ctx->stream);
    if (e != cudaSuccess) {
        cs_free(s->nu, ctx->stream);
        cs_free(s->w, ctx->stream);
        delete s;
        cs_set_error("cudaMalloc(sigma workspace %zu bytes): %s", sizeof(double) * (size_t)nnu * nnode, cudaGetErrorString(e));
        return CS_ERR_NOMEM;
    }
    // w_j = (dnu_{j-1} + dnu_j)/2: trapz(nu, y) = sum_j w_j y_j with every interval counted once (computed from the device copy)
    trapz_weights_kernel<<<(unsigned)((nnu + 255) / 256), 256, 0, ctx->stream>>>(s->nu, nnu, s->w);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    CS_CUDA(cudaMemsetAsync(s->sig, 0, sizeof(double) * (size_t)nnu * nnode, ctx->stream));
    // the caller's nu is only valid for the duration of the call: wait for its copy (everything else is stream-ordered)
    CS_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = s;
    return CS_OK;
}

extern "C" int32_t cs_sigma_zero(cs_sigma* s)
{
    CS_REQUIRE(s, CS_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(s->ctx->mtx);
    CS_CUDA(cudaSetDevice(s->ctx->device));
    CS_CUDA(cudaMemsetAsync(s->sig, 0, sizeof(double) * (size_t)s->nnu * s->nnode, s->ctx->stream));
    return CS_OK;
}

extern "C" int32_t cs_sigma_free(cs_sigma* s)
{
    if (!s) return CS_OK;
    cudaSetDevice(s->ctx->device);
    if (s->own_nu) cs_free(s->nu, s->ctx->stream);
    cs_free(s->w, s->ctx->stream);
    cs_free(s->sig, s->ctx->stream);
    delete s;
    return CS_OK;
}

extern "C" int32_t cs_sigma_read(cs_sigma* s, double* out)
{
    CS_REQUIRE(s && out, CS_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(s->ctx->mtx);
    CS_CUDA(cudaSetDevice(s->ctx->device));
    CS_CUDA(cudaMemcpyAsync(out, s->sig, sizeof(double) * (size_t)s->nnu * s->nnode, cudaMemcpyDeviceToHost, s->ctx->stream));
    CS_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return CS_OK;
}

extern "C" int32_t cs_sigma_add_lines(cs_sigma* s, cs_lines* L, int32_t shape, const double* T, const double* P,
                                      const double* C, double cut)
{
    CS_REQUIRE(s && L && T && P && C, CS_ERR_ARG, "null argument");
    CS_REQUIRE(s->ctx == L->ctx, CS_ERR_ARG, "workspace and lines live on different contexts");
    std::lock_guard<std::recursive_mutex> lk(s->ctx->mtx);
    CS_CUDA(cudaSetDevice(s->ctx->device));
    std::vector<double> Pp((size_t)s->nnode);
    for (int64_t k = 0; k < s->nnode; k++) {
        // bake's assert on concentrations (gases.jl:124)
        CS_REQUIRE(C[k] >= 0 && C[k] <= 1, CS_ERR_ARG, "gas molar concentrations must be in [0,1], not %g", C[k]);
        Pp[(size_t)k] = C[k] * P[k];
    }
    return cs_lines_accumulate(L, shape, s->nnu, s->nu, s->h_nu.data(), s->nnode, T, P, Pp.data(), C, cut, s->sig, 1);
}

__global__ void add_gray_kernel(double* sig, const double* nu, int64_t nnu, int64_t nnode, double value, double nu_cut)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnu) return;
    if (nu[i] <= nu_cut)
        for (int64_t k = 0; k < nnode; k++) sig[(size_t)k * nnu + i] += value;
}

extern "C" int32_t cs_sigma_add_gray(cs_sigma* s, double value, double nu_cut)
{
    CS_REQUIRE(s, CS_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(s->ctx->mtx);
    CS_CUDA(cudaSetDevice(s->ctx->device));
    add_gray_kernel<<<(unsigned)((s->nnu + 255) / 256), 256, 0, s->ctx->stream>>>(s->sig, s->nu, s->nnu, s->nnode, value, nu_cut);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(s->ctx);
    return CS_OK;
}

__global__ void add_array_kernel(double* dst, const double* src, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] += src[i];
}

extern "C" int32_t cs_sigma_add_host(cs_sigma* s, const double* h)
{
    CS_REQUIRE(s && h, CS_ERR_ARG, "null argument");
    cs_ctx* ctx = s->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    size_t n = (size_t)s->nnu * s->nnode;
    CS_TRY(ctx->s_sigma.reserve(sizeof(double) * n));
    CS_CUDA(cudaMemcpyAsync(ctx->s_sigma.p, h, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    add_array_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(s->sig, ctx->s_sigma.as<double>(), n);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    CS_CUDA(cudaStreamSynchronize(ctx->stream));
    return CS_OK;
}
