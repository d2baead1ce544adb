// cs_par.cu -- byte-parallel parser of HITRAN 160-column .par records (src/hitran/par.jl:127-152).
//
// HBM/PCIe-bound byte work: a warp stages its 32 consecutive records into shared memory with coalesced 16-byte loads,
// then every lane parses the ten numeric fields of one record.  Decimal -> binary64 is correctly rounded: the mantissa
// is accumulated exactly in a uint64; for |decimal exponent| <= 22 one IEEE multiply/divide of two exact doubles gives
// the correctly rounded result (Clinger's fast path); larger exponents (line intensities ~1e-30) go through a
// double-double division by two exact powers of ten, whose 106-bit intermediate is rounded once.
#include "cs_internal.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cstring>
#include <limits>

namespace {

__constant__ double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                               1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

struct Field { int a, n; };   // 0-based start column, width  (par.jl:131-140)

// m * 10^e10, correctly rounded for the HITRAN field domain (see header comment)
__device__ double dec2double(unsigned long long m, int e10, bool& bad)
{
    if (m == 0) return 0.0;
    if (m >> 53) { bad = true; return (double)m; }        // more than 15-16 significant digits: not a HITRAN field
    double a = (double)(long long)m;                      // exact
    if (e10 == 0) return a;
    if (e10 > 0) {
        if (e10 <= 22) return a * P10[e10];
        if (e10 > 44) { bad = true; return a; }
        // double-double product a * 1e22 * 10^(e10-22)
        double p = a * P10[22], pe = fma(a, P10[22], -p);
        double b = P10[e10 - 22];
        double q = p * b, qe = fma(p, b, -q) + pe * b;
        return q + qe;
    }
    int k = -e10;
    if (k <= 22) return a / P10[k];
    if (k > 88) { bad = true; return 0.0; }
    // double-double quotient: divide by exact powers of ten (<= 1e22 each), carrying a ~106-bit (hi, lo) pair, and
    // round once at the end
    double h = a, l = 0.0;
    while (k > 0) {
        const double d = P10[k > 22 ? 22 : k];
        k -= 22;
        double q = h / d;
        double r = fma(-q, d, h) + l;      // exact remainder of the hi part + the lo part
        double ql = r / d;
        h = q + ql;                        // renormalise
        l = ql - (h - q);
    }
    return h + l;
}

__device__ double parse_field(const char* s, int n, bool& bad)
{
    int i = 0;
    while (i < n && s[i] == ' ') i++;
    if (i == n) { bad = true; return 0.0; }
    bool neg = false;
    if (s[i] == '-') { neg = true; i++; } else if (s[i] == '+') i++;
    unsigned long long m = 0;
    int nfrac = 0, ndig = 0;
    bool dot = false;
    for (; i < n; i++) {
        char c = s[i];
        if (c >= '0' && c <= '9') {
            if (ndig < 19) { m = m * 10ULL + (unsigned long long)(c - '0'); ndig += (m != 0 || ndig > 0) ? 1 : 0; if (dot) nfrac++; }
            else if (!dot) nfrac--;            // digits beyond 19: drop (never in HITRAN fields)
        } else if (c == '.' && !dot) dot = true;
        else break;
    }
    int ex = 0;
    if (i < n && (s[i] == 'E' || s[i] == 'e' || s[i] == 'D' || s[i] == 'd')) {
        i++;
        bool eneg = false;
        if (i < n && s[i] == '-') { eneg = true; i++; } else if (i < n && s[i] == '+') i++;
        int nd = 0;
        for (; i < n && s[i] >= '0' && s[i] <= '9'; i++) { ex = ex * 10 + (s[i] - '0'); nd++; }
        if (nd == 0) bad = true;
        if (eneg) ex = -ex;
    }
    while (i < n && s[i] == ' ') i++;
    if (i != n) bad = true;                    // trailing garbage
    double v = dec2double(m, ex - nfrac, bad);
    return neg ? -v : v;
}

struct ParArgs {
    const char* text;
    int64_t nrec;
    int reclen;
    int16_t *M, *I;
    double *nu, *S, *A, *ga, *gs, *Epp, *na, *da;
    uint8_t* flags;
};

constexpr int PAR_WARPS = 4;
constexpr int PAR_COLS = 67;      // the numeric fields end at column 67

__global__ void __launch_bounds__(PAR_WARPS * 32) par_parse_kernel(ParArgs a)
{
    extern __shared__ __align__(16) char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t rec0 = ((int64_t)blockIdx.x * PAR_WARPS + warp) * 32;
    if (rec0 >= a.nrec) return;
    const int nr = (int)min((int64_t)32, a.nrec - rec0);
    // coalesced staging of the warp's records (contiguous in the file) with aligned 16-byte loads; the records start
    // `mis` bytes into the staged block
    const size_t stride = ((size_t)32 * a.reclen + 16 + 15) & ~(size_t)15;
    char* buf = sm + (size_t)warp * stride;
    const char* src = a.text + rec0 * a.reclen;
    const int mis = (int)((uintptr_t)src & 15);
    const int nvec = (nr * a.reclen + mis + 15) >> 4;
    const uint4* src4 = reinterpret_cast<const uint4*>(src - mis);
    uint4* buf4 = reinterpret_cast<uint4*>(buf);
    for (int i = lane; i < nvec; i += 32) buf4[i] = src4[i];
    __syncwarp();
    if (lane >= nr) return;
    const char* r = buf + mis + (size_t)lane * a.reclen;
    bool bad = false;
    const int64_t o = rec0 + lane;
    // M: Int16 from columns 1-2 (may have a leading blank)
    int M = 0;
    for (int i = 0; i < 2; i++) {
        char c = r[i];
        if (c >= '0' && c <= '9') M = M * 10 + (c - '0'); else if (c != ' ') bad = true;
    }
    a.M[o] = (int16_t)M;
    // isotopologue character -> ISOINDEX (par.jl:6-13)
    char ic = r[2];
    int I = (ic >= '1' && ic <= '9') ? ic - '0' : (ic == '0' ? 10 : ((ic >= 'A' && ic <= 'Z') ? 11 + (ic - 'A') : -1));
    if (I < 0) bad = true;
    a.I[o] = (int16_t)I;
    a.nu[o] = parse_field(r + 3, 12, bad);
    a.S[o] = parse_field(r + 15, 10, bad);
    a.A[o] = parse_field(r + 25, 10, bad);
    a.ga[o] = parse_field(r + 35, 5, bad);
    a.gs[o] = parse_field(r + 40, 5, bad);
    a.Epp[o] = parse_field(r + 45, 10, bad);
    a.na[o] = parse_field(r + 55, 4, bad);
    a.da[o] = parse_field(r + 59, 8, bad);
    a.flags[o] = bad ? 1 : 0;
}

}  // namespace

extern "C" int32_t cs_par_parse(cs_ctx* ctx, int64_t nbytes, const char* text, int32_t reclen, int64_t nrec, int16_t* M,
                                int16_t* I, double* nu, double* S, double* A, double* ga, double* gs, double* Epp,
                                double* na, double* da, uint8_t* flags)
{
    CS_REQUIRE(ctx && text && M && I && nu && S && A && ga && gs && Epp && na && da && flags, CS_ERR_ARG, "null argument");
    CS_REQUIRE(reclen >= PAR_COLS && reclen <= 512, CS_ERR_ARG, "record length %d outside [%d, 512]", reclen, PAR_COLS);
    CS_REQUIRE(nrec > 0 && nbytes >= (nrec - 1) * (int64_t)reclen + PAR_COLS, CS_ERR_ARG, "buffer shorter than nrec records");
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // device layout: text | 8 double columns | 2 int16 columns | flags
    size_t tb = (((size_t)nrec * reclen + 255) / 256) * 256;
    size_t off_d = tb, off_i = off_d + sizeof(double) * 8 * (size_t)nrec, off_f = off_i + sizeof(int16_t) * 2 * (size_t)nrec;
    CS_TRY(ctx->s_sigma.reserve(off_f + (size_t)nrec + 256));
    char* base = ctx->s_sigma.as<char>();
    size_t ncopy = std::min<size_t>((size_t)nbytes, (size_t)nrec * reclen);
    if (ncopy < (size_t)nrec * reclen) CS_CUDA(cudaMemsetAsync(base, ' ', (size_t)nrec * reclen, st));   // last record without terminator
    CS_CUDA(cudaMemcpyAsync(base, text, ncopy, cudaMemcpyHostToDevice, st));
    ParArgs a;
    a.text = base; a.nrec = nrec; a.reclen = reclen;
    double* d = (double*)(base + off_d);
    a.nu = d; a.S = d + nrec; a.A = d + 2 * nrec; a.ga = d + 3 * nrec; a.gs = d + 4 * nrec; a.Epp = d + 5 * nrec;
    a.na = d + 6 * nrec; a.da = d + 7 * nrec;
    a.M = (int16_t*)(base + off_i); a.I = a.M + nrec;
    a.flags = (uint8_t*)(base + off_f);
    const int sp_parse = cs_span_begin(ctx, CS_T_TOTAL, true);     // parse kernel time (ms)
    int64_t nblk = (nrec + PAR_WARPS * 32 - 1) / (PAR_WARPS * 32);
    // the text buffer is padded to a multiple of 256 bytes, so the last aligned 16-byte load stays inside it
    size_t smem = (size_t)PAR_WARPS * ((((size_t)32 * reclen + 16 + 15) & ~(size_t)15));
    par_parse_kernel<<<(unsigned)nblk, PAR_WARPS * 32, smem, st>>>(a);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    cs_span_end(ctx, sp_parse);
    double* outs[8] = {nu, S, A, ga, gs, Epp, na, da};
    for (int k = 0; k < 8; k++)
        CS_CUDA(cudaMemcpyAsync(outs[k], d + (size_t)k * nrec, sizeof(double) * (size_t)nrec, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaMemcpyAsync(M, a.M, sizeof(int16_t) * (size_t)nrec, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaMemcpyAsync(I, a.I, sizeof(int16_t) * (size_t)nrec, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaMemcpyAsync(flags, a.flags, (size_t)nrec, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    cs_spans_collect(ctx, false);
    return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// readpar's vector operations on the device (src/hitran/par.jl:154-191): filters, the maxlines truncation and the final
// sort by wavenumber, so that only the surviving records cross PCIe, already in output order.
namespace {

// order-preserving map double -> uint64 (isless order: -0.0 before +0.0, as Julia's sortperm uses)
__device__ __forceinline__ unsigned long long ord_key(double x)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

__global__ void par_mask_kernel(int64_t n, const double* __restrict__ nu, const double* __restrict__ S,
                                const int16_t* __restrict__ I, double numin, double numax, double Scut, int nI,
                                const int16_t* __restrict__ Ilist, uint8_t* __restrict__ keep)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    bool k = (nu[j] >= numin) && (nu[j] <= numax) && (S[j] >= Scut);     // par.jl:155-157
    if (nI > 0) {                                                        // :158-169
        bool in = false;
        for (int q = 0; q < nI; q++) in |= (Ilist[q] == I[j]);
        k &= in;
    }
    keep[j] = k ? 1 : 0;
}

__global__ void par_keys_kernel(int64_t n, const int64_t* __restrict__ idx, const double* __restrict__ col,
                                unsigned long long* __restrict__ keys)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) keys[k] = ord_key(col[idx[k]]);
}

// reverse(sortperm(S))[1:m]: the last m entries of the ascending stable order, reversed (par.jl:181)
__global__ void par_reverse_tail_kernel(const int64_t* __restrict__ asc, int64_t nf, int64_t m, int64_t* __restrict__ q)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) q[k] = asc[nf - 1 - k];
}

struct ParGather {
    const int64_t* order;
    int64_t n;
    const double* in[8];
    double* out[8];
    const int16_t *Min, *Iin;
    int16_t *Mout, *Iout;
};
__global__ void par_gather_kernel(ParGather g)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.n) return;
    const int64_t j = g.order[k];
#pragma unroll
    for (int c = 0; c < 8; c++) g.out[c][k] = g.in[c][j];
    g.Mout[k] = g.Min[j];
    g.Iout[k] = g.Iin[j];
}

__global__ void par_count_bad_kernel(int64_t n, const uint8_t* __restrict__ flags, unsigned long long* nbad)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bad = (j < n && flags[j]) ? 1u : 0u;
    unsigned m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(nbad, (unsigned long long)__popc(m));
}

// host text -> device through two pinned staging buffers: the host memcpy of chunk c+1 overlaps the DMA of chunk c (a
// pageable cudaMemcpy of the whole file serialises the two)
int32_t staged_text_h2d(cs_ctx* ctx, char* dst, const char* src, size_t bytes)
{
    constexpr size_t CHUNK = (size_t)4 << 20;
    static thread_local void* pin[2] = {nullptr, nullptr};
    static thread_local cudaEvent_t done[2] = {nullptr, nullptr};
    for (int b = 0; b < 2; b++) {
        if (!pin[b]) CS_CUDA(cudaMallocHost(&pin[b], CHUNK));
        if (!done[b]) CS_CUDA(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
    }
    int b = 0;
    for (size_t off = 0; off < bytes; off += CHUNK, b ^= 1) {
        const size_t n = std::min(CHUNK, bytes - off);
        CS_CUDA(cudaEventSynchronize(done[b]));        // an event never recorded is complete
        memcpy(pin[b], src + off, n);
        CS_CUDA(cudaMemcpyAsync(dst + off, pin[b], n, cudaMemcpyHostToDevice, ctx->stream));
        CS_CUDA(cudaEventRecord(done[b], ctx->stream));
    }
    return CS_OK;
}

}  // namespace

extern "C" int32_t cs_par_read(cs_ctx* ctx, int64_t nbytes, const char* text, int32_t reclen, int64_t nrec, double numin,
                               double numax, double Scut, int32_t nI, const int16_t* Ilist, int64_t maxlines, int16_t* M,
                               int16_t* I, double* nu, double* S, double* A, double* ga, double* gs, double* Epp, double* na,
                               double* da, int64_t* index, int64_t* nout, int64_t* nbad)
{
    CS_REQUIRE(ctx && text && M && I && nu && S && A && ga && gs && Epp && na && da && nout, CS_ERR_ARG, "null argument");
    CS_REQUIRE(reclen >= PAR_COLS && reclen <= 512, CS_ERR_ARG, "record length %d outside [%d, 512]", reclen, PAR_COLS);
    CS_REQUIRE(nrec > 0 && nbytes >= (nrec - 1) * (int64_t)reclen + PAR_COLS, CS_ERR_ARG, "buffer shorter than nrec records");
    CS_REQUIRE(nI >= 0 && (nI == 0 || Ilist), CS_ERR_ARG, "bad isotopologue filter");
    CS_REQUIRE(nrec < ((int64_t)1 << 31), CS_ERR_ARG, "too many records for one call");
    *nout = 0;
    if (nbad) *nbad = 0;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    auto al = [](size_t b) { return ((b + 255) / 256) * 256; };
    const size_t n = (size_t)nrec;
    // device layout: text | 8 parsed columns | M, I | parse flags | keep | Ilist | counters | idx a/b | keys a/b | 8 output columns | M, I out
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
    const size_t o_text = take(n * reclen + 256), o_col = take(8 * n * sizeof(double)), o_mi = take(2 * n * sizeof(int16_t));
    const size_t o_flags = take(n), o_keep = take(n), o_il = take(sizeof(int16_t) * (size_t)std::max(nI, 1)), o_cnt = take(64);
    const size_t o_ia = take(n * 8), o_ib = take(n * 8), o_ka = take(n * 8), o_kb = take(n * 8);
    const size_t o_out = take(8 * n * sizeof(double)), o_mio = take(2 * n * sizeof(int16_t));
    size_t tmp_sel = 0, tmp_sort = 0;
    {
        thrust::counting_iterator<int64_t> it(0);
        cub::DeviceSelect::Flagged((void*)nullptr, tmp_sel, it, (const uint8_t*)nullptr, (int64_t*)nullptr, (int64_t*)nullptr, (int)nrec, st);
        cub::DeviceRadixSort::SortPairs((void*)nullptr, tmp_sort, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                        (const int64_t*)nullptr, (int64_t*)nullptr, (int)nrec, 0, 64, st);
    }
    const size_t tmp_bytes = std::max(tmp_sel, tmp_sort);
    const size_t o_tmp = take(tmp_bytes + 256);
    CS_TRY(ctx->s_sigma.reserve(off));
    char* base = ctx->s_sigma.as<char>();
    const size_t ncopy = std::min<size_t>((size_t)nbytes, n * reclen);
    if (ncopy < n * reclen) CS_CUDA(cudaMemsetAsync(base + o_text, ' ', n * reclen, st));
    CS_TRY(staged_text_h2d(ctx, base + o_text, text, ncopy));
    ParArgs a;
    a.text = base + o_text; a.nrec = nrec; a.reclen = reclen;
    double* d = (double*)(base + o_col);
    a.nu = d; a.S = d + n; a.A = d + 2 * n; a.ga = d + 3 * n; a.gs = d + 4 * n; a.Epp = d + 5 * n; a.na = d + 6 * n; a.da = d + 7 * n;
    a.M = (int16_t*)(base + o_mi); a.I = a.M + n;
    a.flags = (uint8_t*)(base + o_flags);
    const int sp = cs_span_begin(ctx, CS_T_TOTAL, true);
    const int64_t nblk = (nrec + PAR_WARPS * 32 - 1) / (PAR_WARPS * 32);
    const size_t smem = (size_t)PAR_WARPS * ((((size_t)32 * reclen + 16 + 15) & ~(size_t)15));
    par_parse_kernel<<<(unsigned)nblk, PAR_WARPS * 32, smem, st>>>(a);
    CS_CUDA(cudaGetLastError());
    // ---- filters (par.jl:154-170)
    int16_t* dIl = (int16_t*)(base + o_il);
    if (nI > 0) CS_TRY(cs_stage_h2d(ctx, dIl, Ilist, sizeof(int16_t) * (size_t)nI));
    uint8_t* keep = (uint8_t*)(base + o_keep);
    unsigned long long* cnt = (unsigned long long*)(base + o_cnt);      // [0] selected, [1] malformed
    CS_CUDA(cudaMemsetAsync(cnt, 0, 64, st));
    const unsigned g256 = (unsigned)((nrec + 255) / 256);
    par_mask_kernel<<<g256, 256, 0, st>>>(nrec, a.nu, a.S, a.I, numin, numax, Scut, nI, dIl, keep);
    par_count_bad_kernel<<<g256, 256, 0, st>>>(nrec, a.flags, cnt + 1);
    int64_t *ia = (int64_t*)(base + o_ia), *ib = (int64_t*)(base + o_ib);
    unsigned long long *ka = (unsigned long long*)(base + o_ka), *kb = (unsigned long long*)(base + o_kb);
    {
        thrust::counting_iterator<int64_t> it(0);
        size_t tb = tmp_bytes;
        CS_CUDA(cub::DeviceSelect::Flagged(base + o_tmp, tb, it, keep, ia, (int64_t*)cnt, (int)nrec, st));
    }
    unsigned long long hc[2] = {0, 0};
    CS_CUDA(cudaMemcpyAsync(hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    cs_count_launch(ctx, 4);
    int64_t nf = (int64_t)hc[0];
    if (nbad) *nbad = (int64_t)hc[1];
    CS_REQUIRE(nf > 0, CS_ERR_ARG, "par information has been filtered to nothing!");      // par.jl:172
    int64_t* cur = ia;
    int64_t* other = ib;
    // ---- strongest maxlines lines (par.jl:178-185; the comparison is against the PRE-filter count, as in the reference)
    if (maxlines > 0 && nrec > maxlines) {
        CS_REQUIRE(nf >= maxlines, CS_ERR_ARG,
                   "maxlines = %lld exceeds the %lld lines left after filtering (the reference raises a BoundsError here)",
                   (long long)maxlines, (long long)nf);
        const unsigned gf = (unsigned)((nf + 255) / 256);
        par_keys_kernel<<<gf, 256, 0, st>>>(nf, cur, a.S, ka);
        size_t tb = tmp_bytes;
        CS_CUDA(cub::DeviceRadixSort::SortPairs(base + o_tmp, tb, ka, kb, cur, other, (int)nf, 0, 64, st));     // stable, ascending
        par_reverse_tail_kernel<<<(unsigned)((maxlines + 255) / 256), 256, 0, st>>>(other, nf, maxlines, cur);
        CS_CUDA(cudaGetLastError());
        cs_count_launch(ctx, 3);
        nf = maxlines;
    }
    // ---- sort by wavenumber, stable on the current order (par.jl:187-191)
    {
        const unsigned gf = (unsigned)((nf + 255) / 256);
        par_keys_kernel<<<gf, 256, 0, st>>>(nf, cur, a.nu, ka);
        size_t tb = tmp_bytes;
        CS_CUDA(cub::DeviceRadixSort::SortPairs(base + o_tmp, tb, ka, kb, cur, other, (int)nf, 0, 64, st));
        std::swap(cur, other);
        ParGather g;
        g.order = cur; g.n = nf;
        double* od = (double*)(base + o_out);
        for (int c = 0; c < 8; c++) { g.in[c] = d + (size_t)c * n; g.out[c] = od + (size_t)c * n; }
        g.Min = a.M; g.Iin = a.I;
        g.Mout = (int16_t*)(base + o_mio); g.Iout = g.Mout + n;
        par_gather_kernel<<<gf, 256, 0, st>>>(g);
        CS_CUDA(cudaGetLastError());
        cs_count_launch(ctx, 3);
        cs_span_end(ctx, sp);
        double* outs[8] = {nu, S, A, ga, gs, Epp, na, da};
        const size_t bo = sizeof(double) * (size_t)nf;
        for (int c = 0; c < 8; c++) CS_CUDA(cudaMemcpyAsync(outs[c], od + (size_t)c * n, bo, cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaMemcpyAsync(M, g.Mout, sizeof(int16_t) * (size_t)nf, cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaMemcpyAsync(I, g.Iout, sizeof(int16_t) * (size_t)nf, cudaMemcpyDeviceToHost, st));
        if (index) CS_CUDA(cudaMemcpyAsync(index, cur, sizeof(int64_t) * (size_t)nf, cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaStreamSynchronize(st));
    }
    cs_spans_collect(ctx, false);
    *nout = nf;
    return CS_OK;
}
