// cs_par.cu -- byte-parallel parser of HITRAN 160-column .par records (src/hitran/par.jl:127-152).
//
// HBM/PCIe-bound byte work: a warp stages its 32 consecutive records into shared memory with coalesced 16-byte loads,
// then every lane parses the ten numeric fields of one record.  Decimal -> binary64 is correctly rounded: the mantissa
// is accumulated exactly in a uint64; for |decimal exponent| <= 22 one IEEE multiply/divide of two exact doubles gives
// the correctly rounded result (Clinger's fast path); larger exponents (line intensities ~1e-30) go through a
// double-double division by two exact powers of ten, whose 106-bit intermediate is rounded once.
#include "cs_internal.cuh"

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
