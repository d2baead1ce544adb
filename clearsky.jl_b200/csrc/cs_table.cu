// cs_table.cu -- K3 opacity-table fit, K4 table evaluation, K4b/K4c accelerated absorber, K5 CIA, and cs_bake.
//
// Replaces bake (src/absorption/gases.jl:97-145), the OpacityTable constructor and call (:75-85), the Gas
// functor (:278), AcceleratedAbsorber update!/Sigma (src/absorption/absorbers.jl:173-203) and the CIA functor
// (src/absorption/collision_induced_absorption.jl:251-276, 295-303, 378-382, 465).
//
// Layout in HBM: every per-wavenumber quantity is stored "coefficient-major", [k][nu] with nu fastest, so that a
// warp of consecutive wavenumbers always reads one contiguous 256-byte segment per coefficient / node.
#include "cs_internal.cuh"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>

namespace {

constexpr double TINY = DBL_MIN;   // floatmin(Float64)

int32_t upload_d(double** dst, const double* src, size_t n, cudaStream_t st)
{
    *dst = nullptr;
    if (n == 0) return CS_OK;
    if (cs_malloc((void**)dst, sizeof(double) * n, st) != cudaSuccess) {
        cs_set_error("cudaMalloc(%zu bytes) failed", sizeof(double) * n);
        return CS_ERR_NOMEM;
    }
    CS_CUDA(cudaMemcpyAsync(*dst, src, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    return CS_OK;
}

// ---- K3a: zero-mixing repair (gases.jl:131-142) + log with the all-zero rule (gases.jl:75-81), in place
__global__ void __launch_bounds__(256) table_log_kernel(double* blk, int64_t nnu, int nk, unsigned long long* nzeroed)
{
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nnu) return;
    double mn = INFINITY, mx = -INFINITY;
    bool allz = true;
    for (int k = 0; k < nk; k++) {
        double s = blk[(size_t)k * nnu + v];
        mn = fmin(mn, s);
        mx = fmax(mx, s);
        if (!(s <= TINY)) allz = false;
    }
    if (mn == 0 && mx > 0) {
        allz = true;
        atomicAdd(nzeroed, 1ULL);
    }
    const double lt = log(TINY);
    for (int k = 0; k < nk; k++) {
        size_t o = (size_t)k * nnu + v;
        blk[o] = allz ? lt : log(blk[o]);
    }
}
// same decision, but writes sigma (zeroed where mixed) instead of logs: the block handed back to Julia
__global__ void __launch_bounds__(256) table_zero_kernel(double* blk, int64_t nnu, int nk)
{
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nnu) return;
    double mn = INFINITY, mx = -INFINITY;
    for (int k = 0; k < nk; k++) {
        double s = blk[(size_t)k * nnu + v];
        mn = fmin(mn, s);
        mx = fmax(mx, s);
    }
    if (mn == 0 && mx > 0)
        for (int k = 0; k < nk; k++) blk[(size_t)k * nnu + v] = 0.0;
}

// ---- K3b: one separable pass of the 2-D Chebyshev transform.  axis 0: along T (index i of k = i + nT*j),
// axis 1: along ln P (index j).  M is the n x n interpolation-coefficient matrix (row-major, M[out][in]).
__global__ void __launch_bounds__(128) cheb_pass_kernel(const double* in, double* out, int64_t nnu, int nT, int nP,
                                                        int axis, const double* M)
{
    extern __shared__ double sM[];
    const int n = axis == 0 ? nT : nP;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) sM[t] = M[t];
    __syncthreads();
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int other = blockIdx.y;   // j for axis 0, i for axis 1
    if (v >= nnu) return;
    double f[CS_MAX_NODES];
    const size_t base = axis == 0 ? (size_t)nT * other : (size_t)other;
    const size_t step = axis == 0 ? 1 : (size_t)nT;
    for (int q = 0; q < n; q++) f[q] = in[(base + step * q) * nnu + v];
    for (int o = 0; o < n; o++) {
        double acc = 0.0;
        for (int q = 0; q < n; q++) acc = fma(sM[o * n + q], f[q], acc);
        out[(base + step * o) * nnu + v] = acc;
    }
}

// ---- K4: table evaluation, out[l][nu] = exp( sum_j Tp_j(l) * sum_i coef[i + nT*j][nu] * Tt_i(l) ).
// FP64-FMA bound (2*nT*nP flop per (nu, level)); the coefficient block is read from HBM ONCE per block of
// up to 128 levels.  CTA = G warps that share 64 wavenumbers (a lane owns nu and nu+32: one shared-memory operand
// feeds two FMAs); warp g owns levels [g*LB, (g+1)*LB) of the CTA's level block, so the G warps read the same
// coefficient rows and hit L1.  The 1-D Chebyshev values Tt[i][l], Tp[j][l] of the block live in shared memory
// ((nT+nP)*G*LB doubles); coefficients are prefetched U rows ahead in registers.  The inner sum over i runs in
// tmp[], folded into acc[] once per j (nP extra FMAs per level: +1/nT).
// mode 0: out[l][nu] = exp(.)           (rawsigma, gases.jl:85,256)
// mode 1: out[l][nu] += C[l]*exp(.)     (Gas functor, gases.jl:278)
template <int LB, int U>
__global__ void __launch_bounds__(LB == 8 ? 512 : 352) table_eval_kernel(const double* __restrict__ coef, int64_t nnu, int nT, int nP,
                                                         const double* __restrict__ Tt, const double* __restrict__ Tp,
                                                         const double* __restrict__ C, int nlev, int lpb, double* out,
                                                         int mode)
{
    extern __shared__ __align__(16) double te_sm[];
    const int G = blockDim.x >> 5, LC = G * LB;
    double* sTt = te_sm;                 // [nT][LC]
    double* sTp = te_sm + nT * LC;       // [nP][LC]
    const int lb0 = blockIdx.y * lpb;    // first level of this CTA's block
    const int lend = min(lb0 + lpb, nlev);
    for (int t = threadIdx.x; t < nT * LC; t += blockDim.x) {
        int i = t / LC, l = lb0 + t % LC;
        sTt[t] = l < lend ? Tt[(size_t)l * nT + i] : 0.0;
    }
    for (int t = threadIdx.x; t < nP * LC; t += blockDim.x) {
        int j = t / LC, l = lb0 + t % LC;
        sTp[t] = l < lend ? Tp[(size_t)l * nP + j] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t v0 = (int64_t)blockIdx.x * 64 + lane, v1 = v0 + 32;
    const double* c0 = coef + (v0 < nnu ? v0 : nnu - 1);
    const double* c1 = coef + (v1 < nnu ? v1 : nnu - 1);
    const double* tt = sTt + warp * LB;
    const double* tp = sTp + warp * LB;
    const int nk = nT * nP;
    double acc0[LB], acc1[LB], tmp0[LB], tmp1[LB];
#pragma unroll
    for (int l = 0; l < LB; l++) acc0[l] = acc1[l] = tmp0[l] = tmp1[l] = 0.0;
    double a0[U], a1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        a0[u] = u < nk ? c0[(size_t)u * nnu] : 0.0;
        a1[u] = u < nk ? c1[(size_t)u * nnu] : 0.0;
    }
    int i = 0, j = 0;
    for (int k0 = 0; k0 < nk; k0 += U) {
        double b0[U], b1[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            int k = k0 + U + u;
            b0[u] = k < nk ? c0[(size_t)k * nnu] : 0.0;
            b1[u] = k < nk ? c1[(size_t)k * nnu] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (k0 + u < nk) {
                const double2* t2 = reinterpret_cast<const double2*>(tt + i * LC);
#pragma unroll
                for (int l = 0; l < LB; l += 2) {
                    double2 b = t2[l >> 1];
                    tmp0[l] = fma(a0[u], b.x, tmp0[l]);
                    tmp1[l] = fma(a1[u], b.x, tmp1[l]);
                    tmp0[l + 1] = fma(a0[u], b.y, tmp0[l + 1]);
                    tmp1[l + 1] = fma(a1[u], b.y, tmp1[l + 1]);
                }
                if (++i == nT) {
                    const double2* p2 = reinterpret_cast<const double2*>(tp + j * LC);
#pragma unroll
                    for (int l = 0; l < LB; l += 2) {
                        double2 b = p2[l >> 1];
                        acc0[l] = fma(tmp0[l], b.x, acc0[l]);
                        acc1[l] = fma(tmp1[l], b.x, acc1[l]);
                        acc0[l + 1] = fma(tmp0[l + 1], b.y, acc0[l + 1]);
                        acc1[l + 1] = fma(tmp1[l + 1], b.y, acc1[l + 1]);
                        tmp0[l] = tmp1[l] = tmp0[l + 1] = tmp1[l + 1] = 0.0;
                    }
                    i = 0;
                    j++;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) { a0[u] = b0[u]; a1[u] = b1[u]; }
    }
#pragma unroll
    for (int l = 0; l < LB; l++) {
        const int lev = lb0 + warp * LB + l;
        if (lev < lend) {
            if (v0 < nnu) {
                size_t o = (size_t)lev * nnu + v0;
                double s = exp(acc0[l]);
                out[o] = mode ? out[o] + C[lev] * s : s;
            }
            if (v1 < nnu) {
                size_t o = (size_t)lev * nnu + v1;
                double s = exp(acc1[l]);
                out[o] = mode ? out[o] + C[lev] * s : s;
            }
        }
    }
}

// ---- K4 as an FP64 tensor-core GEMM (the one dense contraction on the path):
//   lnsigma[nu][l] = sum_k coef[k][nu] * basis[k][l],   basis[k = i + nT*j][l] = T_i(xi_T(l)) * T_j(xi_P(l)),
// M = nnu, N = levels of one block (<= 128), K = nT*nP.  DMMA m8n8k4 (mma.sync, the only FP64 tensor-core form; tcgen05
// has no f64 kind) takes one A and one B element per lane for 256 FMAs, so operand delivery from shared memory costs
// 1/4 of the tensor time, whereas the FMA-pipe form above needs a broadcast shared-memory operand for every 2 FMAs
// and is bound by the shared-memory pipe at 29 % of the FP64 peak.
// CTA = 128 wavenumbers x <=128 levels: 16 consumer warps (4 along nu x 4 along levels, 32 x 32 each = 4 x 4 MMA tiles,
// 32 accumulator doubles per lane) + 1 producer warp that streams K in slabs of 16 rows (coefficient rows of 1 KB,
// basis rows of 1 KB) through a 4-stage shared-memory ring with cp.async.bulk and full/empty mbarriers.  Rows are
// padded to 132 doubles in shared memory so that the fragment loads (4 rows x 8 columns per half warp) are
// conflict-free.  Requires nnu even (16-byte aligned rows); otherwise table_eval_kernel runs.
constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16, GM_STAGES = 4, GM_LD = 132;
__global__ void __launch_bounds__(128) basis_kernel(const double* __restrict__ Tt, const double* __restrict__ Tp, int nT,
                                                    int nP, int nlev, int lpb, double* __restrict__ basis)
{
    // basis[by][k][GM_BN], zero beyond the levels of the block
    const int k = blockIdx.x, by = blockIdx.y, l = threadIdx.x;
    const int i = k % nT, j = k / nT;
    const int lev = by * lpb + l;
    const bool ok = l < lpb && lev < nlev;
    basis[((size_t)by * nT * nP + k) * GM_BN + l] = ok ? Tt[(size_t)lev * nT + i] * Tp[(size_t)lev * nP + j] : 0.0;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// The same kernel serves the two separable passes of the table fit (K3): there the "levels" are the output Chebyshev
// coefficients of one axis, blockIdx.y selects the other axis' index, rows are addressed through (row0, stride) pairs
// and the epilogue stores the plain sum (mode 2).
struct GemmRows {
    int64_t a_row0_step, a_row_stride;   // A row k of block y: (y*a_row0_step + k*a_row_stride)
    int64_t b_y_stride;                  // basis of block y starts at y*b_y_stride doubles
    int64_t o_row0_step, o_row_stride;   // output column l of block y goes to row (y*o_row0_step + l*o_row_stride)
};
__global__ void __launch_bounds__(32 * 17, 1) table_eval_mma_kernel(const double* coef, int64_t nnu, int nk,
                                                                 const double* __restrict__ basis,
                                                                 const double* __restrict__ C, int nlev, int lpb,
                                                                 double* out, int mode, GemmRows gr)
{
    extern __shared__ __align__(128) double gm_sm[];     // [GM_STAGES][2][GM_BK][GM_LD]
    __shared__ __align__(8) uint64_t full_bar[GM_STAGES], empty_bar[GM_STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lb0 = mode == 2 ? 0 : blockIdx.y * lpb;
    const int nl = min(lpb, nlev - lb0);                  // levels (output columns) of this block
    const int nactive = 4 * ((nl + 31) / 32);             // consumer warps that own at least one level
    const int64_t vbase = (int64_t)blockIdx.x * GM_BM;
    const int nv = (int)min((int64_t)GM_BM, nnu - vbase);
    const int nit = (nk + GM_BK - 1) / GM_BK;
    const double* bas = basis + (size_t)blockIdx.y * gr.b_y_stride;
    const double* arow = coef + (size_t)blockIdx.y * gr.a_row0_step * nnu + vbase;
    const size_t astride = (size_t)gr.a_row_stride * nnu;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GM_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], nactive); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == 16) {
        if (lane == 0) {
            const uint32_t abytes = (uint32_t)nv * 8u, bbytes = (uint32_t)(((nl + 1) & ~1) * 8);
            for (int it = 0; it < nit; it++) {
                const int s = it % GM_STAGES;
                if (it >= GM_STAGES) mbar_wait(&empty_bar[s], ((it / GM_STAGES) - 1) & 1);
                const int rows = min(GM_BK, nk - it * GM_BK);
                double* sA = gm_sm + (size_t)s * 2 * GM_BK * GM_LD;
                double* sB = sA + GM_BK * GM_LD;
                mbar_arrive_expect_tx(&full_bar[s], (abytes + bbytes) * (uint32_t)rows);
                for (int u = 0; u < rows; u++) {
                    const size_t k = (size_t)it * GM_BK + u;
                    tma_bulk_g2s(sA + u * GM_LD, arow + k * astride, abytes, &full_bar[s]);
                    tma_bulk_g2s(sB + u * GM_LD, bas + k * GM_BN, bbytes, &full_bar[s]);
                }
            }
        }
        return;
    }
    if (warp >= nactive) return;
    const int wm = warp & 3, wn = warp >> 2;
    const int ntc = min(4, (nl - 32 * wn + 7) / 8);       // level tiles of this warp that hold a real level
    const int fr = lane >> 2, fk = lane & 3;              // fragment row/column index, k index
    double acc[4][4][2];
#pragma unroll
    for (int t = 0; t < 4; t++)
#pragma unroll
        for (int u = 0; u < 4; u++) acc[t][u][0] = acc[t][u][1] = 0.0;
    for (int it = 0; it < nit; it++) {
        const int s = it % GM_STAGES;
        mbar_wait(&full_bar[s], (it / GM_STAGES) & 1);
        const double* sA = gm_sm + (size_t)s * 2 * GM_BK * GM_LD + 32 * wm + fr;
        const double* sB = gm_sm + (size_t)s * 2 * GM_BK * GM_LD + GM_BK * GM_LD + 32 * wn + fr;
        const int rows = min(GM_BK, nk - it * GM_BK);
#pragma unroll
        for (int kk = 0; kk < GM_BK / 4; kk++) {
            if (kk * 4 < rows) {
                const int kr = kk * 4 + fk;
                const bool kok = kr < rows;               // only the last slab can hold a partial group of four rows
                double a[4], b[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    double av = sA[kr * GM_LD + 8 * t], bv = sB[kr * GM_LD + 8 * t];
                    a[t] = kok ? av : 0.0;
                    b[t] = (kok && t < ntc) ? bv : 0.0;
                }
#pragma unroll
                for (int t = 0; t < 4; t++)
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (u < ntc) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    // epilogue: lane holds D[row = lane/4][col = 2*(lane%4) + {0,1}] of every 8 x 8 tile
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int64_t v = vbase + 32 * wm + 8 * t + fr;
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int l = 32 * wn + 8 * u + 2 * fk + c;
                if (u < ntc && l < nl && v < nnu) {
                    if (mode == 2) {
                        out[((size_t)blockIdx.y * gr.o_row0_step + (size_t)l * gr.o_row_stride) * nnu + v] = acc[t][u][c];
                    } else {
                        const int lev = lb0 + l;
                        size_t o = (size_t)lev * nnu + v;
                        double sg = exp(acc[t][u][c]);
                        out[o] = mode ? out[o] + C[lev] * sg : sg;
                    }
                }
            }
        }
    }
}

// ---- K3 fused: the whole table fit of 8 wavenumbers in ONE pass over HBM.
// The fit is lnsigma[i + nT j][nu] -> coef[i + nT j][nu] = sum_{i', j'} Mx[i][i'] My[j][j'] lnsigma[i' + nT j'][nu] (2-D Chebyshev
// transform, gases.jl:75-81 via BichebyshevInterpolator), with the zero-mixing repair (gases.jl:131-142) and the all-zero rule
// (:75-81) applied first.  As two separable GEMM passes over the block it costs 4 sweeps of HBM plus 3 for the log / repair
// pass (140 GB on configs[3], 61 ms); here a CTA owns 8 wavenumbers, keeps their whole [nT*nP] x 8 slab in shared memory
// (nP (nT|1) 64 B = 163 KB for 50 x 50), and does log, repair flags, T pass (in place, one j block per warp) and P pass
// (straight to global memory) on it: the block is read once and the coefficients written once (40 GB).  Both passes run on the
// FP64 tensor pipe (mma.sync.m8n8k4.f64: M = the 8 wavenumbers, N = output index, K = input index), A fragments from the slab
// and B fragments from padded copies of Mx^T / My^T.  A 64-bit fragment load is served per half warp (4 k rows x 4 columns), so
// the four rows must fall in four different bank groups: slab rows are 8 doubles wide and swizzled (columns 0-3 and 4-7 swap in
// rows with bit 1 set), j blocks are nTp = nT|1 rows apart (the P pass' four rows, nTp apart, are then distinct mod 4), and the
// matrix rows are padded to a pitch of 8 NT8 + 4 doubles.  2 (nT + nP) nT nP FMAs
// per wavenumber = 5e11 flop on configs[3]: the kernel is bound by the FP64 tensor pipe, not by HBM.
constexpr int FF_V = 8;            // wavenumbers per CTA (= M of the MMA)
constexpr int FF_MAXT = CS_MAX_NODES / 8;
struct FitArgs {
    const double* blk;      // [nk][nnu] sigma (is_log = 0) or ln sigma (is_log = 1)
    double* coef;           // [nk][nnu]
    const double* MxT;      // [4 KT][8 NT8 + 4] zero padded: MxT[q][o] = Mx[o][q]
    const double* MyT;      // [4 KP][8 NP8 + 4]
    unsigned long long* nzeroed;
    int64_t nnu;
    int nT, nP, nTp, is_log;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(800, 1) table_fit_fused_kernel(FitArgs a)
{
    extern __shared__ __align__(128) double fs[];
    __shared__ int sflag[2][FF_V];
    const int nT = a.nT, nP = a.nP, nTp = a.nTp;
    const int KT = (nT + 3) / 4, KP = (nP + 3) / 4, NT8 = (nT + 7) / 8, NP8 = (nP + 7) / 8;
    const int ldx = 8 * NT8 + 4, ldy = 8 * NP8 + 4;          // matrix row pitches (= 4 or 12 mod 16 doubles)
    double* slab = fs;                                       // [nP][nTp][8], row r holds column c at r*8 + (c ^ ((r & 2) << 1))
    double* sMx = slab + (size_t)nP * nTp * FF_V;            // [4 KT][ldx]
    double* sMy = sMx + (size_t)4 * KT * ldx;                // [4 KP][ldy]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const int part = lane & 3;                               // a lane always loads the same pair of wavenumbers
    const int64_t ntile = (a.nnu + FF_V - 1) / FF_V;
    const double lt = log(TINY);

    // The CTA is persistent.  A warp owns the T indices i = warp, warp + nwarp, ..: it loads the rows (i, all j) of a tile,
    // and as soon as its P pass is done with them it loads the same rows of the CTA's NEXT tile into the same place, so the
    // next tile's HBM reads overlap this tile's P pass and its stores.
    auto load_rows = [&](int i, int64_t tile) {
        const int64_t v0 = tile * FF_V;
        const int nv = (int)min((int64_t)FF_V, a.nnu - v0);
        for (int c = lane; c < 4 * nP; c += 32) {
            const int j = c >> 2, r = j * nTp + i;
            double* dst = slab + ((size_t)r * FF_V + ((2 * part) ^ ((r & 2) << 1)));
            if (2 * part < nv) cp_async16(dst, a.blk + ((size_t)i + (size_t)nT * j) * a.nnu + v0 + 2 * part);
            else { dst[0] = 0.0; dst[1] = 0.0; }
        }
    };
    if (tid < 2 * FF_V) (&sflag[0][0])[tid] = 0;
    int64_t tile = blockIdx.x;
    if (tile < ntile)
        for (int i = warp; i < nT; i += nwarp) load_rows(i, tile);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int c = tid; c < 4 * KT * ldx; c += nthr) sMx[c] = a.MxT[c];
    for (int c = tid; c < 4 * KP * ldy; c += nthr) sMy[c] = a.MyT[c];
    __syncthreads();

    for (int it = 0; tile < ntile; tile += gridDim.x, it ^= 1) {
        const int64_t v0 = tile * FF_V;
        const int nv = (int)min((int64_t)FF_V, a.nnu - v0);   // even (nnu is even), >= 2
        int* flg = sflag[it];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // ---- log + repair flags on the lane's own chunks (bit 0: a zero, bit 1: a positive value, bit 2: a value > floatmin)
        if (2 * part < nv) {
            int f0 = 0, f1 = 0;
            for (int i = warp; i < nT; i += nwarp) {
                for (int c = lane; c < 4 * nP; c += 32) {
                    const int j = c >> 2, r = j * nTp + i;
                    double2* q = reinterpret_cast<double2*>(slab + ((size_t)r * FF_V + ((2 * part) ^ ((r & 2) << 1))));
                    double2 x = *q;
                    if (a.is_log) {
                        f0 |= (x.x == -INFINITY ? 1 : 2) | (x.x > lt ? 4 : 0);
                        f1 |= (x.y == -INFINITY ? 1 : 2) | (x.y > lt ? 4 : 0);
                    } else {
                        f0 |= (x.x == 0.0 ? 1 : 0) | (x.x > 0.0 ? 2 : 0) | (x.x > TINY ? 4 : 0);
                        f1 |= (x.y == 0.0 ? 1 : 0) | (x.y > 0.0 ? 2 : 0) | (x.y > TINY ? 4 : 0);
                        x.x = log(x.x);
                        x.y = log(x.y);
                        *q = x;
                    }
                }
            }
            atomicOr(&flg[2 * part], f0);
            atomicOr(&flg[2 * part + 1], f1);
        }
        __syncthreads();
        if (tid < FF_V) {
            sflag[it ^ 1][tid] = 0;                          // the other buffer was last read in the previous tile's P pass
            if (tid < nv && (flg[tid] & 3) == 3) atomicAdd(a.nzeroed, 1ULL);     // min == 0 < max: zero-mixing (gases.jl:131-142)
        }
        // ---- T pass, in place: one j block per warp at a time (nobody else touches the block)
        for (int j = warp; j < nP; j += nwarp) {
            const int r0 = j * nTp;
            double acc[FF_MAXT][2];
#pragma unroll
            for (int t = 0; t < FF_MAXT; t++) acc[t][0] = acc[t][1] = 0.0;
            for (int ks = 0; ks < KT; ks++) {
                const int k = 4 * ks + fk;
                const int r = r0 + k;
                const double av = (k < nT) ? slab[(size_t)r * FF_V + (fr ^ ((r & 2) << 1))] : 0.0;
                const double* brow = sMx + (size_t)k * ldx + fr;
#pragma unroll
                for (int t = 0; t < FF_MAXT; t++)
                    if (t < NT8) dmma884(acc[t][0], acc[t][1], av, brow[8 * t]);
            }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < FF_MAXT; t++) {
                if (t < NT8) {
                    const int o = 8 * t + 2 * fk;
                    const int r = r0 + o;
                    if (o < nT) slab[(size_t)r * FF_V + (fr ^ ((r & 2) << 1))] = acc[t][0];
                    if (o + 1 < nT) slab[(size_t)(r + 1) * FF_V + (fr ^ (((r + 1) & 2) << 1))] = acc[t][1];
                }
            }
        }
        __syncthreads();
        // ---- P pass: unit = T index i; output straight to the coefficient block, with the repaired columns replaced by the
        // transform of the constant log(floatmin): c00 = log(floatmin), everything else 0
        const int fl = flg[fr];
        const bool repaired = ((fl & 3) == 3) || !(fl & 4);
        const bool vok = v0 + fr < a.nnu;
        const int64_t next = tile + gridDim.x;
        for (int i = warp; i < nT; i += nwarp) {
            double acc[FF_MAXT][2];
#pragma unroll
            for (int t = 0; t < FF_MAXT; t++) acc[t][0] = acc[t][1] = 0.0;
            for (int ks = 0; ks < KP; ks++) {
                const int k = 4 * ks + fk;
                const int r = k * nTp + i;
                const double av = (k < nP) ? slab[(size_t)r * FF_V + (fr ^ ((r & 2) << 1))] : 0.0;
                const double* brow = sMy + (size_t)k * ldy + fr;
#pragma unroll
                for (int t = 0; t < FF_MAXT; t++)
                    if (t < NP8) dmma884(acc[t][0], acc[t][1], av, brow[8 * t]);
            }
            __syncwarp();                                    // every lane has read the rows (i, .): they can be overwritten
            if (next < ntile) load_rows(i, next);
#pragma unroll
            for (int t = 0; t < FF_MAXT; t++) {
                if (t < NP8) {
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const int o = 8 * t + 2 * fk + c;
                        if (o < nP && vok) {
                            const double val = repaired ? ((i == 0 && o == 0) ? lt : 0.0) : acc[t][c];
                            a.coef[((size_t)i + (size_t)nT * o) * a.nnu + v0 + fr] = val;
                        }
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
}

// ---- accelerated absorber
__global__ void accel_snapshot_kernel(const double* sig, double* lnsig, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    const double lt = log(TINY);
    for (; i < n; i += stride) {
        double l = log(sig[i]);
        lnsig[i] = (l < lt) ? lt : l;     // absorbers.jl:193-195
    }
}
// node descriptors: cell index and ln P of each node
__global__ void accel_eval_kernel(const double* lnsig, int64_t nnu, const double* lnP, const int* cell, const double* q,
                                  int nnode, double* out)
{
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int m = blockIdx.y;
    if (v >= nnu || m >= nnode) return;
    int i = cell[m];
    double ya = lnsig[(size_t)i * nnu + v], yb = lnsig[(size_t)(i + 1) * nnu + v];
    double y = (q[m] - lnP[i]) * (yb - ya) / (lnP[i + 1] - lnP[i]) + ya;
    out[(size_t)m * nnu + v] += exp(y);   // absorbers.jl:203
}

// ---- K5: CIA
struct CiaNode {
    double T, P, P1, P2;
};
struct CiaGridNode {   // per (grid, node): how this node uses grid g
    int use;           // 0: skip, 1: interpolate at (nu, Tq)
    int jT;
    double Tq;         // T or the clamped T
};

__device__ __forceinline__ int findcell_dev(const double* x, int n, double q)
{
    if (q <= x[0]) return 0;
    if (q >= x[n - 1]) return n - 2;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int m = (lo + hi) >> 1;
        if (x[m] > q) hi = m; else lo = m;
    }
    return lo;
}

__global__ void __launch_bounds__(128) cia_kernel(const double* nu, int64_t nnu, int ngrid, const int64_t* desc,
                                                  const double* gnu, const double* gT, const double* glnk,
                                                  int nsingle, const int64_t* sdesc, const double* snu,
                                                  const double* slnk, int use_singles, const CiaNode* nodes,
                                                  const CiaGridNode* gn, int nnode, double* out)
{
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nnu) return;
    const double x = nu[v];
    for (int m = 0; m < nnode; m++) {
        double k = 0.0;
        for (int g = 0; g < ngrid; g++) {
            const int64_t nx = desc[5 * g + 0], ny = desc[5 * g + 1];
            const double* gx = gnu + desc[5 * g + 2];
            const double* gy = gT + desc[5 * g + 3];
            const double* Z = glnk + desc[5 * g + 4];
            CiaGridNode d = gn[(size_t)g * nnode + m];
            if (d.use && gx[0] <= x && x <= gx[nx - 1]) {
                int i = findcell_dev(gx, (int)nx, x);
                int j = d.jT;
                double xx = (x - gx[i]) / (gx[i + 1] - gx[i]);
                double yy = (d.Tq - gy[j]) / (gy[j + 1] - gy[j]);
                double z = (1 - xx) * (1 - yy) * Z[i + nx * j] + xx * (1 - yy) * Z[i + 1 + nx * j] +
                           xx * yy * Z[i + 1 + nx * (j + 1)] + (1 - xx) * yy * Z[i + nx * (j + 1)];
                k += exp(z);
            }
            (void)ny;
        }
        if (use_singles) {
            for (int s = 0; s < nsingle; s++) {
                const int64_t n = sdesc[2 * s + 0];
                const double* sx = snu + sdesc[2 * s + 1];
                const double* sy = slnk + sdesc[2 * s + 1];
                if (sx[0] <= x && x <= sx[n - 1]) {
                    int i = findcell_dev(sx, (int)n, x);
                    double y = (x - sx[i]) * (sy[i + 1] - sy[i]) / (sx[i + 1] - sx[i]) + sy[i];
                    k += exp(y);
                }
            }
        }
        CiaNode nd = nodes[m];
        // cia(k, T, Pa, P1, P2)  (collision_induced_absorption.jl:295-303)
        double rho1 = (nd.P1 / CS_ATM) * (CS_T0 / nd.T);
        double rho2 = (nd.P2 / CS_ATM) * (CS_T0 / nd.T);
        double rhoa = 1e-6 * nd.P / (CS_KB * nd.T);
        out[(size_t)m * nnu + v] += (k * CS_LO2) * rho1 * rho2 / rhoa;
    }
}

// interpolation-coefficient matrix of chebygrid(n) (ascending nodes): c_j = sum_k M[j][k] f_k
void cheb_matrix(int n, std::vector<double>& M)
{
    M.assign((size_t)n * n, 0.0);
    for (int j = 0; j < n; j++)
        for (int k = 0; k < n; k++) {
            double theta = CS_PI * (double)(n - 1 - k) / (double)(n - 1);
            double w = (k == 0 || k == n - 1) ? 0.5 : 1.0;
            double s = w * cos(j * theta) * (2.0 / (double)(n - 1));
            if (j == 0 || j == n - 1) s *= 0.5;
            M[(size_t)j * n + k] = s;
        }
}

int32_t check_grid(int32_t nT, const double* Tg, int32_t nP, const double* Pg)
{
    CS_REQUIRE(nT >= 2 && nT <= CS_MAX_NODES && nP >= 2 && nP <= CS_MAX_NODES, CS_ERR_ARG,
               "table grid must have 2..%d nodes per axis (got %d x %d)", CS_MAX_NODES, nT, nP);
    for (int i = 1; i < nT; i++) CS_REQUIRE(Tg[i] > Tg[i - 1], CS_ERR_ARG, "temperature grid must ascend");
    for (int i = 1; i < nP; i++) CS_REQUIRE(Pg[i] > Pg[i - 1], CS_ERR_ARG, "pressure grid must ascend");
    CS_REQUIRE(Pg[0] > 0, CS_ERR_ARG, "pressure range must be positive");
    // the interpolator requires Chebyshev (extrema) nodes; verify like BichebyshevInterpolator does
    for (int i = 0; i < nT; i++) {
        double xi = cos(CS_PI * (double)(nT - 1 - i) / (double)(nT - 1));
        double e = (xi + 1) * ((Tg[nT - 1] - Tg[0]) / 2) + Tg[0];
        CS_REQUIRE(fabs(e - Tg[i]) <= 1e-9 * fabs(Tg[nT - 1]), CS_ERR_ARG, "temperature nodes are not a Chebyshev grid");
    }
    for (int i = 0; i < nP; i++) {
        double a = log(Pg[0]), b = log(Pg[nP - 1]);
        double xi = cos(CS_PI * (double)(nP - 1 - i) / (double)(nP - 1));
        double e = (xi + 1) * ((b - a) / 2) + a;
        CS_REQUIRE(fabs(e - log(Pg[i])) <= 1e-9 * (fabs(a) + fabs(b) + 1), CS_ERR_ARG,
                   "log-pressure nodes are not a Chebyshev grid");
    }
    return CS_OK;
}

// does the single-sweep fused fit apply?  (even nnu for 16-byte row alignment, slab of 8 wavenumbers within shared memory)
size_t fused_fit_smem(int nT, int nP)
{
    const int KT = (nT + 3) / 4, KP = (nP + 3) / 4, NT8 = (nT + 7) / 8, NP8 = (nP + 7) / 8, nTp = nT | 1;
    return sizeof(double) * ((size_t)nP * nTp * FF_V + (size_t)4 * KT * (8 * NT8 + 4) + (size_t)4 * KP * (8 * NP8 + 4));
}
bool fit_is_fused(const cs_ctx* ctx, int64_t nnu, int nT, int nP)
{
    return nnu % 2 == 0 && !ctx->table_no_mma && !ctx->table_no_fused && fused_fit_smem(nT, nP) <= 227 * 1024 - 256;
}

// fit coefficients from a device block [nk][nnu] of sigma, or of ln sigma when block_is_log (destroyed either way)
int32_t fit_table(cs_ctx* ctx, cs_table* tb, double* d_block, bool block_is_log = false)
{
    cudaStream_t st = ctx->stream;
    const int nT = tb->nT, nP = tb->nP, nk = nT * nP;
    const int64_t nnu = tb->nnu;
    std::vector<double> Mx, My;
    cheb_matrix(nT, Mx);
    cheb_matrix(nP, My);
    // transposed, zero-padded copies [q][GM_BN] for the tensor-core passes (B operand of the GEMM)
    std::vector<double> MxT((size_t)nT * GM_BN, 0.0), MyT((size_t)nP * GM_BN, 0.0);
    for (int o = 0; o < nT; o++)
        for (int q = 0; q < nT; q++) MxT[(size_t)q * GM_BN + o] = Mx[(size_t)o * nT + q];
    for (int o = 0; o < nP; o++)
        for (int q = 0; q < nP; q++) MyT[(size_t)q * GM_BN + o] = My[(size_t)o * nP + q];
    auto al = [](size_t b) { return ((b + 255) / 256) * 256; };
    const size_t offy = al(Mx.size() * sizeof(double));
    const size_t offc = offy + al(My.size() * sizeof(double));
    const size_t offxt = offc + 256;
    const size_t offyt = offxt + al(MxT.size() * sizeof(double));
    CS_TRY(ctx->s_misc.reserve(offyt + al(MyT.size() * sizeof(double))));
    char* base = ctx->s_misc.as<char>();
    CS_CUDA(cudaMemcpyAsync(base, Mx.data(), Mx.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaMemcpyAsync(base + offy, My.data(), My.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaMemcpyAsync(base + offxt, MxT.data(), MxT.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaMemcpyAsync(base + offyt, MyT.data(), MyT.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaMemsetAsync(base + offc, 0, 8, st));
    const int sp_fit = cs_span_begin(ctx, CS_T_TABLE_FIT, false);
    // fused single-sweep fit when the slab of 8 wavenumbers fits in shared memory (any grid up to about 56 x 56)
    {
        const int KT = (nT + 3) / 4, KP = (nP + 3) / 4, NT8 = (nT + 7) / 8, NP8 = (nP + 7) / 8, nTp = nT | 1;
        const int ldx = 8 * NT8 + 4, ldy = 8 * NP8 + 4;
        const size_t nMx = (size_t)4 * KT * ldx, nMy = (size_t)4 * KP * ldy;
        const size_t smf = fused_fit_smem(nT, nP);
        if (fit_is_fused(ctx, nnu, nT, nP)) {
            std::vector<double> pad(nMx + nMy, 0.0);
            for (int o = 0; o < nT; o++)
                for (int q = 0; q < nT; q++) pad[(size_t)q * ldx + o] = Mx[(size_t)o * nT + q];
            for (int o = 0; o < nP; o++)
                for (int q = 0; q < nP; q++) pad[nMx + (size_t)q * ldy + o] = My[(size_t)o * nP + q];
            CS_TRY(ctx->s_w.reserve(pad.size() * sizeof(double)));
            CS_CUDA(cudaMemcpyAsync(ctx->s_w.p, pad.data(), pad.size() * sizeof(double), cudaMemcpyHostToDevice, st));
            FitArgs fa;
            fa.blk = d_block; fa.coef = tb->coef; fa.MxT = ctx->s_w.as<double>(); fa.MyT = ctx->s_w.as<double>() + nMx;
            fa.nzeroed = (unsigned long long*)(base + offc); fa.nnu = nnu; fa.nT = nT; fa.nP = nP; fa.nTp = nTp;
            fa.is_log = block_is_log ? 1 : 0;
            CS_CUDA(cudaFuncSetAttribute(table_fit_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf));
            int nsm = 0;
            CS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
            const int64_t ntile = (nnu + FF_V - 1) / FF_V;
            table_fit_fused_kernel<<<(unsigned)std::min<int64_t>(ntile, nsm), 800, smf, st>>>(fa);   // persistent: 1 CTA per SM
            CS_CUDA(cudaGetLastError());
            cs_count_launch(ctx, 1);
            cs_span_end(ctx, sp_fit);
            unsigned long long nzf = 0;
            CS_CUDA(cudaMemcpyAsync(&nzf, base + offc, 8, cudaMemcpyDeviceToHost, st));
            CS_CUDA(cudaStreamSynchronize(st));
            cs_spans_collect(ctx, false);
            tb->nzeroed = (int64_t)nzf;
            return CS_OK;
        }
    }
    CS_REQUIRE(!block_is_log, CS_ERR_ARG, "internal: the two-pass table fit expects a sigma block");
    unsigned nb = (unsigned)((nnu + 255) / 256);
    table_log_kernel<<<nb, 256, 0, st>>>(d_block, nnu, nk, (unsigned long long*)(base + offc));
    CS_CUDA(cudaGetLastError());
    // pass along T in place (a thread / CTA reads its whole line of nT rows before it writes any), then the pass along
    // ln P from the block into coef: the block is read twice and written once, coef written once
    if (nnu % 2 == 0 && !ctx->table_no_mma) {
        // tensor-core passes: per (128-wavenumber tile, index of the other axis) an [128 x n] = [128 x n] . [n x n] product
        const size_t sm2 = sizeof(double) * GM_STAGES * 2 * GM_BK * GM_LD;
        CS_CUDA(cudaFuncSetAttribute(table_eval_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        const unsigned gx = (unsigned)((nnu + GM_BM - 1) / GM_BM);
        GemmRows gT = {nT, 1, 0, nT, 1};          // block y = j: rows nT*j + q -> rows nT*j + o
        table_eval_mma_kernel<<<dim3(gx, (unsigned)nP), 32 * 17, sm2, st>>>(d_block, nnu, nT, (const double*)(base + offxt),
                                                                            nullptr, nT, nT, d_block, 2, gT);
        CS_CUDA(cudaGetLastError());
        GemmRows gP = {1, nT, 0, 1, nT};          // block y = i: rows i + nT*q -> rows i + nT*o
        table_eval_mma_kernel<<<dim3(gx, (unsigned)nT), 32 * 17, sm2, st>>>(d_block, nnu, nP, (const double*)(base + offyt),
                                                                            nullptr, nP, nP, tb->coef, 2, gP);
        CS_CUDA(cudaGetLastError());
    } else {
        dim3 gA((unsigned)((nnu + 127) / 128), (unsigned)nP);
        cheb_pass_kernel<<<gA, 128, sizeof(double) * nT * nT, st>>>(d_block, d_block, nnu, nT, nP, 0, (const double*)base);
        CS_CUDA(cudaGetLastError());
        dim3 gB((unsigned)((nnu + 127) / 128), (unsigned)nT);
        cheb_pass_kernel<<<gB, 128, sizeof(double) * nP * nP, st>>>(d_block, tb->coef, nnu, nT, nP, 1,
                                                                    (const double*)(base + offy));
        CS_CUDA(cudaGetLastError());
    }
    cs_count_launch(ctx, 3);
    cs_span_end(ctx, sp_fit);
    unsigned long long nz = 0;
    CS_CUDA(cudaMemcpyAsync(&nz, base + offc, 8, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    cs_spans_collect(ctx, false);
    tb->nzeroed = (int64_t)nz;
    return CS_OK;
}

cs_table* new_table(cs_ctx* ctx, int64_t nnu, int32_t nT, const double* Tg, int32_t nP, const double* Pg)
{
    cs_table* tb = new cs_table();
    tb->ctx = ctx;
    tb->nnu = nnu;
    tb->nT = nT;
    tb->nP = nP;
    tb->Ta = Tg[0];
    tb->Tb = Tg[nT - 1];
    tb->lnPa = log(Pg[0]);          // the table is built on log.(Omega.P) (gases.jl:78)
    tb->lnPb = log(Pg[nP - 1]);
    tb->coef = nullptr;
    tb->sigma_block = nullptr;
    tb->nzeroed = 0;
    return tb;
}

int32_t eval_table(cs_table* tb, int64_t nlev, const double* T, const double* P, const double* C, double* d_out, int mode)
{
    cs_ctx* ctx = tb->ctx;
    cudaStream_t st = ctx->stream;
    const int nT = tb->nT, nP = tb->nP;
    std::vector<double> basis((size_t)nlev * (nT + nP));   // Tt[l][i] then Tp[l][j]
    double* ct = basis.data();
    double* cp = basis.data() + (size_t)nlev * nT;
    for (int64_t l = 0; l < nlev; l++, ct += nT, cp += nP) {
        double lp = log(P[l]);
        // StrictBoundaries: out-of-domain coordinates are an error (gases.jl:85 via BichebyshevInterpolator)
        CS_REQUIRE(T[l] >= tb->Ta && T[l] <= tb->Tb, CS_ERR_DOMAIN,
                   "node %lld of %lld: temperature %g K (at %g Pa) outside the opacity-table domain [%g, %g] K", (long long)(l + 1),
                   (long long)nlev, T[l], P[l], tb->Ta, tb->Tb);
        CS_REQUIRE(lp >= tb->lnPa && lp <= tb->lnPb, CS_ERR_DOMAIN,
                   "node %lld of %lld: pressure %g Pa (at %g K) outside the opacity-table domain [%g, %g] Pa", (long long)(l + 1),
                   (long long)nlev, P[l], T[l], exp(tb->lnPa), exp(tb->lnPb));
        double xt = 2 * (T[l] - tb->Ta) / (tb->Tb - tb->Ta) - 1;
        double xp = 2 * (lp - tb->lnPa) / (tb->lnPb - tb->lnPa) - 1;
        ct[0] = 1; ct[1] = xt;
        for (int k = 2; k < nT; k++) ct[k] = 2 * xt * ct[k - 1] - ct[k - 2];
        cp[0] = 1; cp[1] = xp;
        for (int k = 2; k < nP; k++) cp[k] = 2 * xp * cp[k - 1] - cp[k - 2];
    }
    // level blocking: blocks of <= 128 levels (each re-reads the coefficients once)
    const int nk = nT * nP;
    const int nby = (int)((nlev + 127) / 128);
    const int lpb = (int)((nlev + nby - 1) / nby);
    // the tensor-core kernel needs 16-byte aligned coefficient rows: nnu even (coef itself is 256-byte aligned)
    const bool mma = (tb->nnu % 2 == 0) && !ctx->table_no_mma;
    const size_t offP = sizeof(double) * (size_t)nlev * nT;
    const size_t offC = ((basis.size() * sizeof(double) + 255) / 256) * 256;
    const size_t offB = offC + ((sizeof(double) * (size_t)nlev + 255) / 256) * 256;
    CS_TRY(ctx->s_misc.reserve(offB + (mma ? sizeof(double) * (size_t)nby * nk * GM_BN : 0)));
    char* base = ctx->s_misc.as<char>();
    CS_TRY(cs_stage_h2d(ctx, base, basis.data(), basis.size() * sizeof(double)));
    if (C) CS_TRY(cs_stage_h2d(ctx, base + offC, C, sizeof(double) * (size_t)nlev));
    const int sp_ev = cs_span_begin(ctx, CS_T_TABLE_EVAL, false);
    const double* dTt = (const double*)base;
    const double* dTp = (const double*)(base + offP);
    const double* dC = (const double*)(base + offC);
    if (mma) {
        double* dB = (double*)(base + offB);
        basis_kernel<<<dim3((unsigned)nk, (unsigned)nby), GM_BN, 0, st>>>(dTt, dTp, nT, nP, (int)nlev, lpb, dB);
        CS_CUDA(cudaGetLastError());
        const size_t sm2 = sizeof(double) * GM_STAGES * 2 * GM_BK * GM_LD;
        CS_CUDA(cudaFuncSetAttribute(table_eval_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        dim3 g2((unsigned)((tb->nnu + GM_BM - 1) / GM_BM), (unsigned)nby);
        GemmRows gr = {0, 1, (int64_t)nk * GM_BN, 0, 0};
        table_eval_mma_kernel<<<g2, 32 * 17, sm2, st>>>(tb->coef, tb->nnu, nk, dB, dC, (int)nlev, lpb, d_out, mode, gr);
        cs_count_launch(ctx);
    } else {
        // FMA-pipe kernel: within a level block G warps of LB levels each, LB from {8, 12} chosen for the least padding
        int LB = 8, best = 1 << 30;
        for (int cand : {8, 12}) {
            int pad = ((lpb + cand - 1) / cand) * cand;
            if (pad < best || (pad == best && cand > LB)) { best = pad; LB = cand; }
        }
        const int G = best / LB;
        const size_t smem = sizeof(double) * (size_t)(nT + nP) * best;
        dim3 grid((unsigned)((tb->nnu + 63) / 64), (unsigned)nby);
#define CS_TE_LAUNCH(LBV)                                                                                             \
    do {                                                                                                              \
        if (smem > 48 * 1024)                                                                                         \
            CS_CUDA(cudaFuncSetAttribute(table_eval_kernel<LBV, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                         (int)smem));                                                                 \
        table_eval_kernel<LBV, 4><<<grid, 32 * G, smem, st>>>(tb->coef, tb->nnu, nT, nP, dTt, dTp, dC, (int)nlev, lpb, \
                                                              d_out, mode);                                           \
    } while (0)
        if (LB == 8) CS_TE_LAUNCH(8);
        else CS_TE_LAUNCH(12);
    }
#undef CS_TE_LAUNCH
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    cs_span_end(ctx, sp_ev);
    cs_spans_collect(ctx, false);
    return CS_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" int32_t cs_table_from_block(cs_ctx* ctx, int64_t nnu, int32_t nT, const double* Tg, int32_t nP,
                                       const double* Pg, const double* block, cs_table** out)
{
    CS_REQUIRE(ctx && Tg && Pg && block && out, CS_ERR_ARG, "null argument");
    *out = nullptr;
    CS_REQUIRE(nnu > 0, CS_ERR_ARG, "no wavenumbers");
    CS_TRY(check_grid(nT, Tg, nP, Pg));
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cs_table* tb = new_table(ctx, nnu, nT, Tg, nP, Pg);
    size_t bytes = sizeof(double) * (size_t)nT * nP * nnu;
    if (cs_malloc((void**)&tb->coef, bytes, ctx->stream) != cudaSuccess) {
        delete tb;
        cs_set_error("cudaMalloc(table %zu bytes) failed", bytes);
        return CS_ERR_NOMEM;
    }
    int32_t rc = ctx->s_sigma.reserve(bytes);
    if (rc) { cs_table_free(tb); return rc; }
    CS_CUDA(cudaMemcpyAsync(ctx->s_sigma.p, block, bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = fit_table(ctx, tb, ctx->s_sigma.as<double>());
    if (rc) { cs_table_free(tb); return rc; }
    *out = tb;
    return CS_OK;
}

extern "C" int32_t cs_bake(cs_lines* L, int32_t shape, int64_t nnu, const double* nu, int32_t nT, const double* Tg,
                           int32_t nP, const double* Pg, const double* C, double cut, int32_t keep_block,
                           cs_table** out)
{
    CS_REQUIRE(L && nu && Tg && Pg && C && out, CS_ERR_ARG, "null argument");
    *out = nullptr;
    cs_ctx* ctx = L->ctx;
    CS_TRY(check_grid(nT, Tg, nP, Pg));
    CS_REQUIRE(nnu > 0, CS_ERR_ARG, "no wavenumbers");
    for (int64_t i = 1; i < nnu; i++)
        CS_REQUIRE(nu[i] > nu[i - 1], CS_ERR_ARG, "wavenumbers must be unique and in ascending order");   // gases.jl:92
    CS_REQUIRE(nu[0] >= 0, CS_ERR_ARG, "wavenumbers must be positive");                                   // gases.jl:93
    // AtmosphericDomain's Qref/Q range asserts (gases.jl:51-52)
    CS_REQUIRE(Tg[0] >= CS_TMIN && Tg[nT - 1] <= CS_TMAX, CS_ERR_DOMAIN, "temperature grid outside [%g,%g] K", CS_TMIN, CS_TMAX);
    const int nk = nT * nP;
    std::vector<double> T((size_t)nk), P((size_t)nk), Pp((size_t)nk);
    for (int j = 0; j < nP; j++)
        for (int i = 0; i < nT; i++) {
            size_t k = (size_t)i + (size_t)nT * j;
            // gases.jl:124
            CS_REQUIRE(C[k] >= 0 && C[k] <= 1, CS_ERR_ARG,
                       "gas molar concentrations must be in [0,1], not %g (encountered @ %g K, %g Pa)", C[k], Tg[i], Pg[j]);
            T[k] = Tg[i];
            P[k] = Pg[j];
            Pp[k] = C[k] * Pg[j];   // shape!(..., T, P, C*P, cut)  (gases.jl:126)
        }
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cs_reset_timers(ctx);
    cs_table* tb = new_table(ctx, nnu, nT, Tg, nP, Pg);
    size_t bytes = sizeof(double) * (size_t)nk * nnu;
    double* blk = nullptr;
    if (cs_malloc((void**)&tb->coef, bytes, ctx->stream) != cudaSuccess || cs_malloc((void**)&blk, bytes, ctx->stream) != cudaSuccess) {
        if (tb->coef) cs_free(tb->coef, ctx->stream);
        delete tb;
        cs_set_error("cudaMalloc(table 2 x %zu bytes) failed", bytes);
        return CS_ERR_NOMEM;
    }
    int32_t rc = ctx->s_nu.reserve(sizeof(double) * (size_t)nnu);
    bool want_log = false;
    if (!rc) {
        cudaMemcpyAsync(ctx->s_nu.p, nu, sizeof(double) * (size_t)nnu, cudaMemcpyHostToDevice, ctx->stream);
        // the fused fit reads ln sigma: let the line sum write it directly (one log per output point is free there),
        // unless the caller wants the sigma block kept
        want_log = !keep_block && fit_is_fused(ctx, nnu, nT, nP);
        rc = cs_lines_accumulate(L, shape, nnu, ctx->s_nu.as<double>(), nu, nk, T.data(), P.data(), Pp.data(), nullptr,
                                 cut, blk, want_log ? 2 : 0);
    }
    if (!rc && keep_block) {
        // keep sigma with the zero-mixing repair applied, exactly what bake hands to OpacityTable
        if (cs_malloc((void**)&tb->sigma_block, bytes, ctx->stream) != cudaSuccess) {
            cs_set_error("cudaMalloc(sigma block %zu bytes) failed", bytes);
            rc = CS_ERR_NOMEM;
        } else {
            cudaMemcpyAsync(tb->sigma_block, blk, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
            table_zero_kernel<<<(unsigned)((nnu + 255) / 256), 256, 0, ctx->stream>>>(tb->sigma_block, nnu, nk);
            cs_count_launch(ctx);
        }
    }
    if (!rc) rc = fit_table(ctx, tb, blk, want_log);
    cs_free(blk, ctx->stream);
    if (rc) { cs_table_free(tb); return rc; }
    *out = tb;
    return CS_OK;
}

extern "C" int32_t cs_table_eval(cs_table* tb, int64_t nlev, const double* T, const double* P, double* sigma)
{
    CS_REQUIRE(tb && T && P && sigma && nlev > 0, CS_ERR_ARG, "null argument");
    cs_ctx* ctx = tb->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cs_reset_timers(ctx);
    size_t bytes = sizeof(double) * (size_t)nlev * tb->nnu;
    CS_TRY(ctx->s_sigma.reserve(bytes));
    CS_TRY(eval_table(tb, nlev, T, P, nullptr, ctx->s_sigma.as<double>(), 0));
    CS_CUDA(cudaMemcpyAsync(sigma, ctx->s_sigma.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CS_CUDA(cudaStreamSynchronize(ctx->stream));
    return CS_OK;
}

extern "C" int32_t cs_sigma_add_table(cs_sigma* s, cs_table* tb, const double* T, const double* P, const double* C)
{
    CS_REQUIRE(s && tb && T && P && C, CS_ERR_ARG, "null argument");
    CS_REQUIRE(s->ctx == tb->ctx, CS_ERR_ARG, "workspace and table live on different contexts");
    CS_REQUIRE(s->nnu == tb->nnu, CS_ERR_ARG, "gases must have identical wavenumber vectors");   // absorbers.jl:227
    std::lock_guard<std::recursive_mutex> lk(s->ctx->mtx);
    CS_CUDA(cudaSetDevice(s->ctx->device));
    return eval_table(tb, s->nnode, T, P, C, s->sig, 1);
}

extern "C" int32_t cs_table_block(cs_table* tb, double* block)
{
    CS_REQUIRE(tb && block, CS_ERR_ARG, "null argument");
    CS_REQUIRE(tb->sigma_block, CS_ERR_ARG, "the sigma block was not kept (pass keep_block=1 to cs_bake)");
    std::lock_guard<std::recursive_mutex> lk(tb->ctx->mtx);
    CS_CUDA(cudaSetDevice(tb->ctx->device));
    CS_CUDA(cudaMemcpyAsync(block, tb->sigma_block, sizeof(double) * (size_t)tb->nT * tb->nP * tb->nnu,
                            cudaMemcpyDeviceToHost, tb->ctx->stream));
    CS_CUDA(cudaStreamSynchronize(tb->ctx->stream));
    return CS_OK;
}

extern "C" int32_t cs_table_info(cs_table* tb, int64_t* nnu, int32_t* nT, int32_t* nP, int64_t* nz)
{
    CS_REQUIRE(tb, CS_ERR_ARG, "null argument");
    if (nnu) *nnu = tb->nnu;
    if (nT) *nT = tb->nT;
    if (nP) *nP = tb->nP;
    if (nz) *nz = tb->nzeroed;
    return CS_OK;
}

extern "C" int32_t cs_table_free(cs_table* tb)
{
    if (!tb) return CS_OK;
    cudaSetDevice(tb->ctx->device);
    cs_free(tb->coef, tb->ctx->stream);
    cs_free(tb->sigma_block, tb->ctx->stream);
    delete tb;
    return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// accelerated absorber
extern "C" int32_t cs_accel_from_sigma(cs_sigma* s, const double* P, cs_accel** out)
{
    CS_REQUIRE(s && P && out, CS_ERR_ARG, "null argument");
    *out = nullptr;
    CS_REQUIRE(s->nnode >= 2, CS_ERR_ARG, "need at least two pressure levels");
    for (int64_t i = 1; i < s->nnode; i++)
        CS_REQUIRE(P[i] > P[i - 1], CS_ERR_ARG, "AcceleratedAbsorber pressure levels must ascend (absorbers.jl:141-143)");
    cs_ctx* ctx = s->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cs_accel* A = new cs_accel();
    A->ctx = ctx;
    A->nnu = s->nnu;
    A->nlev = s->nnode;
    A->h_lnP.resize((size_t)s->nnode);
    for (int64_t i = 0; i < s->nnode; i++) A->h_lnP[(size_t)i] = log(P[i]);
    size_t n = (size_t)s->nnu * s->nnode;
    if (cs_malloc((void**)&A->lnsig, sizeof(double) * n, ctx->stream) != cudaSuccess) {
        delete A;
        cs_set_error("cudaMalloc(accelerated absorber %zu bytes) failed", sizeof(double) * n);
        return CS_ERR_NOMEM;
    }
    accel_snapshot_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(s->sig, A->lnsig, n);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    CS_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = A;
    return CS_OK;
}

extern "C" int32_t cs_accel_upload(cs_ctx* ctx, int64_t nnu, int64_t nlev, const double* P, const double* lnsig, cs_accel** out)
{
    CS_REQUIRE(ctx && P && lnsig && out, CS_ERR_ARG, "null argument");
    *out = nullptr;
    CS_REQUIRE(nnu > 0 && nlev >= 2, CS_ERR_ARG, "need at least one wavenumber and two pressure levels");
    for (int64_t i = 1; i < nlev; i++)
        CS_REQUIRE(P[i] > P[i - 1], CS_ERR_ARG, "AcceleratedAbsorber pressure levels must ascend (absorbers.jl:141-143)");
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cs_accel* A = new cs_accel();
    A->ctx = ctx;
    A->nnu = nnu;
    A->nlev = nlev;
    A->h_lnP.resize((size_t)nlev);
    for (int64_t i = 0; i < nlev; i++) A->h_lnP[(size_t)i] = log(P[i]);
    int32_t rc = upload_d(&A->lnsig, lnsig, (size_t)nnu * nlev, ctx->stream);
    if (rc) { delete A; return rc; }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cs_set_error("cs_accel_upload: %s", cudaGetErrorString(e));
        cs_accel_free(A);
        return CS_ERR_CUDA;
    }
    *out = A;
    return CS_OK;
}

extern "C" int32_t cs_accel_free(cs_accel* A)
{
    if (!A) return CS_OK;
    cudaSetDevice(A->ctx->device);
    cs_free(A->lnsig, A->ctx->stream);
    delete A;
    return CS_OK;
}

extern "C" int32_t cs_sigma_add_accel(cs_sigma* s, cs_accel* A, const double* P)
{
    CS_REQUIRE(s && A && P, CS_ERR_ARG, "null argument");
    CS_REQUIRE(s->ctx == A->ctx && s->nnu == A->nnu, CS_ERR_ARG, "workspace and accelerated absorber do not match");
    cs_ctx* ctx = s->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t nn = s->nnode, nl = A->nlev;
    std::vector<int> cell((size_t)nn);
    std::vector<double> q((size_t)nn);
    const std::vector<double>& x = A->h_lnP;
    for (int64_t m = 0; m < nn; m++) {
        double v = log(P[m]);
        int64_t i;
        if (v <= x[0]) i = 0;
        else if (v >= x[(size_t)nl - 1]) i = nl - 2;
        else i = (std::upper_bound(x.begin(), x.end(), v) - x.begin()) - 1;
        cell[(size_t)m] = (int)i;
        q[(size_t)m] = v;
    }
    size_t offq = (((size_t)nl * sizeof(double) + 255) / 256) * 256;
    size_t offc = offq + (((size_t)nn * sizeof(double) + 255) / 256) * 256;
    CS_TRY(ctx->s_misc.reserve(offc + sizeof(int) * (size_t)nn));
    char* base = ctx->s_misc.as<char>();
    CS_TRY(cs_stage_h2d(ctx, base, x.data(), sizeof(double) * (size_t)nl));
    CS_TRY(cs_stage_h2d(ctx, base + offq, q.data(), sizeof(double) * (size_t)nn));
    CS_TRY(cs_stage_h2d(ctx, base + offc, cell.data(), sizeof(int) * (size_t)nn));
    dim3 grid((unsigned)((s->nnu + 255) / 256), (unsigned)nn);
    accel_eval_kernel<<<grid, 256, 0, st>>>(A->lnsig, s->nnu, (const double*)base, (const int*)(base + offc),
                                            (const double*)(base + offq), (int)nn, s->sig);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// CIA
extern "C" int32_t cs_cia_upload(cs_ctx* ctx, int32_t ngrid, const int64_t* g_nnu, const int64_t* g_nT,
                                 const double* g_nu, const double* g_T, const double* g_lnk, int32_t nsingle,
                                 const int64_t* s_n, const double* s_nu, const double* s_lnk, int32_t extrapolate,
                                 int32_t singles, cs_cia** out)
{
    CS_REQUIRE(ctx && out, CS_ERR_ARG, "null argument");
    *out = nullptr;
    CS_REQUIRE(ngrid >= 0 && nsingle >= 0 && ngrid + nsingle > 0, CS_ERR_ARG, "empty CIA tables");
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cs_cia* c = new cs_cia();
    c->ctx = ctx;
    c->ngrid = ngrid;
    c->nsingle = nsingle;
    c->extrapolate = extrapolate;
    c->singles = singles;
    c->d_nu = c->d_T = c->d_lnk = c->d_snu = c->d_slnk = nullptr;
    c->d_desc = nullptr;
    int64_t onu = 0, oT = 0, ok = 0;
    std::vector<int64_t> desc;
    for (int g = 0; g < ngrid; g++) {
        CS_REQUIRE(g_nnu[g] >= 2 && g_nT[g] >= 2, CS_ERR_ARG, "CIA grid %d needs >= 2 points per axis", g);
        c->g_nnu.push_back(g_nnu[g]); c->g_nT.push_back(g_nT[g]);
        c->g_off_nu.push_back(onu); c->g_off_T.push_back(oT); c->g_off_k.push_back(ok);
        desc.push_back(g_nnu[g]); desc.push_back(g_nT[g]); desc.push_back(onu); desc.push_back(oT); desc.push_back(ok);
        onu += g_nnu[g]; oT += g_nT[g]; ok += g_nnu[g] * g_nT[g];
    }
    c->h_T.assign(g_T, g_T + oT);
    int64_t os = 0;
    for (int s = 0; s < nsingle; s++) {
        CS_REQUIRE(s_n[s] >= 2, CS_ERR_ARG, "CIA single-temperature table %d needs >= 2 points", s);
        c->s_n.push_back(s_n[s]); c->s_off.push_back(os);
        desc.push_back(s_n[s]); desc.push_back(os);
        os += s_n[s];
    }
    cudaStream_t st = ctx->stream;
    int32_t rc = CS_OK;
    if ((ngrid && ((rc = upload_d(&c->d_nu, g_nu, (size_t)onu, st)) || (rc = upload_d(&c->d_T, g_T, (size_t)oT, st)) ||
                   (rc = upload_d(&c->d_lnk, g_lnk, (size_t)ok, st)))) ||
        (nsingle && ((rc = upload_d(&c->d_snu, s_nu, (size_t)os, st)) || (rc = upload_d(&c->d_slnk, s_lnk, (size_t)os, st))))) {
        cs_cia_free(c);
        return rc;
    }
    if (cs_malloc((void**)&c->d_desc, sizeof(int64_t) * desc.size(), st) != cudaSuccess) {
        cs_cia_free(c);
        cs_set_error("cudaMalloc(CIA descriptors) failed");
        return CS_ERR_NOMEM;
    }
    CS_CUDA(cudaMemcpyAsync(c->d_desc, desc.data(), sizeof(int64_t) * desc.size(), cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaStreamSynchronize(st));
    *out = c;
    return CS_OK;
}

extern "C" int32_t cs_cia_free(cs_cia* c)
{
    if (!c) return CS_OK;
    cudaSetDevice(c->ctx->device);
    cudaStream_t st = c->ctx->stream;
    cs_free(c->d_nu, st); cs_free(c->d_T, st); cs_free(c->d_lnk, st); cs_free(c->d_snu, st); cs_free(c->d_slnk, st); cs_free(c->d_desc, st);
    delete c;
    return CS_OK;
}

extern "C" int32_t cs_sigma_add_cia(cs_sigma* s, cs_cia* c, const double* T, const double* P, const double* C1,
                                    const double* C2)
{
    CS_REQUIRE(s && c && T && P && C1 && C2, CS_ERR_ARG, "null argument");
    CS_REQUIRE(s->ctx == c->ctx, CS_ERR_ARG, "workspace and CIA tables live on different contexts");
    cs_ctx* ctx = s->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mtx);
    CS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t nn = s->nnode;
    std::vector<CiaNode> nodes((size_t)nn);
    std::vector<CiaGridNode> gn((size_t)nn * (size_t)std::max(c->ngrid, 1));
    for (int64_t m = 0; m < nn; m++) {
        nodes[(size_t)m] = {T[m], P[m], P[m] * C1[m], P[m] * C2[m]};   // collision_induced_absorption.jl:379-380
        for (int g = 0; g < c->ngrid; g++) {
            const double* y = c->h_T.data() + c->g_off_T[(size_t)g];
            int64_t ny = c->g_nT[(size_t)g];
            CiaGridNode d = {0, 0, T[m]};
            if (y[0] <= T[m] && T[m] <= y[ny - 1]) d.use = 1;                       // :257
            else if (c->extrapolate) { d.use = 1; d.Tq = T[m] > y[ny - 1] ? y[ny - 1] : y[0]; }   // :260-262
            if (d.use) {
                int64_t j;
                if (d.Tq <= y[0]) j = 0;
                else if (d.Tq >= y[ny - 1]) j = ny - 2;
                else j = (std::upper_bound(y, y + ny, d.Tq) - y) - 1;
                d.jT = (int)j;
            }
            gn[(size_t)g * nn + m] = d;
        }
    }
    size_t offg = ((nodes.size() * sizeof(CiaNode) + 255) / 256) * 256;
    CS_TRY(ctx->s_misc.reserve(offg + gn.size() * sizeof(CiaGridNode)));
    char* base = ctx->s_misc.as<char>();
    CS_TRY(cs_stage_h2d(ctx, base, nodes.data(), nodes.size() * sizeof(CiaNode)));
    CS_TRY(cs_stage_h2d(ctx, base + offg, gn.data(), gn.size() * sizeof(CiaGridNode)));
    const int sp_cia = cs_span_begin(ctx, CS_T_CIA, false);
    cia_kernel<<<(unsigned)((s->nnu + 127) / 128), 128, 0, st>>>(
        s->nu, s->nnu, c->ngrid, c->d_desc, c->d_nu, c->d_T, c->d_lnk, c->nsingle, c->d_desc + 5 * c->ngrid, c->d_snu,
        c->d_slnk, c->singles, (const CiaNode*)base, (const CiaGridNode*)(base + offg), (int)nn, s->sig);
    CS_CUDA(cudaGetLastError());
    cs_count_launch(ctx);
    cs_span_end(ctx, sp_cia);
    cs_spans_collect(ctx, false);
    return CS_OK;
}
