/*
 * clearsky_b200.h -- C ABI of libclearsky_b200.so: the B200 (sm_100a) line-by-line radiative-transfer
 * engine that sits behind ClearSky.jl's Julia API for ONE hot path:
 *
 *   HITRAN line summation -> opacity-table build / (T, ln P) interpolation -> CIA -> layer optical depth
 *   (Gauss-Lobatto) -> Schwarzschild up/down sweep over Gauss-Legendre streams -> spectral reduction.
 *
 * The reference (markmbaum/ClearSky.jl) is pure Julia and has no FFI for this path; its extension points
 * are multiple dispatch and function arguments.  Each entry point below names the reference interface it
 * replaces (path:line under the reference repository) and is what a `ccall` from the Julia wrapper
 * (clearsky.jl_b200/julia/ClearSkyB200.jl) binds.
 *
 * Conventions
 *   - plain C, no exceptions: every function returns an int32 status (CS_OK == 0); on failure
 *     cs_last_error() returns a thread-local message.
 *   - all sizes int64_t, all real data double (Julia Float64), isotopologue ids int16_t (Julia Int16).
 *   - arrays are column-major exactly as Julia owns them; "[a][b]" below means b is the fastest index.
 *   - the caller owns every host array for the duration of the call only; the library copies what it
 *     keeps.  Device objects are owned by the library and released by cs_*_free.
 *   - calls are synchronous: outputs are valid when the function returns.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with CS_ERR_CUDA.
 */
#ifndef CLEARSKY_B200_H
#define CLEARSKY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS_MAXCHEB 16      /* padded length of one Qref/Q Chebyshev coefficient row */
#define CS_MAX_NODES 64    /* max temperature / pressure nodes per axis of an opacity table */
#define CS_MAX_STREAMS 16
#define CS_MAX_LOBATTO 16

enum {
    CS_OK = 0,
    CS_ERR_CUDA = 1,        /* CUDA runtime / launch failure (includes "no device") */
    CS_ERR_ARG = 2,         /* invalid argument (mirrors a reference @assert) */
    CS_ERR_NOMEM = 3,
    CS_ERR_DOMAIN = 4       /* T outside [25,1000] K, (T,P) outside a table domain, ... */
};

/* line shapes: doppler! / lorentz! / voigt! / PHCO2!  (src/absorption/line_shapes.jl:200,313,412,527) */
enum { CS_DOPPLER = 0, CS_LORENTZ = 1, CS_VOIGT = 2, CS_PHCO2 = 3 };

/* per-context kernel timers (milliseconds of the most recent call, CUDA events on the context stream) */
enum { CS_T_PREP = 0, CS_T_LINESUM = 1, CS_T_TABLE_FIT = 2, CS_T_TABLE_EVAL = 3, CS_T_CIA = 4,
       CS_T_RT = 5, CS_T_REDUCE = 6, CS_T_TOTAL = 7, CS_NTIMERS = 8 };

typedef struct cs_ctx cs_ctx;       /* one CUDA device + stream */
typedef struct cs_lines cs_lines;   /* SpectralLines resident on the device */
typedef struct cs_table cs_table;   /* Vector{OpacityTable} of one Gas: Chebyshev coefficients per wavenumber */
typedef struct cs_cia cs_cia;       /* CIATables resident on the device */
typedef struct cs_accel cs_accel;   /* AcceleratedAbsorber: ln sigma at fixed levels, per wavenumber */
typedef struct cs_sigma cs_sigma;   /* device workspace: Sigma(A, idx, T, P) at the quadrature nodes, all wavenumbers */

const char* cs_last_error(void);
int32_t cs_version(void);
int32_t cs_device_count(int32_t* n);

/* ---- context ---------------------------------------------------------------------------------- */
int32_t cs_ctx_create(int32_t device, cs_ctx** out);
/* same, but run on a caller-owned CUDA stream (cudaStream_t passed as void*), e.g. torch's current stream */
int32_t cs_ctx_create_on_stream(int32_t device, void* cuda_stream, cs_ctx** out);
int32_t cs_ctx_free(cs_ctx* ctx);
int32_t cs_ctx_synchronize(cs_ctx* ctx);
/* timers[CS_NTIMERS] of the last compute call; launches = kernels launched so far on this context */
int32_t cs_ctx_timers(cs_ctx* ctx, double* timers_ms);
/* same timers accumulated since the context was created (never reset): take differences around a region.  Timers are
 * CUDA-event pairs recorded on the context stream and read lazily, so compute calls never block the host for them;
 * both cs_ctx_timers and cs_ctx_timers_total wait for the spans still in flight. */
int32_t cs_ctx_timers_total(cs_ctx* ctx, double* timers_ms);
int32_t cs_ctx_launches(cs_ctx* ctx, int64_t* launches);
/* Far-wing treatment of the windowed line sum (Voigt, Lorentz, PHCO2; Doppler far lines contribute exactly 0).
 *   CS_FARFIELD_DIRECT (default): every (nu, line) pair inside the cut-off is evaluated, as surf! does
 *     (src/absorption/line_shapes.jl:53-87).
 *   CS_FARFIELD_EXPANSION: lines that are inside the cut-off for ALL points of a 128-point tile and provably in the
 *     far wing are summed through local expansions about the tile centre instead of pair by pair.
 *       Voigt / Lorentz: lines at least 4 half tile widths away, where the reference's Voigt equals
 *         S*gamma/(pi*(dnu^2+gamma^2)): 20-term Taylor expansion, truncation below 3e-11 of each line's own value;
 *         clusters of 32 lines at least 5 (h + cluster radius) away enter through 18 precomputed moments (< 5e-12).
 *       PHCO2: lines at least 30 cm^-1 from every point of the tile (chi classes 30-120 and >= 120, where chi factorises
 *         into a per-point and a per-line exponential): power-law expansions of K ge/dnu^2 (1 - eps + eps^2),
 *         eps = (chi*gamma/dnu)^2, used only at levels where the host bounds eps < 1e-4 and for tiles narrower than
 *         4 cm^-1; truncation below 1e-11.
 *     All contributions are positive, so these bounds also hold for the sum: two orders inside the 1e-9 parity
 *     tolerance.  Everything else (cut-off edges, near lines, line centres, chi-class borders) is evaluated exactly as
 *     in the direct mode.
 * The CS_FARFIELD environment variable ("direct" | "expansion") sets the default of new contexts. */
#define CS_FARFIELD_DIRECT 0
#define CS_FARFIELD_EXPANSION 1
int32_t cs_ctx_set_farfield(cs_ctx* ctx, int32_t mode);
int32_t cs_ctx_get_farfield(cs_ctx* ctx, int32_t* mode);
/* Floor applied to the vertical optical depth of every layer before the Schwarzschild sweep.  Default 1e-6 =
 * the reference's tau_min (src/core/discretized.jl:174); the Radau-equivalent wrappers (many thin layers) lower it
 * to 1e-9 so that the floor does not add opacity in spectral windows. */
int32_t cs_ctx_set_tau_floor(cs_ctx* ctx, double tau_min);
/* sustained FP64 FMA rate of this device, measured with a register-resident DFMA loop [FLOP/s] */
int32_t cs_fp64_peak(cs_ctx* ctx, int32_t iters, double* flops_per_s);

/* ---- lines: replaces the SpectralLines argument of shape!(...)  (src/hitran/par.jl:224-284) ----
 * nu must be ascending (SpectralLines sorts: par.jl:267).  iso[] = 1-based local isotopologue number
 * (sl.I); cheb[niso][CS_MAXCHEB], ncheb[niso] = MOLPARAM[sl.M].cheb / .ncheb (hitran/molparam.jl);
 * hascheb[i] == 0 for an isotopologue that is present in iso[] fails like scaleintensity's throw
 * (line_shapes.jl:115-119). */
int32_t cs_lines_upload(cs_ctx* ctx, int64_t n, const double* nu, const double* S, const double* gamma_a,
                        const double* gamma_s, const double* Epp, const double* na, const double* mu,
                        const int16_t* iso, int32_t niso, const int32_t* ncheb, const double* cheb,
                        const uint8_t* hascheb, cs_lines** out);
int32_t cs_lines_free(cs_lines* lines);
/* nu-sharded runs: tell a (sliced) line list the first and last point of the GLOBAL wavenumber grid.  The strict
 * includedlines prefilter (line_shapes.jl:18-22: nul > min(nu) - cut && nul < max(nu) + cut) is then applied to that grid,
 * as the reference would, instead of to the slice a call happens to see; at interior slice edges only the inclusive
 * per-point rule |nu - nul| <= cut (line_shapes.jl:10) decides, so a sharded run keeps exactly the unsharded run's lines. */
int32_t cs_lines_set_grid_range(cs_lines* lines, double numin, double numax);

/* vector forms of scaleintensity(sl, i, T) (line_shapes.jl:125-132), alpha-doppler(sl, i, T) (:146-148) and
 * gamma-lorentz(sl, i, T, P, Pp) (:259-261) for ALL lines at one (T, P, Pp): outputs [n] each (NULL = skip) */
int32_t cs_line_params(cs_lines* lines, double T, double P, double Pp, double* S_T, double* alpha, double* gamma);

/* S1: shape!(sigma, nu, sl, T, P, Pp, cut), batched over nlev (T,P,Pp) nodes
 * (line_shapes.jl:200-211, 313-324, 412-424, 527-540; surf! :53-87).  sigma is [nlev][nnu] and is
 * OVERWRITTEN like surf! does.  nu strictly ascending (assert :59); T in [25,1000] (assert :29). */
int32_t cs_xsec(cs_lines* lines, int32_t shape, int64_t nnu, const double* nu, int64_t nlev,
                const double* T, const double* P, const double* Pp, double dnu_cut, double* sigma);
/* exact number of surf! inner-loop iterations for one node: sum_i #{j : |nu_i - nul_j| <= cut}
 * restricted to the lines the strict prefilter keeps (line_shapes.jl:18-22,75-82) */
int32_t cs_count_evals(cs_lines* lines, int64_t nnu, const double* nu, double dnu_cut, int64_t* evals);

/* ---- S2: bake(sl, fC, shape!, cut, nu, Omega)  (src/absorption/gases.jl:97-145) ------------------
 * Tgrid[nT], Pgrid[nP] = Omega.T, Omega.P; C[nT][nP]... Julia C[i,j] = fC(T_i,P_j) at index i + nT*j.
 * Builds sigma[nnu,nT,nP], applies the zero-mixing repair (:131-142) and fits one Bichebyshev
 * interpolant of ln(sigma) in (T, ln P) per wavenumber (OpacityTable ctor, gases.jl:75-82).
 * keep_block != 0 keeps the raw sigma block on the device for cs_table_block. */
int32_t cs_bake(cs_lines* lines, int32_t shape, int64_t nnu, const double* nu, int32_t nT,
                const double* Tgrid, int32_t nP, const double* Pgrid, const double* C, double dnu_cut,
                int32_t keep_block, cs_table** out);
/* build the same table from a caller-supplied sigma block [nT*nP][nnu] (Julia sigma[nu,i,j]) */
int32_t cs_table_from_block(cs_ctx* ctx, int64_t nnu, int32_t nT, const double* Tgrid, int32_t nP,
                            const double* Pgrid, const double* sigma_block, cs_table** out);
/* rawsigma(g, T, P) for all wavenumbers at nlev nodes: sigma[nlev][nnu] = exp(Phi_nu(T, ln P))
 * (gases.jl:85,256,263).  (T,P) outside the table domain -> CS_ERR_DOMAIN (StrictBoundaries). */
int32_t cs_table_eval(cs_table* table, int64_t nlev, const double* T, const double* P, double* sigma);
/* copy the baked sigma block [nT*nP][nnu] to the host (to build stock OpacityTables in Julia) */
int32_t cs_table_block(cs_table* table, double* sigma_block);
int32_t cs_table_info(cs_table* table, int64_t* nnu, int32_t* nT, int32_t* nP, int64_t* nzeroed);
int32_t cs_table_free(cs_table* table);

/* ---- CIATables  (src/absorption/collision_induced_absorption.jl:145-276) -----------------------
 * ngrid bilinear grids (>= 2 temperatures): grid g has g_nnu[g] wavenumbers, g_nT[g] temperatures,
 * lnk[g] = log(k) as [nT][nnu] (Julia Z[i_nu, j_T], nu fastest), non-positive k already replaced by
 * floatmin (:205).  nsingle single-temperature tables: s_nu / s_lnk concatenated (log(0) = -Inf allowed). */
int32_t cs_cia_upload(cs_ctx* ctx, int32_t ngrid, const int64_t* g_nnu, const int64_t* g_nT,
                      const double* g_nu, const double* g_T, const double* g_lnk, int32_t nsingle,
                      const int64_t* s_n, const double* s_nu, const double* s_lnk, int32_t extrapolate,
                      int32_t singles, cs_cia** out);
int32_t cs_cia_free(cs_cia* cia);

/* ---- AcceleratedAbsorber  (src/absorption/absorbers.jl:114-209) --------------------------------- */
/* snapshot max(log(Sigma), log(floatmin)) of a sigma workspace whose nodes are the nlev levels P
 * (ascending) -> update!(A, T) (:173-200) */
int32_t cs_accel_from_sigma(cs_sigma* sig, const double* P, cs_accel** out);
/* the same object from host values: lnsig[nlev][nnu] = the phi_i[k] of a stock AcceleratedAbsorber (Julia matrix [nnu, nlev]),
 * already floored at log(floatmin) by update! (:193-195); P[nlev] ascending (:141-143) */
int32_t cs_accel_upload(cs_ctx* ctx, int64_t nnu, int64_t nlev, const double* P, const double* lnsig, cs_accel** out);
int32_t cs_accel_free(cs_accel* accel);

/* ---- sigma workspace: Sigma(A, idx, T, P) = sum of absorbers at the Lobatto nodes ---------------
 * (src/absorption/absorbers.jl:84-97, core/discretized.jl:76-81).  nnode = (np-1)*(nlobatto-1)+1 for
 * fluxes; node order = ascending pressure.  All cs_sigma_add_* ACCUMULATE into sig[nnode][nnu]. */
int32_t cs_sigma_create(cs_ctx* ctx, int64_t nnu, const double* nu, int64_t nnode, cs_sigma** out);
int32_t cs_sigma_zero(cs_sigma* sig);
int32_t cs_sigma_free(cs_sigma* sig);
/* Gas functor g(i,T,P) = fC(T,P) * Pi_i(T,P)  (gases.jl:278): += C[node] * table(T[node], P[node]) */
int32_t cs_sigma_add_table(cs_sigma* sig, cs_table* table, const double* T, const double* P, const double* C);
/* exact line-by-line gas at the nodes (no table): += C[node] * shape(nu, sl, T, P, C*P, cut) */
int32_t cs_sigma_add_lines(cs_sigma* sig, cs_lines* lines, int32_t shape, const double* T, const double* P,
                           const double* C, double dnu_cut);
/* CIA functor (collision_induced_absorption.jl:378-382,465): += cia(nu, x, T, P, P*C1, P*C2) */
int32_t cs_sigma_add_cia(cs_sigma* sig, cs_cia* cia, const double* T, const double* P, const double* C1,
                         const double* C2);
/* AcceleratedAbsorber Sigma = exp(phi_i(ln P)) (absorbers.jl:203): += */
int32_t cs_sigma_add_accel(cs_sigma* sig, cs_accel* accel, const double* P);
/* gray / semigray gases and pre-evaluated user functions sigma(nu,T,P): += host array [nnode][nnu] */
int32_t cs_sigma_add_host(cs_sigma* sig, const double* sigma_nodes);
/* += value for nu <= nu_cut (GrayGas: nu_cut = +Inf; SemiGrayGas gases.jl:386) */
int32_t cs_sigma_add_gray(cs_sigma* sig, double value, double nu_cut);
int32_t cs_sigma_read(cs_sigma* sig, double* sigma_nodes);

/* ---- S3: monochromaticfluxes!(M+, M-, tau, core::Discretized, ...) + integral-F! + Fnet ----------
 * (src/fluxes.jl:238-279,357-383; core/discretized.jl:136-177,249-326; core/shared.jl:125-137).
 * P[np] ascending (index 0 = TOA); mu[nlob][np-1] at the Lobatto nodes (discretized.jl:11-30);
 * Tlev[np] = fT(P[i]) (planckevaluations :46-58); wlob[nlob] Lobatto weights on [0,1];
 * m[nstream], W[nstream] = streamnodes (core/shared.jl:4-21); fS[nnu], fa[nnu] pre-evaluated.
 * Outputs: Fup/Fdn/Fnet[np] always; tau [np-1][nnu]... Julia tau[i,j] at i + (np-1)*j, Mup/Mdn
 * Julia M[i,j] at i + np*j -- NULL means "do not materialise".
 * nu_weights: NULL -> trapz over this workspace's own nu (util.jl:26-33); otherwise caller-supplied
 * per-point trapezoid weights (used by nu-sharded multi-GPU runs so that every interval is counted once). */
int32_t cs_fluxes(cs_sigma* sig, int64_t np, const double* P, int32_t nlob, const double* wlob,
                  const double* mu, const double* Tlev, double g, const double* fS, const double* fa,
                  double theta_s, int32_t nstream, const double* m, const double* W,
                  const double* nu_weights, double* tau, double* Mup, double* Mdn, double* Fup,
                  double* Fdn, double* Fnet);
/* same, but leaves Fup/Fdn (2*np doubles: Fup then Fdn) in DEVICE memory d_F for a following
 * collective (NCCL all-reduce by the caller) */
int32_t cs_fluxes_device(cs_sigma* sig, int64_t np, const double* P, int32_t nlob, const double* wlob,
                         const double* mu, const double* Tlev, double g, const double* fS,
                         const double* fa, double theta_s, int32_t nstream, const double* m,
                         const double* W, const double* nu_weights, double* d_F);
/* jacobian!(R, eps) (src/radiative_convective.jl:154-171) is np+1 flux solves that differ only in the temperature
 * profile.  When Sigma does not depend on T (AcceleratedAbsorber, absorbers.jl:203) they share every layer depth and
 * transmittance: cs_fluxes_batch runs nbatch profiles Tlev[nbatch][np] over ONE workspace, eight per launch, and returns
 * only the integrated fluxes F[nbatch][2*np] (F+ then F- per profile).  Same arguments as cs_fluxes otherwise. */
int32_t cs_fluxes_batch(cs_sigma* sig, int64_t np, const double* P, int32_t nlob, const double* wlob, const double* mu,
                        int64_t nbatch, const double* Tlev, double g, const double* fS, const double* fa, double theta_s,
                        int32_t nstream, const double* m, const double* W, const double* nu_weights, double* F);
/* opticaldepth(P::Vector, g, T, mu, theta, absorbers...; nlobatto) (src/fluxes.jl:68-97,
 * core/discretized.jl:92-134): total slant-path optical depth per wavenumber, no floor. */
int32_t cs_opticaldepth(cs_sigma* sig, int64_t np, const double* P, int32_t nlob, const double* wlob,
                        const double* mu, double g, double theta, double* tau_total);

/* ---- device-resident radiative-convective loop  (src/radiative_convective.jl:6-151) ------------------------------
 * heating!(R) (:109-144) = radiate! on the radiative levels Pr with the AcceleratedAbsorber, Fnet interpolated to the
 * cell edges Pe (AtmosphericProfile(Pr, Fnet), :123-124), cell heating rates H[i] = (g/cp_i) (R[i]-R[i+1]) / (Pe[i+1]-Pe[i])
 * and surface heating H[end] = R[end]/c_surf (:129-140); step!(R, dt) (:147-151) = T += dt*H.  The reference never calls
 * update!(A, T) inside heating!, and an AcceleratedAbsorber ignores T (absorbers.jl:203), so Sigma is the same at every
 * step: cs_rcm_create takes a sigma workspace already filled at the nodes of Pr ((nrad-1)*(nlob-1)+1 nodes), computes all
 * layer optical depths, stream transmittances and the stellar beam ONCE, and a step only evaluates Planck at the new level
 * temperatures (AtmosphericProfile(P, T) at Pr, :112), runs the stream recurrences on the stored transmittances, reduces
 * spectrally and updates the column -- three kernels captured in one CUDA graph; nothing crosses PCIe until cs_rcm_state.
 *   Pe[np] cell edges (ascending), P[np] cell centres + surface (:62-69), T0[np] initial cell temperatures,
 *   cp[np-1] = fcp(T_i, P_i) (held fixed), c_surf surface heat capacity, Pr[nrad] radiative levels (:71-85),
 *   mu[nlob][nrad-1] at the Lobatto nodes, the rest as in cs_fluxes.  The workspace is only read during the call. */
typedef struct cs_rcm cs_rcm;
int32_t cs_rcm_create(cs_sigma* sig, int64_t np, const double* Pe, const double* P, const double* T0, const double* cp,
                      double c_surf, int64_t nrad, const double* Pr, int32_t nlob, const double* wlob, const double* mu,
                      double g, const double* fS, const double* fa, double theta_s, int32_t nstream, const double* m,
                      const double* W, const double* nu_weights, cs_rcm** out);
int32_t cs_rcm_free(cs_rcm* rcm);
/* nsteps x step!(R, dt) on the device (graph replays), synchronised on return; dt = 0 evaluates heating! without moving T */
int32_t cs_rcm_step(cs_rcm* rcm, double dt, int64_t nsteps);
/* overwrite the cell temperatures (e.g. after a convective adjustment on the host) */
int32_t cs_rcm_set_temperature(cs_rcm* rcm, const double* T);
/* copy the column state to the host: T, H, R [np]; F+, F-, Fnet [nrad] of the last step; NULL = skip */
int32_t cs_rcm_state(cs_rcm* rcm, double* T, double* H, double* R, double* Fup, double* Fdn, double* Fnet);
int32_t cs_rcm_info(cs_rcm* rcm, int64_t* np, int64_t* nrad, int64_t* nnu);
int32_t cs_rcm_ctx(cs_rcm* rcm, cs_ctx** ctx);
/* nu-sharded columns (one cs_rcm per slice, global trapezoid weights): the two halves of a step around the caller's
 * all-reduce.  cs_rcm_enqueue_fluxes launches the flux kernels and leaves this slice's partial F+ then F- (2*nrad doubles)
 * in DEVICE memory d_F (NULL: the handle's own buffer, cs_rcm_flux_buffer); cs_rcm_enqueue_update consumes the summed
 * fluxes.  Both only enqueue work on the context stream (no host synchronisation), so a caller can capture
 * fluxes -> all-reduce -> update in its own CUDA graph. */
int32_t cs_rcm_enqueue_fluxes(cs_rcm* rcm, double* d_F);
int32_t cs_rcm_enqueue_update(cs_rcm* rcm, const double* d_F, double dt);
int32_t cs_rcm_flux_buffer(cs_rcm* rcm, double** d_F);
/* The same nu-sharded step with the collective FUSED into the step's last kernel (no NCCL call, two launches per step).
 * Every rank owns a mailbox in its own device memory (cs_rcm_peer_mailbox: a plain cudaMalloc block so that it can be
 * exported with cs_ipc_export and mapped by the other processes with cs_ipc_open; inside one process peer access does the
 * same).  cs_rcm_peer_connect hands a rank the nranks mailbox addresses as visible from ITS device (its own entry may be
 * NULL).  cs_rcm_enqueue_step_peer = flux kernel + one tail kernel that reduces this rank's partial sums in a fixed order,
 * stores the 2*nrad sums straight into every rank's mailbox over NVLink (peer stores, system fence, flag), spins on its own
 * mailbox until the flags of all ranks for this step have arrived, adds the nranks vectors in rank order (bit-identical on
 * every rank) and updates the column (radiative_convective.jl:123-149).  All ranks must enqueue the same number of steps.
 * cs_rcm_peer_status: steps completed and whether a flag ever failed to arrive within ~2 s of SM clocks (the step then
 * went on with incomplete sums instead of hanging the device).  Capturable in a CUDA graph like the two-call form. */
int32_t cs_rcm_peer_mailbox(cs_rcm* rcm, int32_t nranks, void** d_mailbox, int64_t* nbytes);
int32_t cs_rcm_peer_connect(cs_rcm* rcm, int32_t rank, int32_t nranks, void* const* d_mailboxes);
int32_t cs_rcm_enqueue_step_peer(cs_rcm* rcm, double dt);
int32_t cs_rcm_peer_status(cs_rcm* rcm, int64_t* steps, int32_t* timed_out);
/* CUDA IPC plumbing for the mailboxes: a 64-byte handle of a cudaMalloc'ed block / its mapping in another process */
int32_t cs_ipc_export(void* d_ptr, uint8_t* handle64);
int32_t cs_ipc_open(cs_ctx* ctx, const uint8_t* handle64, void** d_ptr);
int32_t cs_ipc_close(cs_ctx* ctx, void* d_ptr);

/* ---- HITRAN .par ingestion on the GPU (the step BEFORE the path; readpar's parse loop, src/hitran/par.jl:127-152) ----
 * text = the file's bytes, nrec fixed-width records of reclen bytes each (160 columns + line terminator), column map
 * of par.jl:131-140.  Outputs (host, nrec each, FILE order -- filtering/sorting stay with the caller exactly like
 * readpar's vector operations): M (molecule number), I (ISOINDEX of the isotopologue character, par.jl:6-13), nu, S, A,
 * gamma_a, gamma_s, Epp, na, delta_a.  Decimal -> binary64 conversion is correctly rounded (bit-identical to
 * parse(Float64, .)): exact one-operation path for |decimal exponent| <= 22, double-double otherwise; flags[i] != 0
 * marks a record with a malformed field (the caller re-parses or rejects it). */
int32_t cs_par_parse(cs_ctx* ctx, int64_t nbytes, const char* text, int32_t reclen, int64_t nrec, int16_t* M, int16_t* I,
                     double* nu, double* S, double* A, double* gamma_a, double* gamma_s, double* Epp, double* na,
                     double* delta_a, uint8_t* flags);

/* readpar(filename; numin, numax, Scut, I, maxlines) (src/hitran/par.jl:91-193) in ONE call: the text goes to the device through
 * pinned double-buffered staging, is parsed there (as cs_par_parse), filtered (nu in [numin, numax], S >= Scut, isotopologue
 * in Ilist[nI] given as ISOINDEX numbers; nI = 0 keeps all; :154-170), truncated to the maxlines strongest lines when
 * maxlines > 0 and nrec > maxlines (reverse(sortperm(S))[1:maxlines], compared against the PRE-filter count like the reference,
 * :178-185), and sorted by wavenumber with a stable sort on that order (:187-191).  Only the surviving records come back,
 * in output order: the ten numeric columns (capacity nrec each, *nout valid), index[k] = 0-based file record of output k
 * (NULL = skip; lets the caller gather the string columns), *nbad = records with a malformed numeric field (NULL = skip).
 * An empty selection fails with the reference's message (:172). */
int32_t cs_par_read(cs_ctx* ctx, int64_t nbytes, const char* text, int32_t reclen, int64_t nrec, double numin, double numax,
                    double Scut, int32_t nI, const int16_t* Ilist, int64_t maxlines, int16_t* M, int16_t* I, double* nu,
                    double* S, double* A, double* gamma_a, double* gamma_s, double* Epp, double* na, double* delta_a,
                    int64_t* index, int64_t* nout, int64_t* nbad);

/* ---- single-process multi-GPU: nu-sharded runs from ONE host process --------------------
 * A group owns one context per device and a single-node NCCL communicator (libnccl.so.2 is dlopen'ed on first use).
 * The host shards nu contiguously, drives each device's cs_* calls from its own thread (every call only blocks its
 * caller), lets each device leave its partial fluxes in its group buffer (cs_fluxes_device with the pointer from
 * cs_group_buffer and GLOBAL trapezoid weights), and then issues the path's only collective: one all-reduce. */
typedef struct cs_group cs_group;
int32_t cs_group_create(int32_t ndev, const int32_t* devices, cs_group** out);
int32_t cs_group_free(cs_group* grp);
int32_t cs_group_size(cs_group* grp, int32_t* ndev);
int32_t cs_group_ctx(cs_group* grp, int32_t i, cs_ctx** ctx);
/* device buffer of `count` doubles owned by the group on device i (grown on demand, contents preserved per call) */
int32_t cs_group_buffer(cs_group* grp, int32_t i, int64_t count, double** d_ptr);
/* in-place sum over all devices of the first `count` doubles of every group buffer (ncclAllReduce over NVLink) */
int32_t cs_group_allreduce_sum(cs_group* grp, int64_t count);
/* copy the first `count` doubles of device i's group buffer to the host */
int32_t cs_group_read(cs_group* grp, int32_t i, int64_t count, double* host);
/* nsteps radiative-convective steps of a nu-sharded column: rcm[i] lives on group member i; per step every device computes
 * its partial fluxes, ONE ncclAllReduce sums the 2*nrad values, every device applies the same column update.  Everything is
 * enqueued from the calling thread; the call returns when all devices have finished the last step. */
int32_t cs_group_rcm_step(cs_group* grp, cs_rcm* const* rcm, double dt, int64_t nsteps);

#ifdef __cplusplus
}
#endif
#endif
