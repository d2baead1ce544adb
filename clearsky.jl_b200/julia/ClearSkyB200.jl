# ClearSkyB200.jl -- thin Julia wrapper that puts libclearsky_b200.so behind ClearSky.jl's own API.
#
# STATUS: NOT EXECUTED in the build image (no Julia runtime there).  Two things stand in for execution:
#   * tests/test_julia_bindings.py parses EVERY `ccall` in this file and checks symbol, arity, return type and each
#     argument type against include/clearsky_b200.h, and that every exported cs_* symbol is bound here;
#   * the Python twin clearsky.jl_b200/clearsky_b200/ issues the same calls in the same order and is what tests/ and
#     bench.py drive on the GPU.
# Treat the file as untested Julia until it has run once under a Julia with ClearSky.jl installed.
#
# Seams (SURVEY.md section 8b) -- no reference file is edited:
#   S1  shape!(σ, ν, sl, T, P, Pₚ, Δνcut)            -> voigt_b200!, lorentz_b200!, doppler_b200!, PHCO2_b200!
#   S2  Gas(sl, fC, ν, Ω, shape!, Δνcut) / bake       -> method of ClearSky.Gas for the B200 shapes: ONE cs_bake call, returns a
#                                                       STOCK `Gas` whose Π elements are handles into the device table
#   S3  monochromaticfluxes!(M⁺, M⁻, τ, core, P, g, T, μ, 𝒻S, 𝒻a, absorbers...) / radiate!(F, core, ...)
#                                                    -> methods for core::B200Discretized <: AbstractNumericalCore that accept
#                                                       what the reference passes: ONE UnifiedAbsorber or AcceleratedAbsorber
#                                                       (src/fluxes.jl:334,377; src/radiative_convective.jl:113) or loose absorbers
module ClearSkyB200

using ClearSky
using ClearSky: SpectralLines, AtmosphericDomain, AbstractGas, Gas, GrayGas, SemiGrayGas, OpacityTable,
                AbstractNumericalCore, AbstractAbsorber, UnifiedAbsorber, AcceleratedAbsorber, CIATables, CIA,
                MOLPARAM, ISOINDEX, FluxPack, RCM, AtmosphericProfile, LinearInterpolator, NoBoundaries,
                formprofile, formprofiles, lobattonodes, streamnodes, lobattoevaluations, unifyabsorbers,
                checkazimuth, checkstreams, checkpressures, checkν, concentration

export B200Discretized, voigt_b200!, lorentz_b200!, doppler_b200!, PHCO2_b200!, B200Shape
export farfield!, taufloor!, fluxes_b200, opticaldepth_b200, outgoing_b200, opticaldepth_between_b200
export AcceleratedAbsorber_b200, update_b200!, refresh!, stockgas, jacobian_b200!, step_b200!, DeviceRCM
export readpar_b200, lineparams_b200, countevals_b200, DeviceGroup, sharded_fluxes_b200, fp64peak, timers, launches

const LIB = get(ENV, "CLEARSKY_B200_LIB", joinpath(@__DIR__, "..", "lib", "libclearsky_b200.so"))
const MAXCHEB = 16
const NTIMERS = 8
const DOPPLER, LORENTZ, VOIGT, PHCO2_ID = Int32(0), Int32(1), Int32(2), Int32(3)
const Handle = Ptr{Cvoid}
const F64 = Float64
const NULLF = Ptr{Float64}(C_NULL)

# =============================================================================================================
# raw bindings: one Julia function per entry point of include/clearsky_b200.h, same name, same argument order
# =============================================================================================================
module Lib
import ..LIB, ..Handle
cs_last_error() = ccall((:cs_last_error, LIB), Cstring, ())
cs_version() = ccall((:cs_version, LIB), Int32, ())
cs_device_count(n) = ccall((:cs_device_count, LIB), Int32, (Ref{Int32},), n)
cs_ctx_create(device, out) = ccall((:cs_ctx_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, out)
cs_ctx_create_on_stream(device, stream, out) =
    ccall((:cs_ctx_create_on_stream, LIB), Int32, (Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, stream, out)
cs_ctx_free(ctx) = ccall((:cs_ctx_free, LIB), Int32, (Ptr{Cvoid},), ctx)
cs_ctx_synchronize(ctx) = ccall((:cs_ctx_synchronize, LIB), Int32, (Ptr{Cvoid},), ctx)
cs_ctx_timers(ctx, t) = ccall((:cs_ctx_timers, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx, t)
cs_ctx_timers_total(ctx, t) = ccall((:cs_ctx_timers_total, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx, t)
cs_ctx_launches(ctx, n) = ccall((:cs_ctx_launches, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), ctx, n)
cs_ctx_set_farfield(ctx, mode) = ccall((:cs_ctx_set_farfield, LIB), Int32, (Ptr{Cvoid}, Int32), ctx, mode)
cs_ctx_get_farfield(ctx, mode) = ccall((:cs_ctx_get_farfield, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}), ctx, mode)
cs_ctx_set_tau_floor(ctx, τmin) = ccall((:cs_ctx_set_tau_floor, LIB), Int32, (Ptr{Cvoid}, Float64), ctx, τmin)
cs_fp64_peak(ctx, iters, flops) = ccall((:cs_fp64_peak, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Float64}), ctx, iters, flops)
cs_lines_upload(ctx, n, ν, S, γa, γs, Epp, na, μ, iso, niso, ncheb, cheb, hascheb, out) =
    ccall((:cs_lines_upload, LIB), Int32,
          (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
           Ptr{Float64}, Ptr{Int16}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{UInt8}, Ref{Ptr{Cvoid}}),
          ctx, n, ν, S, γa, γs, Epp, na, μ, iso, niso, ncheb, cheb, hascheb, out)
cs_lines_free(lines) = ccall((:cs_lines_free, LIB), Int32, (Ptr{Cvoid},), lines)
cs_lines_set_grid_range(lines, νmin, νmax) =
    ccall((:cs_lines_set_grid_range, LIB), Int32, (Ptr{Cvoid}, Float64, Float64), lines, νmin, νmax)
cs_line_params(lines, T, P, Pₚ, S, α, γ) =
    ccall((:cs_line_params, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
          lines, T, P, Pₚ, S, α, γ)
cs_xsec(lines, shape, nν, ν, nlev, T, P, Pₚ, Δνcut, σ) =
    ccall((:cs_xsec, LIB), Int32,
          (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}),
          lines, shape, nν, ν, nlev, T, P, Pₚ, Δνcut, σ)
cs_count_evals(lines, nν, ν, Δνcut, evals) =
    ccall((:cs_count_evals, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64, Ref{Int64}), lines, nν, ν, Δνcut, evals)
cs_bake(lines, shape, nν, ν, nT, Tg, nP, Pg, C, Δνcut, keep, out) =
    ccall((:cs_bake, LIB), Int32,
          (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Int32,
           Ref{Ptr{Cvoid}}),
          lines, shape, nν, ν, nT, Tg, nP, Pg, C, Δνcut, keep, out)
cs_table_from_block(ctx, nν, nT, Tg, nP, Pg, block, out) =
    ccall((:cs_table_from_block, LIB), Int32,
          (Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
          ctx, nν, nT, Tg, nP, Pg, block, out)
cs_table_eval(table, nlev, T, P, σ) =
    ccall((:cs_table_eval, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), table, nlev, T, P, σ)
cs_table_block(table, block) = ccall((:cs_table_block, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), table, block)
cs_table_info(table, nν, nT, nP, nz) =
    ccall((:cs_table_info, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int32}, Ref{Int32}, Ref{Int64}), table, nν, nT, nP, nz)
cs_table_free(table) = ccall((:cs_table_free, LIB), Int32, (Ptr{Cvoid},), table)
cs_cia_upload(ctx, ngrid, gnν, gnT, gν, gT, glnk, nsingle, sn, sν, slnk, extrapolate, singles, out) =
    ccall((:cs_cia_upload, LIB), Int32,
          (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Int64},
           Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ref{Ptr{Cvoid}}),
          ctx, ngrid, gnν, gnT, gν, gT, glnk, nsingle, sn, sν, slnk, extrapolate, singles, out)
cs_cia_free(cia) = ccall((:cs_cia_free, LIB), Int32, (Ptr{Cvoid},), cia)
cs_accel_from_sigma(sig, P, out) =
    ccall((:cs_accel_from_sigma, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{Ptr{Cvoid}}), sig, P, out)
cs_accel_upload(ctx, nν, nlev, P, lnσ, out) =
    ccall((:cs_accel_upload, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
          ctx, nν, nlev, P, lnσ, out)
cs_accel_free(accel) = ccall((:cs_accel_free, LIB), Int32, (Ptr{Cvoid},), accel)
cs_sigma_create(ctx, nν, ν, nnode, out) =
    ccall((:cs_sigma_create, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ref{Ptr{Cvoid}}), ctx, nν, ν, nnode, out)
cs_sigma_zero(sig) = ccall((:cs_sigma_zero, LIB), Int32, (Ptr{Cvoid},), sig)
cs_sigma_free(sig) = ccall((:cs_sigma_free, LIB), Int32, (Ptr{Cvoid},), sig)
cs_sigma_add_table(sig, table, T, P, C) =
    ccall((:cs_sigma_add_table, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), sig, table, T, P, C)
cs_sigma_add_lines(sig, lines, shape, T, P, C, Δνcut) =
    ccall((:cs_sigma_add_lines, LIB), Int32,
          (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64), sig, lines, shape, T, P, C, Δνcut)
cs_sigma_add_cia(sig, cia, T, P, C₁, C₂) =
    ccall((:cs_sigma_add_cia, LIB), Int32,
          (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), sig, cia, T, P, C₁, C₂)
cs_sigma_add_accel(sig, accel, P) =
    ccall((:cs_sigma_add_accel, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), sig, accel, P)
cs_sigma_add_host(sig, σ) = ccall((:cs_sigma_add_host, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), sig, σ)
cs_sigma_add_gray(sig, value, νcut) = ccall((:cs_sigma_add_gray, LIB), Int32, (Ptr{Cvoid}, Float64, Float64), sig, value, νcut)
cs_sigma_read(sig, σ) = ccall((:cs_sigma_read, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), sig, σ)
cs_fluxes(sig, np, P, nlob, 𝓌, μ, Tlev, g, fS, fa, θₛ, nstream, 𝓂, 𝒲, νw, τ, M⁺, M⁻, F⁺, F⁻, Fnet) =
    ccall((:cs_fluxes, LIB), Int32,
          (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64},
           Ptr{Float64}, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
           Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
          sig, np, P, nlob, 𝓌, μ, Tlev, g, fS, fa, θₛ, nstream, 𝓂, 𝒲, νw, τ, M⁺, M⁻, F⁺, F⁻, Fnet)
cs_fluxes_device(sig, np, P, nlob, 𝓌, μ, Tlev, g, fS, fa, θₛ, nstream, 𝓂, 𝒲, νw, dF) =
    ccall((:cs_fluxes_device, LIB), Int32,
          (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64},
           Ptr{Float64}, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
          sig, np, P, nlob, 𝓌, μ, Tlev, g, fS, fa, θₛ, nstream, 𝓂, 𝒲, νw, dF)
cs_fluxes_batch(sig, np, P, nlob, 𝓌, μ, nbatch, Tlev, g, fS, fa, θₛ, nstream, 𝓂, 𝒲, νw, F) =
    ccall((:cs_fluxes_batch, LIB), Int32,
          (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Float64, Ptr{Float64},
           Ptr{Float64}, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
          sig, np, P, nlob, 𝓌, μ, nbatch, Tlev, g, fS, fa, θₛ, nstream, 𝓂, 𝒲, νw, F)
cs_opticaldepth(sig, np, P, nlob, 𝓌, μ, g, θ, τ) =
    ccall((:cs_opticaldepth, LIB), Int32,
          (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Float64}),
          sig, np, P, nlob, 𝓌, μ, g, θ, τ)
cs_rcm_create(sig, np, Pₑ, P, T₀, cₚ, cₛ, nrad, Pᵣ, nlob, 𝓌, μ, g, fS, fa, θₛ, nstream, 𝓂, 𝒲, νw, out) =
    ccall((:cs_rcm_create, LIB), Int32,
          (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Int64, Ptr{Float64}, Int32,
           Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Int32, Ptr{Float64}, Ptr{Float64},
           Ptr{Float64}, Ref{Ptr{Cvoid}}),
          sig, np, Pₑ, P, T₀, cₚ, cₛ, nrad, Pᵣ, nlob, 𝓌, μ, g, fS, fa, θₛ, nstream, 𝓂, 𝒲, νw, out)
cs_rcm_free(rcm) = ccall((:cs_rcm_free, LIB), Int32, (Ptr{Cvoid},), rcm)
cs_rcm_step(rcm, Δt, nsteps) = ccall((:cs_rcm_step, LIB), Int32, (Ptr{Cvoid}, Float64, Int64), rcm, Δt, nsteps)
cs_rcm_set_temperature(rcm, T) = ccall((:cs_rcm_set_temperature, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), rcm, T)
cs_rcm_state(rcm, T, H, R, F⁺, F⁻, Fnet) =
    ccall((:cs_rcm_state, LIB), Int32,
          (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), rcm, T, H, R, F⁺, F⁻, Fnet)
cs_rcm_info(rcm, np, nrad, nν) =
    ccall((:cs_rcm_info, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), rcm, np, nrad, nν)
cs_rcm_ctx(rcm, ctx) = ccall((:cs_rcm_ctx, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), rcm, ctx)
cs_rcm_enqueue_fluxes(rcm, dF) = ccall((:cs_rcm_enqueue_fluxes, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), rcm, dF)
cs_rcm_enqueue_update(rcm, dF, Δt) = ccall((:cs_rcm_enqueue_update, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Float64), rcm, dF, Δt)
cs_rcm_flux_buffer(rcm, dF) = ccall((:cs_rcm_flux_buffer, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Float64}}), rcm, dF)
cs_rcm_peer_mailbox(rcm, nranks, mailbox, nbytes) =
    ccall((:cs_rcm_peer_mailbox, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}, Ref{Int64}), rcm, nranks, mailbox, nbytes)
cs_rcm_peer_connect(rcm, rank, nranks, mailboxes) =
    ccall((:cs_rcm_peer_connect, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Ptr{Cvoid}}), rcm, rank, nranks, mailboxes)
cs_rcm_enqueue_step_peer(rcm, Δt) = ccall((:cs_rcm_enqueue_step_peer, LIB), Int32, (Ptr{Cvoid}, Float64), rcm, Δt)
cs_rcm_peer_status(rcm, steps, timedout) =
    ccall((:cs_rcm_peer_status, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int32}), rcm, steps, timedout)
cs_ipc_export(p, handle) = ccall((:cs_ipc_export, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), p, handle)
cs_ipc_open(ctx, handle, p) = ccall((:cs_ipc_open, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Ref{Ptr{Cvoid}}), ctx, handle, p)
cs_ipc_close(ctx, p) = ccall((:cs_ipc_close, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, p)
cs_par_parse(ctx, nbytes, text, reclen, nrec, M, I, ν, S, A, γa, γs, Epp, na, δa, flags) =
    ccall((:cs_par_parse, LIB), Int32,
          (Ptr{Cvoid}, Int64, Ptr{UInt8}, Int32, Int64, Ptr{Int16}, Ptr{Int16}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
           Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
          ctx, nbytes, text, reclen, nrec, M, I, ν, S, A, γa, γs, Epp, na, δa, flags)
cs_par_read(ctx, nbytes, text, reclen, nrec, νmin, νmax, Scut, nI, Ilist, maxlines, M, I, ν, S, A, γa, γs, Epp, na, δa, index, nout, nbad) =
    ccall((:cs_par_read, LIB), Int32,
          (Ptr{Cvoid}, Int64, Ptr{UInt8}, Int32, Int64, Float64, Float64, Float64, Int32, Ptr{Int16}, Int64, Ptr{Int16},
           Ptr{Int16}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
           Ptr{Float64}, Ptr{Int64}, Ref{Int64}, Ref{Int64}),
          ctx, nbytes, text, reclen, nrec, νmin, νmax, Scut, nI, Ilist, maxlines, M, I, ν, S, A, γa, γs, Epp, na, δa, index, nout, nbad)
cs_group_create(ndev, devices, out) =
    ccall((:cs_group_create, LIB), Int32, (Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}), ndev, devices, out)
cs_group_free(grp) = ccall((:cs_group_free, LIB), Int32, (Ptr{Cvoid},), grp)
cs_group_size(grp, n) = ccall((:cs_group_size, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}), grp, n)
cs_group_ctx(grp, i, ctx) = ccall((:cs_group_ctx, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), grp, i, ctx)
cs_group_buffer(grp, i, count, dptr) =
    ccall((:cs_group_buffer, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Ref{Ptr{Float64}}), grp, i, count, dptr)
cs_group_allreduce_sum(grp, count) = ccall((:cs_group_allreduce_sum, LIB), Int32, (Ptr{Cvoid}, Int64), grp, count)
cs_group_read(grp, i, count, host) =
    ccall((:cs_group_read, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}), grp, i, count, host)
cs_group_rcm_step(grp, rcm, Δt, nsteps) =
    ccall((:cs_group_rcm_step, LIB), Int32, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Float64, Int64), grp, rcm, Δt, nsteps)
end # module Lib

# =============================================================================================================
# error plumbing, context, handles
# =============================================================================================================
# status code + thread-local message -> Julia exception; CS_ERR_ARG mirrors a reference @assert
function check(rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(Lib.cs_last_error())
    rc == 2 ? throw(AssertionError(msg)) : error("libclearsky_b200 [$rc]: $msg")
end

mutable struct Context
    h::Handle
    borrowed::Bool
end
function Context(device::Integer=parse(Int, get(ENV, "CLEARSKY_B200_DEVICE", "0")))
    r = Ref{Handle}(C_NULL)
    check(Lib.cs_ctx_create(Int32(device), r))
    c = Context(r[], false)
    finalizer(x -> x.borrowed || Lib.cs_ctx_free(x.h), c)
end
const CTX = Ref{Union{Nothing,Context}}(nothing)
context() = (CTX[] === nothing && (CTX[] = Context()); CTX[]::Context)

version() = Int(Lib.cs_version())
function devicecount()
    n = Ref{Int32}(0)
    Lib.cs_device_count(n) == 0 ? Int(n[]) : 0
end
synchronize(ctx::Context=context()) = check(Lib.cs_ctx_synchronize(ctx.h))
const TIMERNAMES = (:prep, :linesum, :table_fit, :table_eval, :cia, :rt, :reduce, :total)
# kernel milliseconds of the most recent call / accumulated since context creation (CUDA events, read lazily)
function timers(ctx::Context=context(); total::Bool=false)
    t = zeros(F64, NTIMERS)
    check(total ? Lib.cs_ctx_timers_total(ctx.h, t) : Lib.cs_ctx_timers(ctx.h, t))
    NamedTuple{TIMERNAMES}(Tuple(t))
end
function launches(ctx::Context=context())
    n = Ref{Int64}(0)
    check(Lib.cs_ctx_launches(ctx.h, n))
    Int(n[])
end
function fp64peak(ctx::Context=context(); iters::Integer=20000)
    v = Ref{F64}(0.0)
    check(Lib.cs_fp64_peak(ctx.h, Int32(iters), v))
    v[]
end
# far-wing treatment of the line sum: :direct (every pair, like surf!) or :expansion (include/clearsky_b200.h)
farfield!(mode::Symbol, ctx::Context=context()) =
    check(Lib.cs_ctx_set_farfield(ctx.h, mode === :expansion ? Int32(1) : Int32(0)))
function farfield(ctx::Context=context())
    m = Ref{Int32}(0)
    check(Lib.cs_ctx_get_farfield(ctx.h, m))
    m[] == 1 ? :expansion : :direct
end
# floor on the vertical optical depth of a layer in the flux kernel (reference: 1e-6, src/core/discretized.jl:174)
taufloor!(τmin::Real, ctx::Context=context()) = check(Lib.cs_ctx_set_tau_floor(ctx.h, F64(τmin)))

# =============================================================================================================
# SpectralLines on the device (cs_lines_upload <- src/hitran/par.jl:224-284)
# =============================================================================================================
mutable struct DeviceLines
    h::Handle
    ctx::Context
end
const LINES = IdDict{Any,DeviceLines}()       # key: (sl, ctx)

function devicelines(sl::SpectralLines, ctx::Context=context(); gridrange::Union{Nothing,NTuple{2,Float64}}=nothing)
    get!(LINES, (sl, ctx, gridrange)) do
        mp = MOLPARAM[sl.M]
        niso = length(mp.A)
        cheb = zeros(F64, MAXCHEB, niso)                  # column-major == C [niso][MAXCHEB]
        for i in 1:niso
            cheb[1:mp.ncheb[i], i] .= mp.cheb[i]
        end
        r = Ref{Handle}(C_NULL)
        check(Lib.cs_lines_upload(ctx.h, Int64(sl.N), sl.ν, sl.S, sl.γa, sl.γs, sl.Epp, sl.na, sl.μ, sl.I, Int32(niso),
                                  Int32.(mp.ncheb), cheb, UInt8.(mp.hascheb), r))
        d = DeviceLines(r[], ctx)
        # ν-sharded runs: the strict includedlines prefilter (line_shapes.jl:18-22) refers to the GLOBAL grid
        gridrange === nothing || check(Lib.cs_lines_set_grid_range(d.h, gridrange[1], gridrange[2]))
        finalizer(x -> Lib.cs_lines_free(x.h), d)
    end
end

# vector forms of scaleintensity / αdoppler / γlorentz (line_shapes.jl:125-132,146-148,259-261) for all lines
function lineparams_b200(sl::SpectralLines, T::Real, P::Real, Pₚ::Real)
    S, α, γ = zeros(F64, sl.N), zeros(F64, sl.N), zeros(F64, sl.N)
    check(Lib.cs_line_params(devicelines(sl).h, F64(T), F64(P), F64(Pₚ), S, α, γ))
    S, α, γ
end
# exact number of surf! inner-loop iterations for one (T,P) node (line_shapes.jl:75-82)
function countevals_b200(sl::SpectralLines, ν::AbstractVector, Δνcut::Real)
    νv = collect(F64, ν)
    n = Ref{Int64}(0)
    check(Lib.cs_count_evals(devicelines(sl).h, Int64(length(νv)), νv, F64(Δνcut), n))
    Int(n[])
end

# =============================================================================================================
# S1: in-place line shapes, same seven arguments as the reference (line_shapes.jl:200,313,412,527)
# =============================================================================================================
for (name, id, cut) in ((:voigt_b200!, VOIGT, 25.0), (:lorentz_b200!, LORENTZ, 25.0),
                        (:doppler_b200!, DOPPLER, 25.0), (:PHCO2_b200!, PHCO2_ID, 500.0))
    @eval function $name(σ::AbstractVector, ν::AbstractVector, sl::SpectralLines, T, P, Pₚ, Δνcut=$cut)
        νv = collect(F64, ν)
        out = Vector{F64}(undef, length(νv))
        check(Lib.cs_xsec(devicelines(sl).h, $id, Int64(length(νv)), νv, Int64(1), F64[T], F64[P], F64[Pₚ], F64(Δνcut), out))
        σ .= out
        nothing
    end
    @eval shapeid(::typeof($name)) = $id
end
const B200Shape = Union{typeof(voigt_b200!),typeof(lorentz_b200!),typeof(doppler_b200!),typeof(PHCO2_b200!)}

# batched form: σ[nν, nlev] for nlev (T, P, Pₚ) nodes in one call
function xsec_b200(shape!::B200Shape, ν::AbstractVector, sl::SpectralLines, T::AbstractVector, P::AbstractVector,
                   Pₚ::AbstractVector, Δνcut::Real)
    νv = collect(F64, ν)
    σ = Matrix{F64}(undef, length(νv), length(T))            # column-major [nν, nlev] == C [nlev][nν]
    check(Lib.cs_xsec(devicelines(sl).h, shapeid(shape!), Int64(length(νv)), νv, Int64(length(T)), collect(F64, T), collect(F64, P),
                      collect(F64, Pₚ), F64(Δνcut), σ))
    σ
end

# =============================================================================================================
# S2: Gas whose OpacityTables live on the GPU (bake <- src/absorption/gases.jl:97-145, Gas ctor :225-238)
# =============================================================================================================
mutable struct DeviceTable
    h::Handle
    ctx::Context
    nν::Int
end
# element type of Gas.Π: the i-th wavenumber of a device table.  Callable like an OpacityTable (gases.jl:85) so that every
# stock scalar code path (rawσ(g, i, T, P), g(i, T, P), Σ of a UnifiedAbsorber) keeps working -- one launch per call, slow;
# the batched paths below never go through it.
struct B200Π
    table::DeviceTable
    i::Int
end
function (Π::B200Π)(T, P)
    out = Vector{F64}(undef, Π.table.nν)
    check(Lib.cs_table_eval(Π.table.h, Int64(1), F64[T], F64[P], out))
    out[Π.i]
end

# Gas(sl, fC, ν, Ω, voigt_b200!, Δνcut): ONE cs_bake call for all nT·nP nodes; the result is a STOCK `Gas`, so it pairs with
# CIATables (isa(g, Gas), absorbers.jl:66), counts in pressurelimits (:249) and passes every type check of the reference
function ClearSky.Gas(sl::SpectralLines, fC::F, ν::AbstractVector{<:Real}, Ω::AtmosphericDomain, shape!::B200Shape,
                      Δνcut::Real=(shape! === PHCO2_b200! ? 500 : 25); progress::Bool=true, keepblock::Bool=false,
                      ctx::Context=context()) where {F}
    @assert length(ν) > 0
    μ = sum(sl.A .* sl.μ) / sum(sl.A)                          # gases.jl:233
    νv = collect(F64, ν)
    checkν(νv)
    C = F64[fC(T, P) for T in Ω.T, P in Ω.P]                   # C[i,j] = fC(T_i, P_j), column-major: index i + nT*(j-1)
    r = Ref{Handle}(C_NULL)
    check(Lib.cs_bake(devicelines(sl, ctx).h, shapeid(shape!), Int64(length(νv)), νv, Int32(Ω.nT), Ω.T, Int32(Ω.nP), Ω.P, C,
                      F64(Δνcut), Int32(keepblock), r))
    tb = DeviceTable(r[], ctx, length(νv))
    finalizer(x -> Lib.cs_table_free(x.h), tb)
    Π = [B200Π(tb, i) for i in eachindex(νv)]
    Gas{B200Π,F}(sl.name, sl.formula, μ, νv, Ω, Π, fC)
end
ClearSky.Gas(sl::SpectralLines, C::Real, ν::AbstractVector{<:Real}, Ω::AtmosphericDomain, shape!::B200Shape, args...; kw...) =
    ClearSky.Gas(sl, (T, P) -> float(C), ν, Ω, shape!, args...; kw...)

# rawσ(g, T, P) for all wavenumbers in one launch (gases.jl:263)
function ClearSky.rawσ(g::Gas{B200Π}, T, P)
    out = Vector{F64}(undef, length(g.ν))
    check(Lib.cs_table_eval(g.Π[1].table.h, Int64(1), F64[T], F64[P], out))
    out
end

function tableinfo(tb::DeviceTable)
    nν, nT, nP, nz = Ref{Int64}(0), Ref{Int32}(0), Ref{Int32}(0), Ref{Int64}(0)
    check(Lib.cs_table_info(tb.h, nν, nT, nP, nz))
    (nν=Int(nν[]), nT=Int(nT[]), nP=Int(nP[]), nzeroed=Int(nz[]))
end

# hand a GPU-baked gas back to unmodified ClearSky code: stock OpacityTables from the σ[nν,nT,nP] block (needs keepblock=true)
function stockgas(g::Gas{B200Π})
    tb = g.Π[1].table
    σ = Array{F64,3}(undef, length(g.ν), g.Ω.nT, g.Ω.nP)
    check(Lib.cs_table_block(tb.h, σ))
    Π = [OpacityTable(g.Ω.T, g.Ω.P, σ[i, :, :]) for i in eachindex(g.ν)]       # gases.jl:144
    Gas(g.name, g.formula, g.μ, g.ν, g.Ω, Π, g.fC)
end

# device table of ANY stock Gas: a B200 gas carries it; a gas with host OpacityTables (stock bake, deserialised) is
# re-sampled at its own Chebyshev nodes -- the interpolant reproduces its node values -- and fitted on the device
const TABLES = IdDict{Any,DeviceTable}()
devicetable(g::Gas{B200Π}, ::Context=context()) = g.Π[1].table
function devicetable(g::Gas, ctx::Context=context())
    get!(TABLES, (g.Π, ctx)) do
        Ω = g.Ω
        σ = Array{F64,3}(undef, length(g.ν), Ω.nT, Ω.nP)
        for j in 1:Ω.nP, i in 1:Ω.nT, v in eachindex(g.ν)
            σ[v, i, j] = g.Π[v](Ω.T[i], Ω.P[j])
        end
        r = Ref{Handle}(C_NULL)
        check(Lib.cs_table_from_block(ctx.h, Int64(length(g.ν)), Int32(Ω.nT), Ω.T, Int32(Ω.nP), Ω.P, σ, r))
        tb = DeviceTable(r[], ctx, length(g.ν))
        finalizer(x -> Lib.cs_table_free(x.h), tb)
    end
end

# exact line-by-line gas at the quadrature nodes (no table): engine-native, <: AbstractGas so UnifiedAbsorber accepts it
struct B200LineGas{F} <: AbstractGas
    name::String
    formula::String
    μ::Float64
    ν::Vector{Float64}
    sl::SpectralLines
    shape::Int32
    Δνcut::Float64
    fC::F
end
B200LineGas(sl::SpectralLines, fC::F, ν::AbstractVector{<:Real}, shape!::B200Shape=voigt_b200!,
            Δνcut::Real=(shape! === PHCO2_b200! ? 500 : 25)) where {F} =
    B200LineGas{F}(sl.name, sl.formula, sum(sl.A .* sl.μ) / sum(sl.A), collect(F64, ν), sl, shapeid(shape!), F64(Δνcut), fC)
ClearSky.concentration(g::B200LineGas, T, P) = g.fC(T, P)

# =============================================================================================================
# Σ(𝒜, idx, T, P) for all idx at all quadrature nodes (absorbers.jl:84-97) in a device workspace [node][ν]
# =============================================================================================================
mutable struct Workspace
    h::Handle
    ctx::Context
    nν::Int
    nnode::Int
end
function Workspace(ν::Vector{F64}, nnode::Integer, ctx::Context=context())
    r = Ref{Handle}(C_NULL)
    check(Lib.cs_sigma_create(ctx.h, Int64(length(ν)), ν, Int64(nnode), r))
    w = Workspace(r[], ctx, length(ν), nnode)
    finalizer(x -> Lib.cs_sigma_free(x.h), w)
end
zero!(w::Workspace) = check(Lib.cs_sigma_zero(w.h))
function Base.read(w::Workspace)
    σ = Matrix{F64}(undef, w.nν, w.nnode)                     # column-major [nν, nnode] == C [nnode][nν]
    check(Lib.cs_sigma_read(w.h, σ))
    σ
end

conc(g, T::Vector{F64}, P::Vector{F64}) = F64[concentration(g, t, p) for (t, p) in zip(T, P)]

# --- gases
function addto!(w::Workspace, g::Gas, T::Vector{F64}, P::Vector{F64})
    C = conc(g, T, P)
    check(Lib.cs_sigma_add_table(w.h, devicetable(g, w.ctx).h, T, P, C))
end
function addto!(w::Workspace, g::B200LineGas, T::Vector{F64}, P::Vector{F64})
    C = conc(g, T, P)
    check(Lib.cs_sigma_add_lines(w.h, devicelines(g.sl, w.ctx).h, g.shape, T, P, C, g.Δνcut))
end
addto!(w::Workspace, g::GrayGas, T::Vector{F64}, P::Vector{F64}) = check(Lib.cs_sigma_add_gray(w.h, F64(g.σ), Inf))
addto!(w::Workspace, g::SemiGrayGas, T::Vector{F64}, P::Vector{F64}) = check(Lib.cs_sigma_add_gray(w.h, F64(g.σ), F64(g.νcut)))

# --- CIA: CIATables cross as flattened arrays (collision_induced_absorption.jl:145-242): per >= 2-temperature group the
# grid vectors Φ.G.x / Φ.G.y and log k (ν fastest; floatmin substitution already done by the constructor, :205), per
# single-temperature range ν and log k (:187-188)
mutable struct DeviceCIA
    h::Handle
end
const CIAS = IdDict{Any,DeviceCIA}()
function devicecia(x::CIATables, ctx::Context=context())
    get!(CIAS, (x, ctx)) do
        gnν, gnT = Int64[], Int64[]
        gν, gT, glnk = F64[], F64[], F64[]
        for Φ in x.Φ
            push!(gnν, length(Φ.G.x)); push!(gnT, length(Φ.G.y))
            append!(gν, Φ.G.x); append!(gT, Φ.G.y); append!(glnk, vec(Φ.G.Z))       # Z[iν, jT], ν fastest
        end
        sn, sν, slnk = Int64[], F64[], F64[]
        for ϕ in x.ϕ
            push!(sn, length(ϕ.r.x)); append!(sν, ϕ.r.x); append!(slnk, ϕ.r.y)
        end
        r = Ref{Handle}(C_NULL)
        check(Lib.cs_cia_upload(ctx.h, Int32(length(gnν)), gnν, gnT, gν, gT, glnk, Int32(length(sn)), sn, sν, slnk,
                                Int32(x.extrapolate), Int32(x.singles), r))
        d = DeviceCIA(r[])
        finalizer(y -> Lib.cs_cia_free(y.h), d)
    end
end
# the CIA functor χ(ν, T, P) = cia(ν, χ.x, T, P, P·C₁(T,P), P·C₂(T,P)) (:378-382, :465) for all ν at all nodes
function addto!(w::Workspace, χ::CIA, T::Vector{F64}, P::Vector{F64})
    C₁, C₂ = conc(χ.g₁, T, P), conc(χ.g₂, T, P)
    check(Lib.cs_sigma_add_cia(w.h, devicecia(χ.x, w.ctx).h, T, P, C₁, C₂))
end

# --- user functions σ(ν,T,P) cannot cross the ABI: pre-evaluate on the host, node-major
function addto!(w::Workspace, f::Function, ν::Vector{F64}, T::Vector{F64}, P::Vector{F64})
    σ = F64[f(x, t, p) for x in ν, (t, p) in zip(T, P)]       # [nν, nnode] column-major == C [nnode][nν]
    check(Lib.cs_sigma_add_host(w.h, σ))
end

# --- UnifiedAbsorber: σchain over its three tuples (absorbers.jl:84-95)
function addto!(w::Workspace, U::UnifiedAbsorber, T::Vector{F64}, P::Vector{F64})
    for g in U.gas
        addto!(w, g, T, P)
    end
    for χ in U.cia
        addto!(w, χ, T, P)
    end
    for f in U.fun
        addto!(w, f, U.ν, T, P)
    end
    nothing
end

# --- AcceleratedAbsorber (absorbers.jl:114-209): Σ = exp(ϕᵢ(ln P)), T ignored (:203).  The device twin holds ln σ at the
# object's own levels; it is uploaded from the stock interpolators once and again after refresh!(A) (call that after any
# update!(A, T) done outside this module; AcceleratedAbsorber_b200 / update_b200! below keep both sides in step themselves)
mutable struct DeviceAccel
    h::Handle
end
const ACCELS = IdDict{Any,DeviceAccel}()
lnσmatrix(A::AcceleratedAbsorber) = F64[A.ϕ[i][k] for i in 1:A.nν, k in eachindex(A.P)]     # [nν, nlev]; ϕ[k] = stored ln σ (absorbers.jl:195)
function deviceaccel(A::AcceleratedAbsorber, ctx::Context=context())
    get!(ACCELS, (A, ctx)) do
        r = Ref{Handle}(C_NULL)
        check(Lib.cs_accel_upload(ctx.h, Int64(A.nν), Int64(length(A.P)), A.P, lnσmatrix(A), r))
        d = DeviceAccel(r[])
        finalizer(y -> Lib.cs_accel_free(y.h), d)
    end
end
refresh!(A::AcceleratedAbsorber) = (for k in collect(keys(ACCELS)); k[1] === A && delete!(ACCELS, k); end; nothing)
addto!(w::Workspace, A::AcceleratedAbsorber, T::Vector{F64}, P::Vector{F64}) =
    check(Lib.cs_sigma_add_accel(w.h, deviceaccel(A, w.ctx).h, P))

# batched construction / update of a STOCK AcceleratedAbsorber: Σ of the unified absorber at all levels in one pass on the
# device (the reference's update! is serial over levels and wavenumbers, absorbers.jl:173-200), values written into the
# stock interpolators so that every reference code path sees the same object
function update_b200!(A::AcceleratedAbsorber, T::AbstractVector, ctx::Context=context())
    @assert length(T) == length(A.P)
    Tv = collect(F64, T)
    w = Workspace(A.ν, length(A.P), ctx)
    addto!(w, A.U, Tv, A.P)
    r = Ref{Handle}(C_NULL)
    check(Lib.cs_accel_from_sigma(w.h, A.P, r))              # max(log Σ, log floatmin) snapshot (:193-195)
    d = DeviceAccel(r[])
    finalizer(y -> Lib.cs_accel_free(y.h), d)
    ACCELS[(A, ctx)] = d
    σ = read(w)
    logtiny = log(floatmin(F64))
    for k in eachindex(A.P), i in 1:A.nν
        lnσ = log(σ[i, k])
        A.ϕ[i][k] = lnσ < logtiny ? logtiny : lnσ
    end
    A.T .= Tv
    nothing
end
function AcceleratedAbsorber_b200(T::AbstractVector{<:Real}, P::AbstractVector{<:Real}, absorbers...; ctx::Context=context())
    U, ν, nν = unifyabsorbers(absorbers)
    U isa AcceleratedAbsorber && return U
    idx = sortperm(P)
    Pv, Tv = collect(F64, P[idx]), collect(F64, T[idx])
    lnP = log.(Pv)
    ϕ = [LinearInterpolator(lnP, Vector{F64}(undef, length(Pv)), NoBoundaries()) for _ in 1:nν]      # absorbers.jl:148-151
    A = AcceleratedAbsorber{F64,typeof(U)}(ϕ, ν, nν, Tv, Pv, U)
    update_b200!(A, Tv, ctx)
    A
end

# =============================================================================================================
# S3: the numerical core
# =============================================================================================================
struct B200Discretized <: AbstractNumericalCore
    nstream::Int64
    nlobatto::Int64
    materialize::Bool      # false: radiate! leaves F.M⁺/F.M⁻/F.τ untouched (the RCM loop only consumes Fnet)
end
B200Discretized(; nstream::Int=5, nlobatto::Int=2, materialize::Bool=true) = B200Discretized(nstream, nlobatto, materialize)
B200Discretized(nstream::Int, nlobatto::Int) = B200Discretized(nstream, nlobatto, true)

# the (np-1)(nlobatto-1)+1 distinct quadrature nodes in ascending pressure; the shared end node of a layer uses the exact
# level pressure and T[end, i] (core/discretized.jl:19-27,169)
function uniquenodes(P::Vector{F64}, Tl::Matrix, 𝓍::Vector{F64})
    np, nl = length(P), length(𝓍)
    Pn, Tn = F64[P[1]], F64[Tl[1, 1]]
    for i in 1:np-1, n in 2:nl
        push!(Pn, n == nl ? P[i+1] : P[i] + (P[i+1] - P[i]) * 𝓍[n])
        push!(Tn, Tl[n, i])
    end
    Tn, Pn
end

struct Prepared
    𝒜::Any
    ν::Vector{F64}
    P::Vector{F64}
    μl::Matrix{F64}
    Tlev::Vector{F64}
    w::Workspace
end
# everything the reference's Discretized method pre-evaluates (fluxes.jl:247-268), then Σ at the nodes on the device
function prepare(core::B200Discretized, P, T, μ, absorbers, ctx::Context)
    𝒜, ν, nν = unifyabsorbers(absorbers)
    𝒻T, 𝒻μ = formprofiles(P, T, μ)
    Pv = collect(F64, P)
    Tl, μl = lobattoevaluations(Pv, 𝒻T, 𝒻μ, core.nlobatto)       # [nlobatto, np-1] exactly as the ABI wants them
    𝓍, _ = lobattonodes(core.nlobatto)
    Tn, Pn = uniquenodes(Pv, Tl, 𝓍)
    w = Workspace(ν, length(Pn), ctx)
    addto!(w, 𝒜, Tn, Pn)
    Prepared(𝒜, ν, Pv, Matrix{F64}(μl), F64[𝒻T(p) for p in Pv], w)
end
# loose absorbers (gases, CIATables, functions) are unified first, exactly like the reference does (absorbers.jl:219-223)
addto!(w::Workspace, 𝒜::Tuple, T::Vector{F64}, P::Vector{F64}) = addto!(w, UnifiedAbsorber(𝒜), T, P)

spectral(𝒻, ν::Vector{F64}) = F64[𝒻(x) for x in ν]
hostptr(x::Nothing) = NULLF
hostptr(x::Array{F64}) = pointer(x)

# fused path: fills F⁺, F⁻, Fnet; materialises M⁺ / M⁻ / τ only into arrays the caller passes (Julia layouts:
# M±[np, nν], τ[np-1, nν])
function fluxes_b200(core::B200Discretized, P::AbstractVector{<:Real}, g::Real, T, μ, 𝒻S, 𝒻a, absorbers...;
                     θₛ::Real=0.841, M⁺=nothing, M⁻=nothing, τ=nothing, ctx::Context=context())
    @assert issorted(P) "pressure coordinates must be in ascending order (sorted)"          # fluxes.jl:257
    pr = prepare(core, P, T, μ, absorbers, ctx)
    checkpressures(pr.𝒜, pr.P[end], pr.P[1])                                                # fluxes.jl:265
    checkstreams(core.nstream)
    checkazimuth(θₛ)
    _, 𝓌 = lobattonodes(core.nlobatto)
    𝓂, 𝒲 = streamnodes(core.nstream)
    np = length(pr.P)
    F⁺, F⁻, Fnet = zeros(F64, np), zeros(F64, np), zeros(F64, np)
    fS, fa = spectral(𝒻S, pr.ν), spectral(𝒻a, pr.ν)
    GC.@preserve M⁺ M⁻ τ check(Lib.cs_fluxes(pr.w.h, Int64(np), pr.P, Int32(core.nlobatto), 𝓌, pr.μl, pr.Tlev, F64(g), fS, fa, F64(θₛ),
                                             Int32(core.nstream), 𝓂, 𝒲, NULLF, hostptr(τ), hostptr(M⁺), hostptr(M⁻), F⁺, F⁻, Fnet))
    F⁺, F⁻, Fnet
end

# the seam itself: same signature as the reference's Discretized method (fluxes.jl:238-279); reached from
# monochromaticfluxes (:303), fluxes (:334) and the generic radiate! (:377) with ONE unified object as `absorbers`
function ClearSky.monochromaticfluxes!(M⁺::AbstractMatrix, M⁻::AbstractMatrix, τ::AbstractMatrix, core::B200Discretized,
                                       P::AbstractVector{<:Real}, g::Real, T, μ, 𝒻S, 𝒻a, absorbers...; θₛ::Real=0.841)::Nothing
    np, nν = size(M⁺)
    Mu, Md, tt = Matrix{F64}(undef, np, nν), Matrix{F64}(undef, np, nν), Matrix{F64}(undef, np - 1, nν)
    fluxes_b200(core, P, g, T, μ, 𝒻S, 𝒻a, absorbers...; θₛ=θₛ, M⁺=Mu, M⁻=Md, τ=tt)
    M⁺ .= Mu; M⁻ .= Md; τ .= tt
    nothing
end

# radiate!(F, core, ...) (fluxes.jl:357-383) dispatches on the positional core: the spectral integral is fused on the device
# (no host ∫F!), the monochromatic blocks are copied back only when core.materialize -- this is the method RCM.heating!
# reaches (radiative_convective.jl:113) with 𝒜::AcceleratedAbsorber
function ClearSky.radiate!(F::FluxPack, core::B200Discretized, P::AbstractVector{<:Real}, g::Real, T, μ, 𝒻S, 𝒻a,
                           absorbers...; θₛ::Real=0.841)::Nothing
    _, _, nν = unifyabsorbers(absorbers)
    @assert size(F) == (length(P), nν) "size of FluxPack does not match number of pressure or wavenumber coordinates"
    dense = core.materialize && F.M⁺ isa Matrix{F64}
    F⁺, F⁻, Fnet = dense ? fluxes_b200(core, P, g, T, μ, 𝒻S, 𝒻a, absorbers...; θₛ=θₛ, M⁺=F.M⁺, M⁻=F.M⁻, τ=F.τ) :
                           fluxes_b200(core, P, g, T, μ, 𝒻S, 𝒻a, absorbers...; θₛ=θₛ)
    F.F⁺ .= F⁺; F.F⁻ .= F⁻; F.Fnet .= Fnet
    nothing
end

# opticaldepth(P::Vector, g, T, μ, θ, absorbers...; nlobatto=4) (fluxes.jl:68-97; 𝒹depth discretized.jl:92-134): the
# reference method has no core argument to dispatch on, hence the _b200 name
function opticaldepth_b200(P::AbstractVector{<:Real}, g::Real, T, μ, θ::Real, absorbers...; nlobatto::Int=4, ctx::Context=context())
    Ps = sort(collect(F64, P))
    pr = prepare(B200Discretized(5, nlobatto), Ps, T, μ, absorbers, ctx)
    checkpressures(pr.𝒜, Ps[end], Ps[1])
    checkazimuth(θ)
    _, 𝓌 = lobattonodes(nlobatto)
    τ = Vector{F64}(undef, length(pr.ν))
    check(Lib.cs_opticaldepth(pr.w.h, Int64(length(Ps)), Ps, Int32(nlobatto), 𝓌, pr.μl, F64(g), F64(θ), τ))
    τ
end

# =============================================================================================================
# RCM (src/radiative_convective.jl): batched jacobian! and the device-resident step loop
# =============================================================================================================
# the O(np) tail of heating! (:123-143) for a given net-flux profile
function heatingfrom(ℛ::RCM, Fnet::AbstractVector)
    𝒻F = AtmosphericProfile(ℛ.Pᵣ, collect(Fnet))
    R = [-𝒻F(p) for p in ℛ.Pₑ]
    H = similar(R)
    for i in 1:ℛ.np-1
        H[i] = (ℛ.g / ℛ.𝒻cₚ(ℛ.T[i], ℛ.P[i])) * (R[i] - R[i+1]) / (ℛ.Pₑ[i+1] - ℛ.Pₑ[i])
    end
    H[end] = R[end] / ℛ.cₛ
    R, H
end

# jacobian!(ℛ, ϵ) (:154-171) = np+1 flux solves that differ only in T; the AcceleratedAbsorber ignores T, so ONE
# cs_fluxes_batch call shares every layer depth and transmittance between them
function jacobian_b200!(ℛ::RCM, ϵ::Real=1; ctx::Context=context())::Nothing
    core = ℛ.core isa B200Discretized ? ℛ.core : B200Discretized(ℛ.core.nstream, ℛ.core.nlobatto)
    nr, np = length(ℛ.Pᵣ), ℛ.np
    pr = prepare(core, ℛ.Pᵣ, AtmosphericProfile(ℛ.P, ℛ.T), ℛ.𝒻μ, (ℛ.𝒜,), ctx)
    Tlev = Matrix{F64}(undef, nr, np + 1)                       # column-major [nrad, nbatch] == C [nbatch][nrad]
    for b in 0:np
        Tb = copy(ℛ.T)
        b > 0 && (Tb[b] += ϵ)
        𝒻T = AtmosphericProfile(ℛ.P, Tb)
        Tlev[:, b+1] .= 𝒻T.(ℛ.Pᵣ)
    end
    _, 𝓌 = lobattonodes(core.nlobatto)
    𝓂, 𝒲 = streamnodes(core.nstream)
    F = Matrix{F64}(undef, 2nr, np + 1)
    check(Lib.cs_fluxes_batch(pr.w.h, Int64(nr), pr.P, Int32(core.nlobatto), 𝓌, pr.μl, Int64(np + 1), Tlev, F64(ℛ.g), spectral(ℛ.𝒻S, pr.ν),
                              spectral(ℛ.𝒻a, pr.ν), 0.841, Int32(core.nstream), 𝓂, 𝒲, NULLF, F))
    R₀, H₀ = heatingfrom(ℛ, F[1:nr, 1] .- F[nr+1:2nr, 1])
    ℛ.R .= R₀; ℛ.H .= H₀
    ℛ.F.F⁺ .= F[1:nr, 1]; ℛ.F.F⁻ .= F[nr+1:2nr, 1]; ℛ.F.Fnet .= ℛ.F.F⁺ .- ℛ.F.F⁻
    for i in 1:np
        _, H = heatingfrom(ℛ, F[1:nr, i+1] .- F[nr+1:2nr, i+1])
        ℛ.J[:, i] .= (H .- H₀) ./ ϵ
    end
    nothing
end

# device-resident loop (cs_rcm_*): heating! + step! without leaving the GPU.  heating! never calls update! (:109-144), so Σ,
# the layer depths and the stream transmittances are computed once at construction; a step is three kernels replayed from a
# CUDA graph.  𝒻cₚ and 𝒻μ are evaluated once, at the temperatures the object is created with.
mutable struct DeviceRCM
    h::Handle
    ℛ::RCM
end
function DeviceRCM(ℛ::RCM; θₛ::Real=0.841, ctx::Context=context())
    core = ℛ.core isa B200Discretized ? ℛ.core : B200Discretized(ℛ.core.nstream, ℛ.core.nlobatto)
    pr = prepare(core, ℛ.Pᵣ, AtmosphericProfile(ℛ.P, ℛ.T), ℛ.𝒻μ, (ℛ.𝒜,), ctx)
    checkpressures(ℛ.𝒜, ℛ.Pᵣ[end], ℛ.Pᵣ[1])
    _, 𝓌 = lobattonodes(core.nlobatto)
    𝓂, 𝒲 = streamnodes(core.nstream)
    cₚ = F64[ℛ.𝒻cₚ(ℛ.T[i], ℛ.P[i]) for i in 1:ℛ.np-1]
    r = Ref{Handle}(C_NULL)
    check(Lib.cs_rcm_create(pr.w.h, Int64(ℛ.np), collect(F64, ℛ.Pₑ), collect(F64, ℛ.P), collect(F64, ℛ.T), cₚ, F64(ℛ.cₛ),
                            Int64(length(ℛ.Pᵣ)), pr.P, Int32(core.nlobatto), 𝓌, pr.μl, F64(ℛ.g), spectral(ℛ.𝒻S, pr.ν),
                            spectral(ℛ.𝒻a, pr.ν), F64(θₛ), Int32(core.nstream), 𝓂, 𝒲, NULLF, r))
    d = DeviceRCM(r[], ℛ)
    finalizer(x -> Lib.cs_rcm_free(x.h), d)
end
function rcminfo(d::DeviceRCM)
    np, nrad, nν = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(Lib.cs_rcm_info(d.h, np, nrad, nν))
    (np=Int(np[]), nrad=Int(nrad[]), nν=Int(nν[]))
end
# nsteps × step!(ℛ, Δt) (:147-151); Δt = 0 evaluates heating! only.  ℛ.T, ℛ.H, ℛ.R and ℛ.F are refreshed on return.
function step_b200!(d::DeviceRCM, Δt::Real, nsteps::Integer=1)::Nothing
    ℛ = d.ℛ
    check(Lib.cs_rcm_set_temperature(d.h, collect(F64, ℛ.T)))
    check(Lib.cs_rcm_step(d.h, F64(Δt), Int64(nsteps)))
    T, H, R = zeros(F64, ℛ.np), zeros(F64, ℛ.np), zeros(F64, ℛ.np)
    nr = length(ℛ.Pᵣ)
    F⁺, F⁻, Fnet = zeros(F64, nr), zeros(F64, nr), zeros(F64, nr)
    check(Lib.cs_rcm_state(d.h, T, H, R, F⁺, F⁻, Fnet))
    ℛ.T .= T; ℛ.H .= H; ℛ.R .= R
    ℛ.F.F⁺ .= F⁺; ℛ.F.F⁻ .= F⁻; ℛ.F.Fnet .= Fnet
    nothing
end
step_b200!(ℛ::RCM, Δt::Real, nsteps::Integer=1) = step_b200!(DeviceRCM(ℛ), Δt, nsteps)

# =============================================================================================================
# Radau-core entry points (fluxes.jl:39-66, 133-158) on the Discretized GPU core: layers equally spaced in ln P, doubled until
# two successive Richardson extrapolates agree to tol (clearsky_b200/radau.py is the executed twin)
# =============================================================================================================
function outgoing_b200(Pₛ::Real, g::Real, 𝒻T, 𝒻μ, absorbers...; Ptop::Real=1.0, nstream::Int=5, tol::Real=1e-5)
    taufloor!(1e-9)
    try
        _, ν, nν = unifyabsorbers(absorbers)
        prev, prevE, n = nothing, nothing, 32
        while true
            P = exp.(range(log(Ptop), log(Pₛ), length=n + 1))
            M⁺ = zeros(F64, n + 1, nν)
            fluxes_b200(B200Discretized(nstream, 4), P, g, 𝒻T, 𝒻μ, x -> 0.0, x -> 0.0, absorbers...; M⁺=M⁺)
            olr = M⁺[1, :]
            E = prev === nothing ? nothing : (4 .* olr .- prev) ./ 3          # Richardson extrapolate of the O(n⁻²) scheme
            if E !== nothing && prevE !== nothing
                scale = max.(abs.(E), 1e-3 * maximum(abs.(E)))
                (maximum(abs.(E .- prevE) ./ scale) < tol || 2n + 1 > 4097) && return E
            end
            prev, prevE, n = olr, E, 2n
        end
    finally
        taufloor!(1e-6)
    end
end

function opticaldepth_between_b200(P₁::Real, P₂::Real, g::Real, 𝒻T, 𝒻μ, θ::Real, absorbers...; tol::Real=1e-5)
    P₁, P₂ = max(P₁, P₂), min(P₁, P₂)
    prev, n = nothing, 16
    while true
        τ = opticaldepth_b200(exp.(range(log(P₂), log(P₁), length=n + 1)), g, 𝒻T, 𝒻μ, θ, absorbers...; nlobatto=4)
        if prev !== nothing
            scale = max.(abs.(τ), 1e-3 * maximum(abs.(τ)))
            (maximum(abs.(τ .- prev) ./ scale) < tol || 2n + 1 > 4097) && return τ
        end
        prev, n = τ, 2n
    end
end

# =============================================================================================================
# .par ingestion (readpar, src/hitran/par.jl:91-193): parse on the GPU, filter + sort on the GPU, gather on the host
# =============================================================================================================
function readpar_b200(filename::String; νmin::Real=0, νmax::Real=Inf, Scut::Real=0, I::Vector=[], maxlines::Int=-1,
                      ctx::Context=context())
    @assert filename[end-3:end] == ".par" "expected file with .par extension, downloaded from https://hitran.org/lbl/"
    text = read(filename)
    nl = findfirst(==(UInt8('\n')), text)
    reclen = nl === nothing ? length(text) : nl
    N = cld(length(text), reclen)
    M, Iv = zeros(Int16, N), zeros(Int16, N)
    col() = zeros(F64, N)
    ν, S, A, γa, γs, Epp, na, δa = col(), col(), col(), col(), col(), col(), col(), col()
    index = zeros(Int64, N)
    nout, nbad = Ref{Int64}(0), Ref{Int64}(0)
    # readpar's filters (:154-170), the maxlines truncation (:178-185) and the final sort by ν (:187-191) all happen on the
    # device; only the surviving records come back, in output order
    Ilist = Int16[i isa Char ? ISOINDEX[i] : Int16(i) for i in I]
    check(Lib.cs_par_read(ctx.h, Int64(length(text)), text, Int32(reclen), Int64(N), F64(νmin), F64(νmax), F64(Scut),
                          Int32(length(Ilist)), Ilist, Int64(maxlines), M, Iv, ν, S, A, γa, γs, Epp, na, δa, index, nout, nbad))
    @assert nbad[] == 0 "malformed numeric field in $filename"
    k = 1:nout[]
    # "I" as ISOINDEX numbers; the string columns (Vp, Vpp, ...) can be gathered from `text` with index .+ 1
    Dict("M" => M[k], "I" => Iv[k], "ν" => ν[k], "S" => S[k], "A" => A[k], "γa" => γa[k], "γs" => γs[k], "Epp" => Epp[k],
         "na" => na[k], "δa" => δa[k], "record" => index[k] .+ 1)
end
# the unfiltered, file-order form (parse loop only, par.jl:127-152)
function parsepar_b200(text::Vector{UInt8}, reclen::Integer; ctx::Context=context())
    N = cld(length(text), reclen)
    M, Iv = zeros(Int16, N), zeros(Int16, N)
    col() = zeros(F64, N)
    ν, S, A, γa, γs, Epp, na, δa = col(), col(), col(), col(), col(), col(), col(), col()
    flags = zeros(UInt8, N)
    check(Lib.cs_par_parse(ctx.h, Int64(length(text)), text, Int32(reclen), Int64(N), M, Iv, ν, S, A, γa, γs, Epp, na, δa, flags))
    (M=M, I=Iv, ν=ν, S=S, A=A, γa=γa, γs=γs, Epp=Epp, na=na, δa=δa, flags=flags)
end

# =============================================================================================================
# multi-GPU from ONE Julia process (cs_group_*): contiguous ν slices, lines within slice ± Δνcut, global trapezoid weights,
# one all-reduce of the 2·np integrated fluxes (SURVEY.md section 8e)
# =============================================================================================================
mutable struct DeviceGroup
    h::Handle
    ctx::Vector{Context}
end
function DeviceGroup(devices::AbstractVector{<:Integer}=collect(0:devicecount()-1))
    r = Ref{Handle}(C_NULL)
    check(Lib.cs_group_create(Int32(length(devices)), Int32.(devices), r))
    n = Ref{Int32}(0)
    check(Lib.cs_group_size(r[], n))
    ctxs = Context[]
    for i in 0:n[]-1
        c = Ref{Handle}(C_NULL)
        check(Lib.cs_group_ctx(r[], Int32(i), c))
        push!(ctxs, Context(c[], true))
    end
    grp = DeviceGroup(r[], ctxs)
    finalizer(x -> Lib.cs_group_free(x.h), grp)
end
Base.length(grp::DeviceGroup) = length(grp.ctx)

trapzweights(ν::Vector{F64}) = F64[((j > 1 ? ν[j] - ν[j-1] : 0.0) + (j < length(ν) ? ν[j+1] - ν[j] : 0.0)) / 2 for j in eachindex(ν)]

# fluxes(P, g, T, μ, 𝒻S, 𝒻a, absorbers...) with the spectrum sharded over the group.  build(νslice, ctx) returns the
# absorbers of one slice living on ctx (e.g. Gas(sl, fC, νslice, Ω, voigt_b200!; ctx=ctx)); edges are slice borders (1-based,
# length(grp)+1 entries).
function sharded_fluxes_b200(grp::DeviceGroup, build, ν::AbstractVector{<:Real}, edges::Vector{Int}, core::B200Discretized,
                             P::AbstractVector{<:Real}, g::Real, T, μ, 𝒻S, 𝒻a; θₛ::Real=0.841)
    νv = collect(F64, ν)
    wg = trapzweights(νv)
    np = length(P)
    _, 𝓌 = lobattonodes(core.nlobatto)
    𝓂, 𝒲 = streamnodes(core.nstream)
    @sync for i in 1:length(grp)
        Threads.@spawn begin
            a, b = edges[i], edges[i+1] - 1
            pr = prepare(core, P, T, μ, Tuple(build(νv[a:b], grp.ctx[i])), grp.ctx[i])
            dF = Ref{Ptr{F64}}(C_NULL)
            check(Lib.cs_group_buffer(grp.h, Int32(i - 1), Int64(2np), dF))
            check(Lib.cs_fluxes_device(pr.w.h, Int64(np), pr.P, Int32(core.nlobatto), 𝓌, pr.μl, pr.Tlev, F64(g), spectral(𝒻S, pr.ν),
                                       spectral(𝒻a, pr.ν), F64(θₛ), Int32(core.nstream), 𝓂, 𝒲, wg[a:b], dF[]))
        end
    end
    check(Lib.cs_group_allreduce_sum(grp.h, Int64(2np)))          # ncclAllReduce over NVLink
    F = zeros(F64, 2np)
    check(Lib.cs_group_read(grp.h, Int32(0), Int64(2np), F))
    F[1:np], F[np+1:2np], F[1:np] .- F[np+1:2np]
end

# radiative-convective steps of a ν-sharded column: one DeviceRCM per group member (created with that member's context and the
# slice's GLOBAL trapezoid weights), all advanced together by cs_group_rcm_step
function step_b200!(grp::DeviceGroup, ds::Vector{DeviceRCM}, Δt::Real, nsteps::Integer=1)::Nothing
    hs = Handle[d.h for d in ds]
    check(Lib.cs_group_rcm_step(grp.h, hs, F64(Δt), Int64(nsteps)))
    nothing
end

# lower-level pieces for callers that drive their own collective (one process per GPU, MPI/NCCL.jl): partial fluxes of a step
# into device memory, then the column update from the summed fluxes
function enqueue_fluxes!(d::DeviceRCM, dF::Ptr{F64}=NULLF)
    check(Lib.cs_rcm_enqueue_fluxes(d.h, dF))
end
function enqueue_update!(d::DeviceRCM, Δt::Real, dF::Ptr{F64}=NULLF)
    check(Lib.cs_rcm_enqueue_update(d.h, dF, F64(Δt)))
end
function fluxbuffer(d::DeviceRCM)
    p = Ref{Ptr{F64}}(C_NULL)
    check(Lib.cs_rcm_flux_buffer(d.h, p))
    p[]
end
# ν-sharded step with the collective fused into the step's last kernel (one Julia process per GPU: exchange the 64-byte
# handles of `peermailbox` with MPI.Allgather / Distributed, map them with `ipcopen`, then `peerconnect!`)
function peermailbox(d::DeviceRCM, nranks::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL); n = Ref{Int64}(0)
    check(Lib.cs_rcm_peer_mailbox(d.h, Int32(nranks), p, n))
    p[], n[]
end
function ipcexport(p::Ptr{Cvoid})
    h = Vector{UInt8}(undef, 64)
    GC.@preserve h check(Lib.cs_ipc_export(p, pointer(h)))
    h
end
function ipcopen(c::Context, h::Vector{UInt8})
    p = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve h check(Lib.cs_ipc_open(c.h, pointer(h), p))
    p[]
end
ipcclose(c::Context, p::Ptr{Cvoid}) = check(Lib.cs_ipc_close(c.h, p))
function peerconnect!(d::DeviceRCM, rank::Integer, mailboxes::Vector{Ptr{Cvoid}})
    GC.@preserve mailboxes check(Lib.cs_rcm_peer_connect(d.h, Int32(rank), Int32(length(mailboxes)), pointer(mailboxes)))
end
enqueue_step_peer!(d::DeviceRCM, Δt::Real) = check(Lib.cs_rcm_enqueue_step_peer(d.h, F64(Δt)))
function peerstatus(d::DeviceRCM)
    n = Ref{Int64}(0); t = Ref{Int32}(0)
    check(Lib.cs_rcm_peer_status(d.h, n, t))
    n[], t[] != 0
end
function rcmcontext(d::DeviceRCM)
    c = Ref{Handle}(C_NULL)
    check(Lib.cs_rcm_ctx(d.h, c))
    Context(c[], true)
end
# a context on a caller-owned CUDA stream (CUDA.jl: `Context(0, CUDA.stream().handle)`)
function Context(device::Integer, stream::Ptr{Cvoid})
    r = Ref{Handle}(C_NULL)
    check(Lib.cs_ctx_create_on_stream(Int32(device), stream, r))
    c = Context(r[], false)
    finalizer(x -> x.borrowed || Lib.cs_ctx_free(x.h), c)
end

end # module
