# ClearSkyB200.jl -- thin Julia wrapper that puts libclearsky_b200.so behind ClearSky.jl's own API.
#
# NOT EXECUTED in the build image (no Julia runtime there); it mirrors 1:1 the call sequence of the Python
# twin clearsky.jl_b200/clearsky_b200/ that the tests and bench.py drive.  Every `ccall` below binds an entry
# point declared in include/clearsky_b200.h.
#
# Seams used (SURVEY.md section 8b) -- no reference file is edited:
#   S1  shape!(σ, ν, sl, T, P, Pₚ, Δνcut)          -> voigt_b200!, lorentz_b200!, doppler_b200!, PHCO2_b200!
#   S2  Gas(sl, fC, ν, Ω, shape!, Δνcut) / bake     -> B200Gas(sl, fC, ν, Ω, B200Shape, Δνcut)
#   S3  monochromaticfluxes!(M⁺, M⁻, τ, core, ...)   -> methods for core::B200Discretized <: AbstractNumericalCore
module ClearSkyB200

using ClearSky
using ClearSky: SpectralLines, AtmosphericDomain, AbstractGas, AbstractNumericalCore, AbstractAbsorber,
                MOLPARAM, CIATables, FluxPack, formprofiles, lobattonodes, streamnodes, lobattoevaluations,
                checkazimuth, checkstreams, checkν

export B200Discretized, B200Gas, B200LineGas, voigt_b200!, lorentz_b200!, doppler_b200!, PHCO2_b200!
export farfield!, taufloor!, outgoing_b200, opticaldepth_b200

const LIB = get(ENV, "CLEARSKY_B200_LIB", joinpath(@__DIR__, "..", "lib", "libclearsky_b200.so"))
const MAXCHEB = 16
const DOPPLER, LORENTZ, VOIGT, PHCO2_ID = Int32(0), Int32(1), Int32(2), Int32(3)

# ---- error plumbing: status code + thread-local message -> Julia exception -------------------------------
function check(rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:cs_last_error, LIB), Cstring, ()))
    rc == 2 ? throw(AssertionError(msg)) : error("libclearsky_b200 [$rc]: $msg")
end

# ---- context (one per device; CLEARSKY_B200_DEVICE selects it) -------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=parse(Int, get(ENV, "CLEARSKY_B200_DEVICE", "0")))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:cs_ctx_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, r))
        c = new(r[])
        finalizer(x -> ccall((:cs_ctx_free, LIB), Int32, (Ptr{Cvoid},), x.h), c)
    end
end
const CTX = Ref{Union{Nothing,Context}}(nothing)
context() = (CTX[] === nothing && (CTX[] = Context()); CTX[])

# far-wing treatment of the line sum: :direct (every pair, like surf!) or :expansion (local expansions of well-separated
# far-wing lines about each 128-point tile: Voigt/Lorentz and the PHCO2 classes >= 30 cm^-1; truncation < 3e-11;
# include/clearsky_b200.h)
farfield!(mode::Symbol) = check(ccall((:cs_ctx_set_farfield, LIB), Int32, (Ptr{Cvoid}, Int32), context().h,
                                      mode === :expansion ? Int32(1) : Int32(0)))
# floor on the vertical optical depth of a layer in the flux kernel (reference: 1e-6, src/core/discretized.jl:174)
taufloor!(τmin::Real) = check(ccall((:cs_ctx_set_tau_floor, LIB), Int32, (Ptr{Cvoid}, Float64), context().h, τmin))

# ---- SpectralLines on the device (cs_lines_upload <- src/hitran/par.jl:224-284) --------------------------
mutable struct DeviceLines
    h::Ptr{Cvoid}
end
const LINES = IdDict{SpectralLines,DeviceLines}()

function devicelines(sl::SpectralLines)
    get!(LINES, sl) do
        mp = MOLPARAM[sl.M]
        niso = length(mp.A)
        cheb = zeros(Float64, MAXCHEB, niso)                  # column-major == C [niso][MAXCHEB]
        for i in 1:niso
            cheb[1:mp.ncheb[i], i] .= mp.cheb[i]
        end
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:cs_lines_upload, LIB), Int32,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Int16}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{UInt8}, Ref{Ptr{Cvoid}}),
            context().h, sl.N, sl.ν, sl.S, sl.γa, sl.γs, sl.Epp, sl.na, sl.μ, sl.I, niso,
            Int32.(mp.ncheb), cheb, UInt8.(mp.hascheb), r))
        d = DeviceLines(r[])
        finalizer(x -> ccall((:cs_lines_free, LIB), Int32, (Ptr{Cvoid},), x.h), d)
    end
end

# ---- S1: in-place line shapes, same seven arguments as the reference ------------------------------------
for (name, id, cut) in ((:voigt_b200!, VOIGT, 25.0), (:lorentz_b200!, LORENTZ, 25.0),
                        (:doppler_b200!, DOPPLER, 25.0), (:PHCO2_b200!, PHCO2_ID, 500.0))
    @eval function $name(σ::AbstractVector, ν::AbstractVector, sl::SpectralLines, T, P, Pₚ, Δνcut=$cut)
        νv = collect(Float64, ν)
        out = Vector{Float64}(undef, length(νv))
        check(ccall((:cs_xsec, LIB), Int32,
            (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Int64, Ref{Float64}, Ref{Float64}, Ref{Float64}, Float64, Ptr{Float64}),
            devicelines(sl).h, $id, length(νv), νv, 1, Float64(T), Float64(P), Float64(Pₚ), Float64(Δνcut), out))
        σ .= out
        nothing
    end
    @eval shapeid(::typeof($name)) = $id
end

# ---- S2: Gas whose OpacityTables live on the GPU (bake <- src/absorption/gases.jl:97-145) -----------------
mutable struct B200Gas{F} <: AbstractGas
    name::String
    formula::String
    μ::Float64
    ν::Vector{Float64}
    Ω::AtmosphericDomain
    h::Ptr{Cvoid}          # cs_table
    fC::F
end

function B200Gas(sl::SpectralLines, fC::F, ν::AbstractVector{<:Real}, Ω::AtmosphericDomain,
                 shape!::Function=voigt_b200!, Δνcut::Real=25) where {F}
    ν = collect(Float64, ν)
    checkν(ν)
    C = [fC(T, P) for T in Ω.T, P in Ω.P]                      # C[i,j] = fC(T_i, P_j), column-major
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:cs_bake, LIB), Int32,
        (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Int32,
         Ref{Ptr{Cvoid}}),
        devicelines(sl).h, shapeid(shape!), length(ν), ν, Ω.nT, Ω.T, Ω.nP, Ω.P, C, Float64(Δνcut), 0, r))
    g = B200Gas{F}(sl.name, sl.formula, sum(sl.A .* sl.μ) / sum(sl.A), ν, Ω, r[], fC)
    finalizer(x -> ccall((:cs_table_free, LIB), Int32, (Ptr{Cvoid},), x.h), g)
end

ClearSky.concentration(g::B200Gas, T, P) = g.fC(T, P)

# rawσ(g, T, P) for all wavenumbers (gases.jl:263)
function ClearSky.rawσ(g::B200Gas, T, P)
    out = Vector{Float64}(undef, length(g.ν))
    check(ccall((:cs_table_eval, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{Float64}, Ref{Float64}, Ptr{Float64}),
                g.h, 1, Float64(T), Float64(P), out))
    out
end

# exact line-by-line gas at the quadrature nodes (no table)
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
ClearSky.concentration(g::B200LineGas, T, P) = g.fC(T, P)

# ---- sigma workspace: Σ(𝒜, idx, T, P) for all idx at all nodes (absorbers.jl:84-95) ---------------------
mutable struct Workspace
    h::Ptr{Cvoid}
end
function Workspace(ν::Vector{Float64}, nnode::Integer)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:cs_sigma_create, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ref{Ptr{Cvoid}}),
                context().h, length(ν), ν, nnode, r))
    w = Workspace(r[])
    finalizer(x -> ccall((:cs_sigma_free, LIB), Int32, (Ptr{Cvoid},), x.h), w)
end

addto!(w::Workspace, g::B200Gas, T, P) =
    check(ccall((:cs_sigma_add_table, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                w.h, g.h, T, P, Float64[g.fC(t, p) for (t, p) in zip(T, P)]))
addto!(w::Workspace, g::B200LineGas, T, P) =
    check(ccall((:cs_sigma_add_lines, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64),
                w.h, devicelines(g.sl).h, g.shape, T, P, Float64[g.fC(t, p) for (t, p) in zip(T, P)], g.Δνcut))
addto!(w::Workspace, g::ClearSky.GrayGas, T, P) =
    check(ccall((:cs_sigma_add_gray, LIB), Int32, (Ptr{Cvoid}, Float64, Float64), w.h, Float64(g.σ), Inf))
addto!(w::Workspace, g::ClearSky.SemiGrayGas, T, P) =
    check(ccall((:cs_sigma_add_gray, LIB), Int32, (Ptr{Cvoid}, Float64, Float64), w.h, Float64(g.σ), g.νcut))
# user functions σ(ν,T,P) cannot cross the ABI: pre-evaluate on the host, node-major
function addto!(w::Workspace, f::Function, ν, T, P)
    σ = Float64[f(x, t, p) for x in ν, (t, p) in zip(T, P)]     # [nν, nnode] column-major == C [nnode][nν]
    check(ccall((:cs_sigma_add_host, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), w.h, σ))
end
# (CIATables -> cs_cia_upload + cs_sigma_add_cia follows the same pattern; see INTEGRATION.md)

# ---- S3: the numerical core -----------------------------------------------------------------------------
struct B200Discretized <: AbstractNumericalCore
    nstream::Int64
    nlobatto::Int64
end
B200Discretized(; nstream::Int=5, nlobatto::Int=2) = B200Discretized(nstream, nlobatto)

function ClearSky.monochromaticfluxes!(M⁺::AbstractMatrix, M⁻::AbstractMatrix, τ::AbstractMatrix,
                                       core::B200Discretized, P::AbstractVector{<:Real}, g::Real, T, μ, 𝒻S, 𝒻a,
                                       absorbers...; θₛ::Real=0.841)::Nothing
    F = fluxes_b200(core, P, g, T, μ, 𝒻S, 𝒻a, absorbers...; θₛ=θₛ, M⁺=M⁺, M⁻=M⁻, τ=τ)
    nothing
end

# fused path: never materialises M⁺/M⁻/τ unless asked (radiate!/fluxes only need F⁺, F⁻, Fnet)
function fluxes_b200(core::B200Discretized, P, g, T, μ, 𝒻S, 𝒻a, absorbers...; θₛ=0.841, M⁺=nothing, M⁻=nothing, τ=nothing)
    gases = filter(a -> a isa AbstractGas, collect(absorbers))
    ν = gases[1].ν
    𝒻T, 𝒻μ = formprofiles(P, T, μ)
    @assert issorted(P) "pressure coordinates must be in ascending order (sorted)"
    checkstreams(core.nstream); checkazimuth(θₛ)
    Tl, μl = lobattoevaluations(P, 𝒻T, 𝒻μ, core.nlobatto)          # [nlobatto, np-1] exactly as the ABI wants
    𝓍, 𝓌 = lobattonodes(core.nlobatto)
    𝓂, 𝒲 = streamnodes(core.nstream)
    np, nl = length(P), core.nlobatto
    # unique nodes, ascending pressure: node n of layer i at n + (nl-1)*(i-1)
    Pn = Float64[P[1]]; Tn = Float64[Tl[1, 1]]
    for i in 1:np-1, n in 2:nl
        push!(Pn, n == nl ? P[i+1] : P[i] + (P[i+1] - P[i]) * 𝓍[n]); push!(Tn, Tl[n, i])
    end
    w = Workspace(ν, length(Pn))
    for a in absorbers
        a isa Function ? addto!(w, a, ν, Tn, Pn) : addto!(w, a, Tn, Pn)
    end
    Tlev = Float64[𝒻T(p) for p in P]
    F⁺, F⁻, Fnet = zeros(np), zeros(np), zeros(np)
    ptr(x) = x === nothing ? Ptr{Float64}(C_NULL) : pointer(x)
    GC.@preserve M⁺ M⁻ τ check(ccall((:cs_fluxes, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64},
         Ptr{Float64}, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        w.h, np, collect(Float64, P), nl, 𝓌, μl, Tlev, Float64(g), Float64[𝒻S(x) for x in ν], Float64[𝒻a(x) for x in ν],
        Float64(θₛ), core.nstream, 𝓂, 𝒲, C_NULL, ptr(τ), ptr(M⁺), ptr(M⁻), F⁺, F⁻, Fnet))
    (F⁺, F⁻, Fnet)
end

function ClearSky.radiate!(F::FluxPack, core::B200Discretized, P::AbstractVector{<:Real}, g::Real, T, μ, 𝒻S, 𝒻a,
                           absorbers...; kwargs...)::Nothing
    F⁺, F⁻, Fnet = fluxes_b200(core, P, g, T, μ, 𝒻S, 𝒻a, absorbers...; M⁺=F.M⁺, M⁻=F.M⁻, τ=F.τ, kwargs...)
    F.F⁺ .= F⁺; F.F⁻ .= F⁻; F.Fnet .= Fnet
    nothing
end

# ---- Radau-core equivalents (src/fluxes.jl:39-66, 133-158): same entry points on the Discretized GPU core, layers
# equally spaced in ln P doubled until two successive Richardson extrapolates agree to tol (see clearsky_b200/radau.py, the
# executed twin)
function outgoing_b200(Pₛ::Real, g::Real, 𝒻T, 𝒻μ, absorbers...; Ptop::Real=1.0, nstream::Int=5, tol::Real=1e-5)
    taufloor!(1e-9)
    try
        prev, prevE, n = nothing, nothing, 32
        while true
            P = exp.(range(log(Ptop), log(Pₛ), length=n + 1))
            ν = first(a for a in absorbers if a isa AbstractGas).ν
            M⁺, M⁻ = zeros(n + 1, length(ν)), zeros(n + 1, length(ν))
            fluxes_b200(B200Discretized(nstream, 4), P, g, 𝒻T, 𝒻μ, x -> 0.0, x -> 0.0, absorbers...; M⁺=M⁺, M⁻=M⁻)
            olr = M⁺[1, :]
            E = prev === nothing ? nothing : (4 .* olr .- prev) ./ 3          # Richardson extrapolate of the O(n⁻²) scheme
            if prevE !== nothing
                scale = max.(abs.(E), 1e-3 * maximum(abs.(E)))
                (maximum(abs.(E .- prevE) ./ scale) < tol || 2n + 1 > 1025) && return E
            end
            prev, prevE, n = olr, E, 2n
        end
    finally
        taufloor!(1e-6)
    end
end

function opticaldepth_b200(P₁::Real, P₂::Real, g::Real, 𝒻T, 𝒻μ, θ::Real, absorbers...; tol::Real=1e-5)
    P₁, P₂ = max(P₁, P₂), min(P₁, P₂)
    prev, n = nothing, 16
    while true
        τ = ClearSky.opticaldepth(exp.(range(log(P₂), log(P₁), length=n + 1)), g, 𝒻T, 𝒻μ, θ, absorbers...; nlobatto=4)
        prev !== nothing && (maximum(abs.(τ .- prev) ./ max.(abs.(τ), 1e-3 * maximum(abs.(τ)))) < tol || 2n + 1 > 1025) && return τ
        prev, n = τ, 2n
    end
end

end # module
