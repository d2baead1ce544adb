#!/usr/bin/env julia
#=
Pin recipe for the CPU oracle (SURVEY.md section 8c): run the UNMODIFIED reference once, anywhere Julia is available,
and write golden vectors that tests/test_reference_golden.py compares with oracle/oracle.c.

    julia --project=/path/to/ClearSky.jl tools/julia_golden.jl [outdir = tests/golden/ref]

Needs only the reference package (and its dependencies, Faddeyeva985 among them) and `gunzip` on PATH.  Inputs: the
reference's own fixtures (test/HITRAN/CO2.par, CO2-CO2_2018.cia) and the committed synthetic slice of BASELINE configs[1]
(tests/golden/ref_inputs/, written by tools/julia_golden_inputs.py).  Output: one .npy file per array (NumPy format 1.0,
column-major, written by hand below: no Julia dependency beyond ClearSky itself) plus MANIFEST.txt with the package
versions that produced them.  This script has NOT been executed in the build container (no Julia there); it only uses
exported reference functions with the signatures cited beside each call.
=#
using ClearSky
using Pkg

const REPO = normpath(joinpath(@__DIR__, ".."))
const OUT = length(ARGS) >= 1 ? ARGS[1] : joinpath(REPO, "tests", "golden", "ref")
const REFDATA = joinpath(pkgdir(ClearSky), "test", "HITRAN")
mkpath(OUT)

# ---- .npy writer (format 1.0; fortran_order so Julia's column-major memory is written as is)
npydescr(::Type{Float64}) = "<f8"
npydescr(::Type{Int64}) = "<i8"
function writenpy(name::String, A::AbstractArray{T}) where {T}
    A = Array(A)
    shape = ndims(A) == 1 ? "($(length(A)),)" : "(" * join(size(A), ", ") * ")"
    hdr = "{'descr': '$(npydescr(T))', 'fortran_order': True, 'shape': $shape, }"
    pad = (64 - (10 + length(hdr) + 1) % 64) % 64
    hdr = hdr * " "^pad * "\n"
    open(joinpath(OUT, name * ".npy"), "w") do io
        write(io, UInt8[0x93, 0x4e, 0x55, 0x4d, 0x50, 0x59, 0x01, 0x00])        # \x93NUMPY 1.0
        write(io, htol(UInt16(length(hdr))))
        write(io, hdr)
        write(io, htol.(A))
    end
    println("wrote ", name, " ", size(A))
end

function gunzipped(path::String)
    tmp = joinpath(mktempdir(), replace(basename(path), ".gz" => ""))
    run(pipeline(`gunzip -c $path`, stdout=tmp))
    tmp
end

# ===================================================================================================================
# 1. faddeyeva(x, y) as fvoigt calls it (src/absorption/line_shapes.jl:375), on points straddling every border of the
#    BOTH candidate region maps of the oracle (oracle/oracle.c: map 1 |z|^2 = 2.5, 30, 62, 256, 3.8e4; y^2 = 1e-13, 0.072 and
#    map 0 |z|^2 = 3.5, 28.5, 107, 160, 1.6e4; y^2 = 6e-14, 0.026) and on a log grid: one run decides which map the package uses
# ===================================================================================================================
let
    xs, ys = Float64[], Float64[]
    offs = [-1e-1, -1e-3, -1e-6, -1e-9, -1e-12, 0.0, 1e-12, 1e-9, 1e-6, 1e-3, 1e-1]
    for s in (2.5, 3.5, 28.5, 30.0, 62.0, 107.0, 160.0, 256.0, 1.6e4, 3.8e4), f in offs, φ in range(0.0, π / 2, length=33)
        r = sqrt(s * (1 + f))
        push!(xs, r * cos(φ)); push!(ys, r * sin(φ))
    end
    for y2 in (6e-14, 1e-13, 0.026, 0.072), f in offs, x in vcat(0.0, 10 .^ range(-3, 5, length=49))
        push!(xs, x); push!(ys, sqrt(y2 * (1 + f)))
    end
    for x in vcat(0.0, 10 .^ range(-6, 5, length=56)), y in 10 .^ range(-30, 5, length=71)
        push!(xs, x); push!(ys, y)
    end
    w = [ClearSky.faddeyeva(xs[i], ys[i]) for i in eachindex(xs)]
    writenpy("fad_x", xs); writenpy("fad_y", ys); writenpy("fad_w", Float64.(real.(w)))
end

# ===================================================================================================================
# 2. BASELINE configs[0] (C1): reference fixture CO2.par, nu_i = 1 + 2.5(i-1), 21 levels on a dry adiabat
# ===================================================================================================================
co2 = SpectralLines(joinpath(REFDATA, "CO2.par"), progress=false)               # par.jl:286
ν = collect(1.0 .+ 2.5 .* (0:999))
P = pressuregrid(10.0, 1e5, 21)                                                 # util.jl:19
Γ = DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)                           # atmospherics.jl:321
T = Γ.(P)
C = 400e-6
writenpy("c1_nu", ν); writenpy("c1_P", P); writenpy("c1_T", T)

# per-line quantities at one (T, P) (line_shapes.jl:125-132, 146-148, 259-261)
let idx = collect(1:co2.N), Tq = 250.0, Pq = 5e4
    writenpy("c1_line_S", scaleintensity(co2, idx, Tq))
    writenpy("c1_line_alpha", αdoppler(co2, idx, Tq))
    writenpy("c1_line_gamma", γlorentz(co2, idx, Tq, Pq, C * Pq))
end

# the four in-place shapes (line_shapes.jl:200, 313, 412, 527)
function shapeblock(shape!, sl, ν, T, P, Pₚ, Δνcut)
    σ = zeros(length(ν), length(T))
    for k in eachindex(T)
        shape!(view(σ, :, k), ν, sl, T[k], P[k], Pₚ[k], Δνcut)
    end
    σ
end
lev5, lev10 = collect(1:5:21), collect(1:10:21)
writenpy("c1_sigma_voigt", shapeblock(voigt!, co2, ν, T[lev5], P[lev5], C .* P[lev5], 25.0))
writenpy("c1_sigma_lorentz", shapeblock(lorentz!, co2, ν, T[lev5], P[lev5], C .* P[lev5], 25.0))
writenpy("c1_sigma_doppler", shapeblock(doppler!, co2, ν, T[lev5], P[lev5], C .* P[lev5], 25.0))
writenpy("c1_sigma_phco2", shapeblock(PHCO2!, co2, ν, T[lev10], P[lev10], P[lev10], 500.0))
# a fine grid across the 15 micron Q branch, where the near-centre Faddeyeva branches are exercised at low pressure
νq = collect(range(667.0, 668.0, length=2001))
writenpy("c1q_nu", νq)
writenpy("c1q_sigma_voigt", shapeblock(voigt!, co2, νq, T[lev5], P[lev5], C .* P[lev5], 25.0))

# bake + OpacityTable on a 6 x 8 domain (gases.jl:26-61, 68-85, 97-145): table values at the nodes and at the C1 levels
Ω = AtmosphericDomain((140.0, 300.0), 6, (5.0, 1.1e5), 8)
gas = Gas(co2, (T, P) -> C, ν, Ω, voigt!, 25.0, progress=false)                 # gases.jl:225
writenpy("c1_tab_Tnodes", Ω.T); writenpy("c1_tab_Pnodes", Ω.P)
writenpy("c1_tab_nodes", [rawσ(gas, i, Ω.T[j], Ω.P[k]) for i in 1:length(ν), j in 1:Ω.nT, k in 1:Ω.nP])   # gases.jl:256
writenpy("c1_tab_levels", [rawσ(gas, i, T[k], P[k]) for i in 1:length(ν), k in 1:length(P)])

# fluxes through the 12 x 24 table of configs[0] (fluxes.jl:311-340), Discretized(5, 2), no star, black surface
Ω1 = AtmosphericDomain((140.0, 300.0), 12, (5.0, 1.1e5), 24)
gas1 = Gas(co2, (T, P) -> C, ν, Ω1, voigt!, 25.0, progress=false)
F⁺, F⁻ = fluxes(P, 9.8, Γ, 0.029, ν -> 0.0, ν -> 0.0, gas1; core=Discretized(5, 2))
writenpy("c1_tab_Fup", F⁺); writenpy("c1_tab_Fdn", F⁻)
writenpy("c1_tab12_levels", [rawσ(gas1, i, T[k], P[k]) for i in 1:length(ν), k in 1:length(P)])
# same with a stellar beam and a reflecting surface, 3 Lobatto nodes, 4 streams
F⁺, F⁻ = fluxes(P, 9.8, Γ, 0.029, ν -> 1e-3, ν -> 0.3, gas1; core=Discretized(4, 3), θₛ=0.5)
writenpy("c1_tab_Fup_sun", F⁺); writenpy("c1_tab_Fdn_sun", F⁻)
# monochromatic fluxes and layer depths (fluxes.jl:281-306)
M⁺, M⁻ = monochromaticfluxes(P, 9.8, Γ, 0.029, ν -> 0.0, ν -> 0.0, gas1; core=Discretized(5, 2))
writenpy("c1_tab_Mup", M⁺); writenpy("c1_tab_Mdn", M⁻)
# optical depth of the column (fluxes.jl:68-97; default nlobatto = 4)
writenpy("c1_tab_depth", opticaldepth(P, 9.8, Γ, 0.029, 0.0, gas1))

# CIA (collision_induced_absorption.jl:295-303, 378-382): CO2-CO2 on a pure-CO2 column, both extrapolate settings
let νc = collect(range(1.0, 3000.0, length=1500))
    writenpy("cia_nu", νc)
    for (tag, ex) in (("ex", true), ("noex", false))
        x = CIATables(joinpath(REFDATA, "CO2-CO2_2018.cia"), extrapolate=ex, verbose=false)
        writenpy("cia_sigma_" * tag, [cia(νc[i], x, T[k], P[k], P[k], P[k]) for i in eachindex(νc), k in eachindex(P)])
    end
end

# ===================================================================================================================
# 3. BASELINE configs[1] (C2) on a 1500-point slice of its grid: 2 x ~5500 synthetic Voigt lines, 101 levels
# ===================================================================================================================
let
    inp = joinpath(REPO, "tests", "golden", "ref_inputs")
    i0, n = 149250, 1500                      # tests/golden/ref_inputs/c2slice.json
    ν2 = collect(0.01 .* ((i0 + 1):(i0 + n)))
    P2 = pressuregrid(10.0, 1e5, 101)
    T2 = Γ.(P2)
    writenpy("c2_nu", ν2); writenpy("c2_P", P2); writenpy("c2_T", T2)
    levs = collect(1:10:101)
    for (name, Cg) in (("CO2", 400e-6), ("H2O", 1e-3))
        sl = SpectralLines(gunzipped(joinpath(inp, "c2slice_$(name).par.gz")), progress=false)
        writenpy("c2_lines_nu_" * name, sl.ν); writenpy("c2_lines_S_" * name, sl.S)
        writenpy("c2_sigma_voigt_" * name, shapeblock(voigt!, sl, ν2, T2[levs], P2[levs], Cg .* P2[levs], 25.0))
    end
end

open(joinpath(OUT, "MANIFEST.txt"), "w") do io
    println(io, "generated by tools/julia_golden.jl with Julia ", VERSION)
    for (_, info) in Pkg.dependencies()
        info.name in ("ClearSky", "Faddeyeva985", "BasicInterpolators", "FastGaussQuadrature") &&
            println(io, info.name, " ", info.version, " ", something(info.git_revision, ""), " ", something(info.tree_hash, ""))
    end
end
println("done: ", OUT)
