"""One-shot API wrappers on the Discretized core (seam S3 of SURVEY.md section 8b).

Reference: src/fluxes.jl -- opticaldepth(P::Vector...) :68-97, monochromaticfluxes!(…Discretized…) :238-279,
monochromaticfluxes :281-306, fluxes :311-340, netfluxes :342-352, radiate! :357-383, radiate :385-403.
Profile callables (𝒻T, 𝒻μ, 𝒻S, 𝒻a, fC) stay host-side and are pre-evaluated to arrays exactly where the
reference does it (lobattoevaluations core/discretized.jl:11-30, planckevaluations :46-58).
"""
import warnings

import numpy as np

from ._lib import check, f64, lib, ptr
from .absorbers import SigmaWorkspace, unifyabsorbers
from .core import Discretized, FluxPack
from .quadrature import lobattonodes, streamnodes
from .util import AtmosphericProfile


def checkazimuth(θ):
    """fluxes.jl:4-6"""
    assert 0 <= θ < np.pi / 2, "azimuth angle θ must be ∈ [0,π/2)"


def checkstreams(n):
    """fluxes.jl:8-10"""
    if n < 4:
        warnings.warn("careful! using nstream < 4 is likely to be inaccurate!")


def formprofile(P, x):
    """fluxes.jl:13-16: vector -> AtmosphericProfile, number -> constant function, callable -> as is"""
    if callable(x):
        return x
    if np.ndim(x) == 0:
        c = float(x)
        return lambda *_: c
    return AtmosphericProfile(P, x)


def _vec(f, *args):
    """evaluate a profile callable on arrays when it supports them, element by element otherwise"""
    shape = np.shape(args[0])
    try:
        out = np.asarray(f(*args), dtype=np.float64)
        if out.shape == shape:
            return out
        if out.shape == ():
            return np.full(shape, float(out))
    except Exception:
        pass
    flat = [np.ravel(a) for a in args]
    return np.array([float(f(*xs)) for xs in zip(*flat)], dtype=np.float64).reshape(shape)


def lobattoevaluations(P, fT, fμ, nlobatto):
    """core/discretized.jl:11-30 -> T, μ [nlobatto, np-1] (returned as C arrays [np-1, nlobatto], i.e. the same
    memory order as Julia's column-major [nlobatto, np-1]) plus the node pressures"""
    x, _ = lobattonodes(nlobatto)
    P = np.asarray(P, dtype=np.float64)
    ΔP = P[1:] - P[:-1]
    Pn = P[:-1, None] + ΔP[:, None] * x[None, :]
    T = _vec(fT, Pn)
    μ = _vec(fμ, T, Pn)
    return T, μ, Pn


def _unique_nodes(P, T, Pn, nlobatto):
    """the (np-1)*(nlobatto-1)+1 distinct quadrature nodes in ascending pressure: node n of layer i sits at
    index n + (nlobatto-1)*i; the shared end node uses the exact level pressure P[i+1] and T[end,i]
    (core/discretized.jl:169)."""
    L = len(P) - 1
    nn = L * (nlobatto - 1) + 1
    Tn, Pq = np.empty(nn), np.empty(nn)
    Tn[0], Pq[0] = T[0, 0], P[0]
    for i in range(L):
        for n in range(1, nlobatto - 1):
            Tn[n + (nlobatto - 1) * i] = T[i, n]
            Pq[n + (nlobatto - 1) * i] = Pn[i, n]
        Tn[(nlobatto - 1) * (i + 1)] = T[i, nlobatto - 1]
        Pq[(nlobatto - 1) * (i + 1)] = P[i + 1]
    return Tn, Pq


def _prepare(P, T, μ, absorbers, nlobatto):
    A, ν, nν = unifyabsorbers(absorbers)
    P = f64(np.asarray(P, dtype=np.float64))
    fT, fμ = formprofile(P, T), formprofile(P, μ)
    Tl, μl, Pn = lobattoevaluations(P, fT, fμ, nlobatto)
    Tn, Pq = _unique_nodes(P, Tl, Pn, nlobatto)
    # one device workspace per (absorber, node count), reused across calls (the RCM loop calls this every step)
    cache = A.__dict__.setdefault("_ws_cache", {})
    ws = cache.get(len(Tn))
    if ws is None:
        ws = cache[len(Tn)] = SigmaWorkspace(ν, len(Tn))
    else:
        ws.zero()
    A.sigma_nodes(ws, Tn, Pq)
    return A, ν, nν, P, fT, μl, ws


def opticaldepth(P, *args, **kwargs):
    """opticaldepth(P::Vector, g, T, μ, θ, absorbers...; nlobatto=4)    -- fluxes.jl:68-97 (Discretized core)
    opticaldepth(P₁::Real, P₂::Real, g, 𝒻T, 𝒻μ, θ, absorbers...; tol=1e-5) -- fluxes.jl:39-66 (Radau equivalent, radau.py)"""
    if np.ndim(P) == 0:
        from .radau import opticaldepth_between
        return opticaldepth_between(P, *args, **kwargs)
    return _opticaldepth_levels(P, *args, **kwargs)


def _opticaldepth_levels(P, g, T, μ, θ, *absorbers, nlobatto=4):
    P = np.sort(np.asarray(P, dtype=np.float64))
    A, ν, nν, P, fT, μl, ws = _prepare(P, T, μ, absorbers, nlobatto)
    A.checkpressures(P[-1], P[0])
    checkazimuth(θ)
    _, w = lobattonodes(nlobatto)
    τ = np.empty(nν)
    check(lib().cs_opticaldepth(ws.h, len(P), ptr(P), nlobatto, ptr(f64(w)), ptr(f64(μl)), float(g), float(θ), ptr(τ)))
    return τ


def transmittance(*args, **kwargs):
    """fluxes.jl:109"""
    return np.exp(-opticaldepth(*args, **kwargs))


def _eval_spectral(f, ν):
    if f is None:
        return None
    if callable(f):
        try:
            v = np.asarray(f(ν), dtype=np.float64)
            if v.shape == ν.shape:
                return f64(v)
        except Exception:
            pass
        return f64(np.array([f(x) for x in ν], dtype=np.float64))
    return f64(np.full(len(ν), float(f)))


def monochromaticfluxes_(Mup, Mdn, τ, core, P, g, T, μ, fS, fa, *absorbers, θs=0.841, _F=None, ν_weights=None):
    """monochromaticfluxes!(M⁺, M⁻, τ, core::Discretized, P, g, T, μ, 𝒻S, 𝒻a, absorbers...; θₛ) -- fluxes.jl:238-279.
    Any of Mup/Mdn/τ may be None (not materialised).  Returns (F⁺, F⁻, Fnet) since the spectral integral is fused."""
    if not isinstance(core, Discretized):
        from .radau import Radau, monochromaticfluxes_radau
        assert isinstance(core, Radau), "core must be Discretized or Radau"
        assert _F is None and ν_weights is None
        return monochromaticfluxes_radau(Mup, Mdn, τ, core, P, g, T, μ, fS, fa, *absorbers, θs=θs)
    nstream, nlobatto = core.nstream, core.nlobatto
    assert np.all(np.diff(P) >= 0), "pressure coordinates must be in ascending order (sorted)"
    A, ν, nν, P, fT, μl, ws = _prepare(P, T, μ, absorbers, nlobatto)
    Tlev = f64(_vec(fT, P))
    A.checkpressures(P[-1], P[0])
    checkstreams(nstream)
    checkazimuth(θs)
    m, W = streamnodes(nstream)
    _, w = lobattonodes(nlobatto)
    fSν, faν = _eval_spectral(fS, ν), _eval_spectral(fa, ν)
    npl = len(P)
    Fup, Fdn, Fnet = np.empty(npl), np.empty(npl), np.empty(npl)
    for arr, shp in ((Mup, (nν, npl)), (Mdn, (nν, npl)), (τ, (nν, npl - 1))):
        if arr is not None:
            assert arr.shape == shp and arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
    check(lib().cs_fluxes(ws.h, npl, ptr(P), nlobatto, ptr(f64(w)), ptr(f64(μl)), ptr(Tlev), float(g),
                          ptr(fSν), ptr(faν), float(θs), nstream, ptr(f64(m)), ptr(f64(W)),
                          ptr(f64(ν_weights)) if ν_weights is not None else None,
                          ptr(τ), ptr(Mup), ptr(Mdn), ptr(Fup), ptr(Fdn), ptr(Fnet)))
    return Fup, Fdn, Fnet


def monochromaticfluxes(P, g, T, μ, fS, fa, *absorbers, core=None, θs=0.841):
    """fluxes.jl:281-306 -> (M⁺, M⁻) as [nν, np] arrays"""
    core = core or Discretized()
    _, ν, nν = unifyabsorbers(absorbers)
    Mup, Mdn = np.zeros((nν, len(P))), np.zeros((nν, len(P)))
    monochromaticfluxes_(Mup, Mdn, None, core, P, g, T, μ, fS, fa, *absorbers, θs=θs)
    return Mup, Mdn


def fluxes(P, g, T, μ, fS, fa, *absorbers, core=None, θs=0.841):
    """fluxes(P, g, T, μ, 𝒻S, 𝒻a, absorbers...; core, θₛ) -> (F⁺, F⁻) -- fluxes.jl:311-340.
    The monochromatic blocks are never materialised: K6 reduces spectrally on the fly."""
    core = core or Discretized()
    Fup, Fdn, _ = monochromaticfluxes_(None, None, None, core, P, g, T, μ, fS, fa, *absorbers, θs=θs)
    return Fup, Fdn


def fluxes_batch(P, g, Ts, μ, fS, fa, A, core=None, θs=0.841):
    """fluxes for several temperature profiles `Ts` (callables or vectors on P) over ONE AcceleratedAbsorber, whose Σ does
    not depend on T (absorbers.jl:203): all layer depths and transmittances are shared (cs_fluxes_batch).  This is what
    jacobian! (radiative_convective.jl:154-171) amounts to.  -> F⁺[nbatch, np], F⁻[nbatch, np]"""
    from .absorbers import AcceleratedAbsorber
    assert isinstance(A, AcceleratedAbsorber), "batched fluxes need an absorber whose Σ ignores T (AcceleratedAbsorber)"
    core = core or Discretized()
    P = f64(np.asarray(P, dtype=np.float64))
    assert np.all(np.diff(P) >= 0), "pressure coordinates must be in ascending order (sorted)"
    A.checkpressures(P[-1], P[0])
    checkstreams(core.nstream)
    checkazimuth(θs)
    fTs = [formprofile(P, T) for T in Ts]
    _, ν, nν, P, _, μl, ws = _prepare(P, fTs[0], μ, (A,), core.nlobatto)      # Σ and μ at the nodes (T-independent here)
    Tlev = f64(np.stack([_vec(fT, P) for fT in fTs]))
    m, W = streamnodes(core.nstream)
    _, w = lobattonodes(core.nlobatto)
    fSν, faν = _eval_spectral(fS, ν), _eval_spectral(fa, ν)
    F = np.empty((len(fTs), 2 * len(P)))
    check(lib().cs_fluxes_batch(ws.h, len(P), ptr(P), core.nlobatto, ptr(f64(w)), ptr(f64(μl)), len(fTs), ptr(Tlev), float(g),
                                ptr(fSν), ptr(faν), float(θs), core.nstream, ptr(f64(m)), ptr(f64(W)), None, ptr(F)))
    return F[:, :len(P)].copy(), F[:, len(P):].copy()


def netfluxes(P, g, T, μ, fS, fa, *absorbers, **kwargs):
    """fluxes.jl:342-352"""
    Fup, Fdn = fluxes(P, g, T, μ, fS, fa, *absorbers, **kwargs)
    return Fup - Fdn


def radiate_(F, core, P, g, T, μ, fS, fa, *absorbers, θs=0.841, materialize=True):
    """radiate!(F::FluxPack, core, P, g, T, μ, 𝒻S, 𝒻a, absorbers...) -- fluxes.jl:357-383.
    materialize=False skips the D2H of τ/M⁺/M⁻ (the RCM loop only consumes Fnet)."""
    _, ν, nν = unifyabsorbers(absorbers)
    assert F.size == (len(P), nν), "size of FluxPack does not match number of pressure or wavenumber coordinates"
    if materialize:
        Fup, Fdn, Fnet = monochromaticfluxes_(F.Mup, F.Mdn, F.τ, core, P, g, T, μ, fS, fa, *absorbers, θs=θs)
    else:
        Fup, Fdn, Fnet = monochromaticfluxes_(None, None, None, core, P, g, T, μ, fS, fa, *absorbers, θs=θs)
    F.Fup[:], F.Fdn[:], F.Fnet[:] = Fup, Fdn, Fnet
    return None


def radiate(P, g, T, μ, fS, fa, *absorbers, core=None, θs=0.841):
    """fluxes.jl:385-403"""
    core = core or Discretized()
    _, ν, nν = unifyabsorbers(absorbers)
    F = FluxPack(len(P), nν)
    radiate_(F, core, P, g, T, μ, fS, fa, *absorbers, θs=θs)
    return F
