"""Radau-core equivalents.

The reference's `Radau` core integrates dτ/dP and the Schwarzschild equation per wavenumber with an adaptive
implicit ODE solver (ScalarRadau.jl, third party) to a tolerance `tol` (src/fluxes.jl:39-66 `opticaldepth(P₁,P₂,…)`,
:133-158 `outgoing(Pₛ,…)`, :197-236 `monochromaticfluxes!(…, core::Radau, …)`; src/core/radau.jl).  An adaptive
per-wavenumber step sequence has no place on a GPU and cannot be matched beyond its own tolerance, so the same entry
points are provided here on the B200 Discretized core (K6): the pressure range is divided into layers equally spaced
in ln P, every layer's optical depth is a 4-point Gauss-Lobatto integral, and the number of layers n is doubled until two
successive Richardson extrapolates E(n) = (4 F(2n) - F(n))/3 of the second-order scheme (linear-in-τ source function;
measured error ratio 4.0 per doubling) agree to `tol` (relative, per wavenumber, against a floor of 1e-3 of the spectral
maximum); the last extrapolate is returned.  One flux call holds at most MAX_LEVELS levels (per-CTA tables of K6 live in
shared memory).
The layer-depth floor of the Discretized core (1e-6, discretized.jl:174) is lowered to 1e-9 for these calls so that
hundreds of thin layers do not add opacity in spectral windows.

These are convenience entry points, not parity targets: they agree with the reference's Radau results to the
integrator tolerance, not to 1e-8.  `outgoing(P::Vector, …)` is the reference's Discretized method (fluxes.jl:160-192),
which cannot run in the reference (𝒹streams uses τ before defining it, discretized.jl:196); it is provided with its
evident meaning: upward streams from a black surface at 𝒻T(Pₛ), no stellar term.
"""
import warnings
from dataclasses import dataclass

import numpy as np

from ._lib import check, f64, lib, ptr
from .absorbers import SigmaWorkspace, unifyabsorbers
from .core import AbstractNumericalCore, Discretized
from .fluxes import (_eval_spectral, _unique_nodes, _vec, checkazimuth, checkstreams, formprofile, lobattoevaluations)
from .quadrature import lobattonodes, streamnodes

RADAU_TAU_FLOOR = 1e-9
REFERENCE_TAU_FLOOR = 1e-6      # discretized.jl:174
MAX_NODE_BYTES = 6 << 30        # cap on the Σ workspace of one refinement level
MAX_LEVELS = 4097               # K6 keeps (2 + nlobatto)·np doubles of per-level tables in shared memory (its per-warp
                                # accumulators move to global memory beyond ~600 levels)


@dataclass
class Radau(AbstractNumericalCore):
    """Radau(; nstream=5, tol=1e-5) -- shared.jl:40-47"""
    nstream: int = 5
    tol: float = 1e-5


def _sigma(A, ν, P, fT, fμ, nlobatto):
    """Σ at the distinct Lobatto nodes of the level vector P (ascending) in a throw-away workspace"""
    Tl, μl, Pn = lobattoevaluations(P, fT, fμ, nlobatto)
    Tn, Pq = _unique_nodes(P, Tl, Pn, nlobatto)
    ws = SigmaWorkspace(ν, len(Tn))
    A.sigma_nodes(ws, Tn, Pq)
    return ws, f64(μl)


def _fits(nν, nlev, nlobatto):
    return (nlev <= MAX_LEVELS and 8 * ((2 + nlobatto) * nlev + 64) <= 200 * 1024
            and nν * ((nlev - 1) * (nlobatto - 1) + 1) * 8 <= MAX_NODE_BYTES)


def _converged(new, old, tol):
    """do two successive results agree to tol (relative, against a floor of 1e-3 of the maximum)?"""
    scale = np.maximum(np.abs(new), 1e-3 * np.max(np.abs(new)) + np.finfo(float).tiny)
    return float(np.max(np.abs(new - old) / scale)) < tol


def _extrapolate(new, old):
    return (4.0 * new - old) / 3.0


class _floor:
    """lower the layer-depth floor of a context for the duration of a Radau-equivalent call"""

    def __init__(self, ctx):
        self.ctx = ctx

    def __enter__(self):
        self.ctx.set_tau_floor(RADAU_TAU_FLOOR)

    def __exit__(self, *a):
        self.ctx.set_tau_floor(REFERENCE_TAU_FLOOR)


def opticaldepth_between(P1, P2, g, fT, fμ, θ, *absorbers, tol=1e-5, nlobatto=4, n0=16):
    """opticaldepth(P₁, P₂, g, 𝒻T, 𝒻μ, θ, absorbers...; tol=1e-5) -- fluxes.jl:39-66: optical depth between two
    pressure levels along a path at angle θ, for every wavenumber of the absorbers"""
    A, ν, nν = unifyabsorbers(absorbers)
    P1, P2 = max(float(P1), float(P2)), min(float(P1), float(P2))
    A.checkpressures(P1, P2)
    checkazimuth(θ)
    fT, fμ = formprofile(None, fT), formprofile(None, fμ)
    _, w = lobattonodes(nlobatto)
    prev, n = None, n0
    while True:
        P = f64(np.exp(np.linspace(np.log(P2), np.log(P1), n + 1)))
        P[0], P[-1] = P2, P1
        ws, μl = _sigma(A, ν, P, fT, fμ, nlobatto)
        τ = np.empty(nν)
        check(lib().cs_opticaldepth(ws.h, len(P), ptr(P), nlobatto, ptr(f64(w)), ptr(μl), float(g), float(θ), ptr(τ)))
        ws.close()
        # the Lobatto rule integrates dτ/dP to high order: no extrapolation, plain successive difference
        if prev is not None and _converged(τ, prev, tol):
            return τ
        if not _fits(nν, 2 * n + 1, nlobatto):
            warnings.warn(f"opticaldepth: refinement stopped at {n} layers before reaching tol = {tol}")
            return τ
        prev, n = τ, 2 * n


def _sweep(A, ν, P, g, fT, fμ, fSν, faν, θs, nstream, nlobatto, want_M):
    """one Discretized solve on the level vector P (ascending); returns (M⁺, M⁻) [nν, np] when want_M"""
    ws, μl = _sigma(A, ν, P, fT, fμ, nlobatto)
    Tlev = f64(_vec(fT, P))
    m, W = streamnodes(nstream)
    _, w = lobattonodes(nlobatto)
    npl, nν = len(P), len(ν)
    Mup = np.empty((nν, npl)) if want_M else None
    Mdn = np.empty((nν, npl)) if want_M else None
    Fup, Fdn, Fnet = np.empty(npl), np.empty(npl), np.empty(npl)
    with _floor(ws.ctx):
        check(lib().cs_fluxes(ws.h, npl, ptr(P), nlobatto, ptr(f64(w)), ptr(μl), ptr(Tlev), float(g), ptr(fSν), ptr(faν),
                              float(θs), nstream, ptr(f64(m)), ptr(f64(W)), None, None, ptr(Mup), ptr(Mdn), ptr(Fup),
                              ptr(Fdn), ptr(Fnet)))
    ws.close()
    return Mup, Mdn, Fup, Fdn


def outgoing(P, g, T, μ, *absorbers, Ptop=1.0, nstream=5, tol=1e-5, nlobatto=None, n0=32):
    """outgoing monochromatic fluxes [W/m²/cm⁻¹] at the top of the atmosphere, one value per wavenumber.

    outgoing(Pₛ::Real, g, 𝒻T, 𝒻μ, absorbers...; Ptop=1.0, nstream=5, tol=1e-5)   -- fluxes.jl:133-158 (Radau)
    outgoing(P::Vector, g, T, μ, absorbers...; nstream=5, nlobatto=3)              -- fluxes.jl:160-192 (Discretized)

    No stellar term and a black surface at 𝒻T(Pₛ) in both forms; total OLR = trapz(ν, outgoing(...))."""
    A, ν, nν = unifyabsorbers(absorbers)
    checkstreams(nstream)
    if np.ndim(P) > 0:
        # Discretized method: any order of P is accepted (the reference sorts descending; K6 wants ascending)
        P = f64(np.sort(np.asarray(P, dtype=np.float64)))
        nlob = 3 if nlobatto is None else int(nlobatto)
        fT, fμ = formprofile(P, T), formprofile(P, μ)
        A.checkpressures(P[-1], P[0])
        ws, μl = _sigma(A, ν, P, fT, fμ, nlob)
        Tlev = f64(_vec(fT, P))
        m, W = streamnodes(nstream)
        _, w = lobattonodes(nlob)
        npl = len(P)
        Mup, Mdn = np.empty((nν, npl)), np.empty((nν, npl))
        Fup, Fdn, Fnet = np.empty(npl), np.empty(npl), np.empty(npl)
        check(lib().cs_fluxes(ws.h, npl, ptr(P), nlob, ptr(f64(w)), ptr(μl), ptr(Tlev), float(g), None, None, 0.841,
                              nstream, ptr(f64(m)), ptr(f64(W)), None, None, ptr(Mup), ptr(Mdn), ptr(Fup), ptr(Fdn),
                              ptr(Fnet)))
        ws.close()
        return np.ascontiguousarray(Mup[:, 0])
    Ps, Ptop = float(P), float(Ptop)
    assert Ps > Ptop > 0, "surface pressure must exceed the top-of-atmosphere pressure"
    A.checkpressures(Ps, Ptop)
    fT, fμ = formprofile(None, T), formprofile(None, μ)
    nlob = 4 if nlobatto is None else int(nlobatto)
    prev, prevE, n = None, None, n0
    while True:
        Pg = f64(np.exp(np.linspace(np.log(Ptop), np.log(Ps), n + 1)))
        Pg[0], Pg[-1] = Ptop, Ps
        Mup, _, _, _ = _sweep(A, ν, Pg, g, fT, fμ, None, None, 0.841, nstream, nlob, True)
        olr = np.ascontiguousarray(Mup[:, 0])
        E = None if prev is None else _extrapolate(olr, prev)
        if prevE is not None and _converged(E, prevE, tol):
            return E
        if not _fits(nν, 2 * n + 1, nlob):
            warnings.warn(f"outgoing: refinement stopped at {n} layers before reaching tol = {tol}")
            return olr if E is None else E
        prev, prevE, n = olr, E, 2 * n


def monochromaticfluxes_radau(Mup, Mdn, τ, core, P, g, T, μ, fS, fa, *absorbers, θs=0.841, nlobatto=4, k0=2):
    """monochromaticfluxes!(M⁺, M⁻, τ, core::Radau, P, g, T, μ, 𝒻S, 𝒻a, absorbers...; θₛ) -- fluxes.jl:197-236.
    Every layer of P is split into k sub-layers equally spaced in ln P (k doubled until M⁺ and M⁻ at the levels of P
    change by less than core.tol); τ is filled with NaN like the reference does (:221).  Returns (F⁺, F⁻, Fnet)."""
    assert isinstance(core, Radau)
    A, ν, nν = unifyabsorbers(absorbers)
    P = f64(np.asarray(P, dtype=np.float64))
    assert np.all(np.diff(P) >= 0), "pressure coordinates must be in ascending order (sorted)"
    fT, fμ = formprofile(P, T), formprofile(P, μ)
    A.checkpressures(P[-1], P[0])
    checkstreams(core.nstream)
    checkazimuth(θs)
    fSν, faν = _eval_spectral(fS, ν), _eval_spectral(fa, ν)
    if τ is not None:
        τ[...] = np.nan
    lnP = np.log(P)
    prev, prevF, prevE, k = None, None, None, k0
    while True:
        frac = np.arange(k) / k
        fine = np.concatenate([np.exp(lnP[:-1, None] + (lnP[1:] - lnP[:-1])[:, None] * frac[None, :]).ravel(), P[-1:]])
        fine[::k] = P
        fine = f64(fine)
        Mu, Md, Fup, Fdn = _sweep(A, ν, fine, g, fT, fμ, fSν, faν, θs, core.nstream, nlobatto, True)
        npl = len(P)
        cur = np.concatenate([Mu[:, ::k], Md[:, ::k]], axis=1)
        curF = np.concatenate([Fup[::k], Fdn[::k]])
        E = None if prev is None else _extrapolate(cur, prev)
        EF = None if prev is None else _extrapolate(curF, prevF)
        done = prevE is not None and _converged(E, prevE, core.tol)
        if not done and not _fits(nν, 2 * k * (npl - 1) + 1, nlobatto):
            warnings.warn(f"monochromaticfluxes!(Radau): refinement stopped at {k} sub-layers before reaching tol = {core.tol}")
            done = True
        if done:
            if E is None:
                E, EF = cur, curF
            if Mup is not None:
                Mup[...] = E[:, :npl]
            if Mdn is not None:
                Mdn[...] = E[:, npl:]
            # the spectral integrals at the levels of P (∫F!, shared.jl:125-137) are linear in M: same extrapolation
            return EF[:npl].copy(), EF[npl:].copy(), EF[:npl] - EF[npl:]
        prev, prevF, prevE, k = cur, curF, E, 2 * k
