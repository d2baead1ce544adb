"""Randomised parity sweep (fixed seeds): small problems with random line densities, grids (uniform at a random
resolution, irregular, clustered with gaps), cut-offs, pressures and temperatures, every shape, in both far-field modes,
each against the CPU oracle at the 1e-9 bar.  The cases are built to land on the class borders the kernels special-case:
cut-off windows narrower than a tile, tiles wider than the cut-off, near-centre ranges that reach the window edges,
empty windows, single-line windows, lines exactly on grid points."""
import numpy as np
import pytest

from conftest import relerr
from helpers import synthetic_lines

pytestmark = pytest.mark.gpu

SHAPES = [("doppler", 0), ("lorentz", 1), ("voigt", 2), ("PHCO2", 3)]


def _grid(rng, kind, n):
    lo = rng.uniform(5.0, 900.0)
    if kind == 0:                       # uniform, resolution from 1e-4 to 3 cm^-1
        return lo + 10 ** rng.uniform(-4, 0.5) * np.arange(n)
    if kind == 1:                       # irregular
        return np.unique(lo + rng.uniform(0, rng.uniform(1.0, 400.0), n))
    parts, x = [], lo                   # clusters separated by gaps
    while sum(len(p) for p in parts) < n:
        m = int(rng.integers(1, 200))
        parts.append(x + 10 ** rng.uniform(-3, -0.5) * np.arange(m))
        x = parts[-1][-1] + rng.uniform(0.5, 120.0)
    return np.unique(np.concatenate(parts))[:n]


@pytest.mark.parametrize("seed", range(48))
def test_random_cases_both_modes(cs, orc, seed):
    rng = np.random.default_rng(1000 + seed)
    ctx = cs.default_context()
    prev = ctx.get_farfield()
    try:
        for case in range(3):
            nl = int(rng.choice([1, 7, 300, 3000]))
            sl = synthetic_lines(cs, nl, seed=seed * 10 + case, νmax=float(rng.choice([200.0, 1500.0])))
            ν = _grid(rng, int(rng.integers(0, 3)), int(rng.choice([1, 33, 128, 129, 1000, 2500])))
            if case == 2 and nl > 1:    # put some grid points exactly on line centres
                ν = np.unique(np.concatenate([ν, sl.ν[:: max(1, nl // 17)]]))
            nlev = int(rng.integers(1, 4))
            T = rng.uniform(100.0, 400.0, nlev)
            P = 10 ** rng.uniform(0.0, 6.5, nlev)
            Pp = P * rng.uniform(0.0, 1.0, nlev)
            for name, sid in SHAPES:
                cut = float(rng.choice([0.03, 1.0, 25.0, 140.0, 500.0])) if name == "PHCO2" else float(rng.choice([0.03, 1.0, 25.0, 300.0]))
                ref = orc.xsec(sid, sl, ν, T, P, Pp, cut, nthreads=0)
                for mode in ("direct", "expansion"):
                    ctx.set_farfield(mode)
                    got = cs.xsec(name, ν, sl, T, P, Pp, cut)
                    assert relerr(got, ref, 1e-290) < 1e-9, (seed, case, name, mode, cut, nl, len(ν))
    finally:
        ctx.set_farfield(prev)


@pytest.mark.parametrize("seed", range(40))
def test_random_flux_solves(cs, orc, seed):
    """K6/K7 and the depth kernel on random atmospheres: Σ supplied from the host (log-uniform over 12 decades so that
    layers range from transparent -- the 1e-6 floor -- to opaque), random level counts / spacing, streams, Lobatto orders,
    stellar beam, albedo, zenith angle; F±, M±, τ and the total optical depth against the oracle at the 1e-8 bar"""
    from clearsky_b200.fluxes import _unique_nodes, formprofile, lobattoevaluations
    rng = np.random.default_rng(5000 + seed)
    nν = int(rng.choice([1, 31, 128, 129, 777]))
    ν = np.sort(rng.uniform(1.0, 3000.0, nν)) if nν > 1 else np.array([667.0])
    npl = int(rng.integers(2, 60))
    P = np.sort(10 ** rng.uniform(0.5, 5.0, npl))
    P = np.unique(P)
    if len(P) < 2:
        P = np.array([10.0, 1e5])
    nstream, nlob = int(rng.integers(1, 12)), int(rng.integers(2, 7))
    Tsurf, Ttop = rng.uniform(220, 330), rng.uniform(120, 220)
    fT = cs.AtmosphericProfile(np.array([P[0], P[-1]]), np.array([Ttop, Tsurf]))
    μ = float(rng.uniform(0.002, 0.044))
    g = float(rng.uniform(1.0, 25.0))
    Tl, μl, Pn = lobattoevaluations(P, fT, formprofile(P, μ), nlob)
    Tn, Pq = _unique_nodes(P, Tl, Pn, nlob)
    σ = 10 ** rng.uniform(-32, -20, (len(Tn), nν)) * (1.0 + Pq[:, None] / 1e4)
    has_sun, has_alb = rng.random() < 0.6, rng.random() < 0.6
    fS = (lambda x: 0.5 + 1e-4 * x) if has_sun else None
    fa = (lambda x: 0.05 + 1e-4 * x) if has_alb else None
    θs = float(rng.uniform(0.0, 1.5))
    ws = cs.SigmaWorkspace(ν, len(Tn))
    ws.add_host(σ)
    from clearsky_b200._lib import check, f64, lib, ptr
    m, W = cs.streamnodes(nstream)
    x, w = cs.lobattonodes(nlob)
    Tlev = f64(fT(P))
    F = cs.FluxPack(len(P), nν)
    Fup, Fdn, Fnet = np.empty(len(P)), np.empty(len(P)), np.empty(len(P))
    check(lib().cs_fluxes(ws.h, len(P), ptr(f64(P)), nlob, ptr(f64(w)), ptr(f64(μl)), ptr(Tlev), g,
                          ptr(f64(fS(ν))) if fS else None, ptr(f64(fa(ν))) if fa else None, θs, nstream, ptr(f64(m)),
                          ptr(f64(W)), None, ptr(F.τ), ptr(F.Mup), ptr(F.Mdn), ptr(Fup), ptr(Fdn), ptr(Fnet)))
    ref = orc.fluxes(ν, P, nlob, w, np.full((len(P) - 1, nlob), μ), Tlev, σ, g, fS(ν) if fS else None,
                     fa(ν) if fa else None, θs, nstream, m, W, nthreads=0)
    tol = 1e-8
    assert relerr(F.τ, ref["τ"]) < tol, (seed, "tau")
    assert relerr(F.Mup, ref["Mup"], 1e-300) < tol and relerr(F.Mdn, ref["Mdn"], 1e-300) < tol, (seed, "M")
    if nν > 1:
        assert relerr(Fup, ref["Fup"], 1e-300) < tol and relerr(Fdn, ref["Fdn"], 1e-300) < tol, (seed, "F")
        assert relerr(Fnet, ref["Fup"] - ref["Fdn"], 1e-12 * np.max(np.abs(ref["Fup"]))) < 1e-6
    τtot = np.empty(nν)
    check(lib().cs_opticaldepth(ws.h, len(P), ptr(f64(P)), nlob, ptr(f64(w)), ptr(f64(μl)), g, min(θs, 1.4), ptr(τtot)))
    assert relerr(τtot, orc.opticaldepth(P, nlob, w, μl, σ, g, min(θs, 1.4)), 1e-300) < tol
