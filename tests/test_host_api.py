"""CPU-side tests of the host logic and of the C-ABI boundary (no GPU): the shared library loads, exports every
symbol include/clearsky_b200.h declares, and fails loudly -- never falls back -- when no device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import DATA, ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "clearsky_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(cs):
    from clearsky_b200 import _lib
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/clearsky_b200.h but not exported"
    # and the Python binding covers them all (cs_last_error / cs_version are bound separately)
    bound = set(_lib.SIGNATURES) | {"cs_last_error", "cs_version"}
    assert set(names) <= bound, sorted(set(names) - bound)
    assert L.cs_version() >= 100


def test_no_cpu_fallback(cs):
    """without a CUDA device the product path must fail loudly"""
    if cs.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cs.ClearSkyError):
        cs.Context(0)
    sl = cs.SpectralLines.from_file(os.path.join(DATA, "CO2.par.gz"), νmin=600, νmax=700)
    with pytest.raises(cs.ClearSkyError):
        cs.voigt(np.linspace(600, 700, 11), sl, 250.0, 1e4, 4.0)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under clearsky.jl_b200/ may reference it"""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "clearsky.jl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(d, f), encoding="utf-8").read()
                if re.search(r"\boracle\b|liboracle|orc_", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_readpar_fixture_and_filters(cs):
    """readpar / SpectralLines (par.jl:91-193, 253-284) on the reference fixtures"""
    fn = os.path.join(DATA, "CO2.par.gz")
    sl = cs.SpectralLines.from_file(fn)
    assert (sl.N, sl.M, sl.name, sl.formula) == (5599, 2, "Carbon Dioxide", "CO2")
    assert np.all(np.diff(sl.ν) >= 0) and sl.I.min() == 1 and sl.I.max() == 12
    assert abs(sl.ν[0] - 0.757206) < 1e-12 and abs(sl.S[0] - 1.751e-34) < 1e-46
    par = cs.readpar(fn, νmin=600, νmax=800, Scut=1e-25)
    assert par["ν"].min() >= 600 and par["ν"].max() <= 800 and par["S"].min() >= 1e-25
    par = cs.readpar(fn, I=[1, "2"])
    assert set(par["I"]) == {"1", "2"}
    par = cs.readpar(fn, maxlines=100)
    assert len(par["ν"]) == 100 and np.all(np.diff(par["ν"]) >= 0)
    full = cs.readpar(fn)
    assert par["S"].min() >= np.sort(full["S"])[-100]
    with pytest.raises(AssertionError):
        cs.readpar(fn, νmin=1e6)
    h2o = cs.SpectralLines.from_file(os.path.join(DATA, "H2O.par.gz"))
    assert h2o.N == 3058 and abs(h2o.na.min() + 0.23) < 1e-12
    niso, ncheb, cheb, has = sl.cheb_table()
    assert niso == 12 and ncheb[0] == 8 and has.all()


def test_molparam_table(cs):
    """the reference's only active test (test/test_molparam.jl:1-18), on the extracted table"""
    assert len(cs.MOLPARAM) == 55 and cs.TMIN == 25.0 and cs.TMAX == 1000.0
    for mp in cs.MOLPARAM:
        if mp.M < 0 or len(mp.I) <= 1:
            continue
        assert all(e <= 0.01 for e in mp.maxrelerr)
        assert all(n == len(c) for n, c in zip(mp.ncheb, mp.cheb))
        assert not any(np.isnan(x) for c in mp.cheb for x in c)
        assert all((len(c) == 0) == (not h) for c, h in zip(mp.cheb, mp.hascheb))
        assert sum(mp.A) <= 1.001


def test_domain_and_grids(cs):
    """AtmosphericDomain asserts (gases.jl:45-61), chebygrid, pressuregrid (util.jl:19-23)"""
    Ω = cs.AtmosphericDomain((150, 320), 12, (5, 1.1e5), 24)
    assert Ω.T[0] == 150 and abs(Ω.T[-1] - 320) < 1e-12 and len(Ω.P) == 24
    assert abs(Ω.P[0] - 5) < 1e-12 and abs(Ω.P[-1] / 1.1e5 - 1) < 1e-14
    x = cs.chebygrid(7)
    assert np.allclose(x, -x[::-1]) and x[0] == -1 and x[-1] == 1
    for bad in (((10, 300), 4, (1, 10), 4), ((100, 1200), 4, (1, 10), 4), ((300, 100), 4, (1, 10), 4),
                ((100, 300), 4, (10, 1), 4)):
        with pytest.raises(AssertionError):
            cs.AtmosphericDomain(*bad)
    P = cs.pressuregrid(10, 1e5, 21)
    assert abs(P[0] - 10) < 1e-12 and np.all(np.diff(P) > 0)


def test_lobatto_evaluations_and_nodes(cs):
    """lobattoevaluations (discretized.jl:11-30) and the unique-node map the C ABI expects"""
    from clearsky_b200.fluxes import _unique_nodes, formprofile, lobattoevaluations
    P = cs.pressuregrid(10, 1e5, 6)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    for nlob in (2, 3, 5):
        T, μ, Pn = lobattoevaluations(P, Γ, formprofile(P, 0.029), nlob)
        assert T.shape == (5, nlob) and np.all(μ == 0.029)
        assert np.allclose(Pn[:, 0], P[:-1]) and np.allclose(Pn[:, -1], P[1:])
        Tn, Pq = _unique_nodes(P, T, Pn, nlob)
        assert len(Tn) == 5 * (nlob - 1) + 1 and np.all(np.diff(Pq) > 0)
        assert Pq[0] == P[0] and Pq[-1] == P[-1]
    prof = cs.AtmosphericProfile(P, Γ(P))
    assert abs(prof(P[2]) - Γ(P[2])) < 1e-12
    # linear in ln P with end-cell extrapolation (atmospherics.jl:16-26)
    lo = prof(P[0] / 2)
    slope = (Γ(P[1]) - Γ(P[0])) / (np.log(P[1]) - np.log(P[0]))
    assert abs(lo - (Γ(P[0]) + slope * np.log(0.5))) < 1e-10


def test_absorber_grouping(cs):
    """UnifiedAbsorber checks (absorbers.jl:50-77, 226-229): gases need identical ν, at least one gas"""
    ν = np.linspace(1, 100, 10)
    g1, g2 = cs.GrayGas(1e-26, ν), cs.SemiGrayGas(1e-27, ν, 50.0)
    U = cs.UnifiedAbsorber(g1, g2, lambda ν, T, P: 0 * ν)
    assert U.nν == 10 and len(U.gas) == 2 and len(U.fun) == 1
    with pytest.raises(AssertionError):
        cs.UnifiedAbsorber(g1, cs.GrayGas(1e-26, ν + 1))
    with pytest.raises(ValueError):
        cs.UnifiedAbsorber(lambda ν, T, P: 0)
    with pytest.raises(AssertionError):
        cs.UnifiedAbsorber(g1, g1)
    assert g2(3, 250, 1e4) == 1e-27 and g2(9, 250, 1e4) == 0.0   # gases.jl:386


def test_bench_sharding_logic():
    """ν slicing balanced by evaluations + global trapezoid weights reproduce trapz exactly once per interval"""
    import bench
    rng = np.random.default_rng(0)
    ν = np.sort(rng.uniform(1, 100, 5000))
    νl = np.sort(rng.uniform(0, 120, 20000))
    cnt = bench.per_point_counts(ν, νl, 25.0)
    for n in (1, 2, 4, 8):
        e = bench.balanced_slices(cnt, n)
        assert e[0] == 0 and e[-1] == len(ν) and all(b >= a for a, b in zip(e, e[1:]))
        loads = [cnt[a:b].sum() for a, b in zip(e, e[1:])]
        assert max(loads) <= 1.05 * cnt.sum() / n + cnt.max()
    y = rng.normal(size=len(ν))
    w = bench.trapz_weights(ν)
    e = bench.balanced_slices(cnt, 4)
    parts = sum(np.dot(w[a:b], y[a:b]) for a, b in zip(e, e[1:]))
    assert abs(parts - np.trapezoid(y, ν)) < 1e-10


def test_radau_refinement_logic(cs):
    """the Radau-equivalent refinement (radau.py): Richardson estimate and extrapolation of a second-order sequence,
    the per-call level cap, and the core tag's defaults (shared.jl:40-47)"""
    from clearsky_b200 import radau
    assert cs.Radau() == cs.Radau(nstream=5, tol=1e-5)
    exact = np.array([1.0, 2.0, 0.5])
    seq = lambda n: exact * (1 + 0.3 / n ** 2)                      # error ~ C/n^2, like the linear-in-tau source scheme
    assert not radau._converged(seq(4), seq(2), 1e-5)
    assert radau._converged(seq(512), seq(256), 1e-5)
    assert np.max(np.abs(radau._extrapolate(seq(8), seq(4)) - exact)) < 1e-15
    # successive Richardson extrapolates of a sequence with a fourth-order remainder agree long before the raw values do
    seq4 = lambda n: exact * (1 + 0.3 / n ** 2 + 2.0 / n ** 4)
    E = lambda n: radau._extrapolate(seq4(2 * n), seq4(n))
    assert radau._converged(E(32), E(16), 1e-5) and not radau._converged(seq4(64), seq4(32), 1e-5)
    # values far below the spectral maximum are judged against 1e-3 of it, not against themselves
    a, b = np.array([1.0, 1e-9]), np.array([1.0, 2e-9])
    assert radau._converged(a, b, 1e-5)
    assert radau._fits(1000, radau.MAX_LEVELS, 4) and not radau._fits(1000, radau.MAX_LEVELS + 1, 4)
    assert not radau._fits(10 ** 7, 1025, 4)                        # the sigma workspace of one refinement level is capped


def test_dispatch_of_scalar_and_vector_pressure_forms(cs):
    """opticaldepth(P::Vector, ...) and opticaldepth(P1, P2, ...) (fluxes.jl:68 and :39) share a name in the reference;
    the Python twin dispatches on the first argument, and both fail loudly without a device (no CPU fallback)"""
    gray = cs.GrayGas(1e-26, np.linspace(1.0, 100.0, 10))
    try:
        have_gpu = cs.device_count() > 0
    except (cs.ClearSkyError, ImportError, OSError):
        have_gpu = False
    if have_gpu:       # this CPU-side test also runs on GPU boxes: there both forms work and agree with each other
        a = cs.opticaldepth(np.array([10.0, 1e5]), 9.8, 250.0, 0.029, 0.0, gray)
        b = cs.opticaldepth(1e5, 10.0, 9.8, 250.0, 0.029, 0.0, gray)
        assert np.all(np.isfinite(a)) and np.all(np.isfinite(b))
    else:
        with pytest.raises((cs.ClearSkyError, ImportError)):
            cs.opticaldepth(np.array([10.0, 1e5]), 9.8, 250.0, 0.029, 0.0, gray)
        with pytest.raises((cs.ClearSkyError, ImportError)):
            cs.opticaldepth(1e5, 10.0, 9.8, 250.0, 0.029, 0.0, gray)
    with pytest.raises(AssertionError):
        cs.opticaldepth(1e5, 10.0, 9.8, 250.0, 0.029, 2.0, gray)       # checkazimuth before any device work


def test_checknu_and_record_geometry(cs):
    """one-pass checkν keeps the reference's two asserts (gases.jl:90-95); .par record geometry of the device reader"""
    from clearsky_b200.gases import checkν
    from clearsky_b200.par import _record_geometry
    checkν(np.array([0.0, 1.0, 2.5]))
    checkν(np.array([3.0]))
    for bad, msg in ((np.array([1.0, 1.0]), "ascending"), (np.array([2.0, 1.0]), "ascending"), (np.array([-1.0, 1.0]), "positive")):
        with pytest.raises(AssertionError, match=msg):
            checkν(bad)
    rec = b" 21" + b"x" * 157
    assert _record_geometry((rec + b"\n") * 3) == (161, 3)
    assert _record_geometry((rec + b"\n") * 3 + b"\n") == (161, 3)            # trailing blank line is not a record
    assert _record_geometry((rec + b"\r\n") * 2) == (162, 2)
    assert _record_geometry((rec + b"\n") * 2 + rec) == (161, 3)               # last record without terminator
