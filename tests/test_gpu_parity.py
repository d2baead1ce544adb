"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI exactly like the
Julia wrapper would, against the CPU oracle on the same inputs and against the committed golden fixtures.

Tolerances are BASELINE.json's: relative error <= 1e-9 on cross-sections, <= 1e-8 on OLR and fluxes."""
import os

import numpy as np
import pytest

from conftest import DATA, GOLDEN, relerr
from helpers import c1_problem, synthetic_lines

pytestmark = pytest.mark.gpu

XSEC_TOL = 1e-9
FLUX_TOL = 1e-8
SHAPES = [("doppler", 0, 25.0), ("lorentz", 1, 25.0), ("voigt", 2, 25.0), ("PHCO2", 3, 500.0)]


@pytest.mark.parametrize("name,sid,cut", SHAPES)
def test_xsec_fixture_c1(cs, orc, co2, name, sid, cut):
    """BASELINE configs[0] grid (ν = 1 + 2.5 i) on the reference's own CO2 fixture, all four shapes"""
    ν, P, Γ = c1_problem(cs)
    T = Γ(P)
    Pp = 400e-6 * P
    got = cs.xsec(name, ν, co2, T, P, Pp, cut)
    ref = orc.xsec(sid, co2, ν, T, P, Pp, cut, nthreads=0)
    assert relerr(got, ref, 1e-290) < XSEC_TOL


def test_xsec_golden(cs, co2):
    G = np.load(os.path.join(GOLDEN, "c1_co2.npz"))
    ν, P, T = G["nu"], G["P"], G["T"]
    for name, key, sl_, Pp, cut in (("voigt", "sigma_voigt", slice(None, None, 5), 400e-6, 25.0),
                                    ("lorentz", "sigma_lorentz", slice(None, None, 5), 400e-6, 25.0),
                                    ("doppler", "sigma_doppler", slice(None, None, 5), 400e-6, 25.0),
                                    ("PHCO2", "sigma_phco2", slice(None, None, 10), 1.0, 500.0)):
        got = cs.xsec(name, ν, co2, T[sl_], P[sl_], Pp * P[sl_], cut)
        assert relerr(got, G[key], 1e-290) < XSEC_TOL, name


@pytest.mark.parametrize("name,sid,cut", SHAPES)
def test_xsec_fine_grid_all_regions(cs, orc, name, sid, cut):
    """0.01 cm^-1 grid over dense synthetic lines from 10 Pa to 1 bar: exercises every Faddeyeva region,
    the paired fast path, the edge predicate and ragged tiles (nν not a multiple of the tile)"""
    sl = synthetic_lines(cs, 3000, seed=7, νmax=120.0)
    ν = 30.0 + 0.01 * np.arange(4999)
    P = np.array([10.0, 300.0, 5e3, 1e5])
    T = np.array([150.0, 210.0, 250.0, 296.0])
    Pp = np.array([1e-3, 0.1, 5.0, 1e5 * 0.5])
    c = min(cut, 40.0)
    got = cs.xsec(name, ν, sl, T, P, Pp, c)
    ref = orc.xsec(sid, sl, ν, T, P, Pp, c, nthreads=0)
    assert relerr(got, ref, 1e-290) < XSEC_TOL


@pytest.fixture
def expansion(cs):
    """switch the default context to the far-field expansion for one test"""
    ctx = cs.default_context()
    prev = ctx.get_farfield()        # "direct" unless the suite runs with CS_FARFIELD=expansion
    ctx.set_farfield("expansion")
    assert ctx.get_farfield() == "expansion"
    yield ctx
    ctx.set_farfield(prev)


@pytest.mark.parametrize("name,sid", [("lorentz", 1), ("voigt", 2)])
def test_xsec_farfield_expansion(cs, orc, expansion, name, sid):
    """CS_FARFIELD_EXPANSION: well-separated far-wing lines go through a 20-term local expansion; same parity bar
    against the oracle (1e-9) and within 1e-10 of the direct mode, on uniform, ragged, gappy and irregular grids"""
    sl = synthetic_lines(cs, 6000, seed=17, νmax=200.0)
    rng = np.random.default_rng(5)
    grids = [30.0 + 0.01 * np.arange(9000),                                    # 0.01 cm^-1, several tiles, ragged end
             np.sort(rng.uniform(20.0, 180.0, 7001)),                           # irregular spacing
             np.concatenate([np.linspace(40, 45, 777), np.linspace(120, 150, 2031)]),   # a gap inside a tile
             30.0 + 2.5 * np.arange(60)]                                        # coarse: tiles wider than the cut-off
    P = np.array([10.0, 5e3, 1e5, 5e5])
    T = np.array([150.0, 230.0, 296.0, 320.0])
    Pp = np.array([1e-3, 5.0, 5e4, 4e5])
    for ν in grids:
        got = cs.xsec(name, ν, sl, T, P, Pp, 25.0)
        ref = orc.xsec(sid, sl, ν, T, P, Pp, 25.0, nthreads=0)
        assert relerr(got, ref, 1e-290) < XSEC_TOL
        expansion.set_farfield("direct")
        direct = cs.xsec(name, ν, sl, T, P, Pp, 25.0)
        expansion.set_farfield("expansion")
        assert relerr(got, direct, 1e-290) < 1e-10
    # tiny and huge cut-offs
    for cut in (0.05, 1.0, 1e4):
        ν = 60.0 + 0.01 * np.arange(3000)
        got = cs.xsec(name, ν, sl, T[:2], P[:2], Pp[:2], cut)
        ref = orc.xsec(sid, sl, ν, T[:2], P[:2], Pp[:2], cut, nthreads=0)
        assert relerr(got, ref, 1e-290) < XSEC_TOL


def test_fluxes_farfield_expansion(cs, orc, co2, expansion):
    """C1 problem on a fine grid with the expansion on: fluxes within 1e-8 of the oracle and 1e-10 of the direct mode"""
    ν = 600.0 + 0.01 * np.arange(12000)
    P = cs.pressuregrid(10.0, 1e5, 21)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    gas = cs.LineGas(co2, 400e-6, ν, "voigt", 25.0)
    F = cs.radiate(P, 9.8, Γ, 0.029, None, None, gas)
    expansion.set_farfield("direct")
    Fd = cs.radiate(P, 9.8, Γ, 0.029, None, None, gas)
    assert relerr(F.Fup, Fd.Fup) < 1e-10 and relerr(F.Fdn[1:], Fd.Fdn[1:]) < 1e-10
    assert not np.array_equal(F.Mup, Fd.Mup)        # the expansion really was used


def test_farfield_expansion_corner_cases(cs, orc, co2, expansion):
    """expansion mode on degenerate grids (one / two points: zero-width tile), through cs_bake, and on the shape it does
    not apply to (Doppler must be bit-identical to the direct mode)"""
    sl = synthetic_lines(cs, 4000, seed=23, νmax=150.0)
    T, P, Pp = np.array([200.0, 290.0]), np.array([2e3, 9e4]), np.array([1.0, 40.0])
    for ν in (np.array([75.0]), np.array([40.0, 110.0]), 70.0 + 1e-4 * np.arange(257)):
        for name, sid in (("lorentz", 1), ("voigt", 2)):
            got = cs.xsec(name, ν, sl, T, P, Pp, 25.0)
            assert relerr(got, orc.xsec(sid, sl, ν, T, P, Pp, 25.0, nthreads=0), 1e-290) < XSEC_TOL
    ν = 60.0 + 0.01 * np.arange(2000)
    x = cs.xsec("doppler", ν, sl, T, P, Pp, 25.0)
    expansion.set_farfield("direct")
    d = cs.xsec("doppler", ν, sl, T, P, Pp, 25.0)
    expansion.set_farfield("expansion")
    assert np.array_equal(x, d)
    # bake with the expansion on: table values against the oracle's direct bake + fit
    ν = 620.0 + 0.01 * np.arange(3000)
    Ω = cs.AtmosphericDomain((150, 300), 6, (10, 1e5), 8)
    gas = cs.Gas(co2, 400e-6, ν, Ω, "voigt", 25.0, keep_block=True)
    block, nz = orc.bake(orc.VOIGT, co2, ν, Ω.T, Ω.P, np.full((Ω.nP, Ω.nT), 400e-6), 25.0, nthreads=0)
    assert relerr(gas.σblock(), block, 1e-290) < XSEC_TOL
    Tq, Pq = np.array([180.0, 260.0]), np.array([300.0, 5e4])
    assert relerr(gas.rawσ(Tq, Pq), orc.gas_nodes(orc.table_fit(block), Ω.T, Ω.P, Tq, Pq, np.ones(2)), 1e-290) < XSEC_TOL


def test_xsec_line_centres(cs, orc, co2):
    """evaluation points sitting exactly on / next to line centres at low pressure (Hui / Humlicek regions)"""
    νl = co2.ν[(co2.ν > 600) & (co2.ν < 760)][::7][:150]
    off = np.array([0.0, 1e-7, 1e-5, 3e-4, 2e-3, 9e-3])
    ν = np.unique(np.sort((νl[:, None] + off[None, :]).ravel()))
    T, P = np.array([200.0, 250.0]), np.array([1.0, 50.0])
    got = cs.xsec("voigt", ν, co2, T, P, 400e-6 * P, 25.0)
    ref = orc.xsec(orc.VOIGT, co2, ν, T, P, 400e-6 * P, 25.0)
    assert relerr(got, ref, 1e-290) < XSEC_TOL


def test_voigt_direct_sum_variants(cs, orc, monkeypatch):
    """direct-mode Voigt is two launches by default (cold classes with the per-point band correction, then far_fold_kernel);
    CS_LINESUM_NO_SPLIT keeps the band in one launch, CS_LINESUM_NO_BAND is the per-tile near range.  All three against the
    oracle on a grid that crosses line centres at 10 Pa .. 1 bar; a level at 1e-5 Pa (damping parameter below the bound the
    band correction needs) must silently take the per-tile path and still agree"""
    sl = synthetic_lines(cs, 4000, seed=11, νmax=160.0)
    ν = 40.0 + 0.01 * np.arange(7013)
    ν[100:140] = np.sort(sl.ν[(sl.ν > 41.0) & (sl.ν < 110.0)][:40])          # points exactly on line centres
    ν = np.unique(ν)
    P = np.array([10.0, 2e3, 1e5])
    T = np.array([160.0, 230.0, 296.0])
    Pp = np.array([1e-2, 1.0, 3e4])
    ref = orc.xsec(orc.VOIGT, sl, ν, T, P, Pp, 25.0, nthreads=0)
    lref = orc.xsec(orc.LORENTZ, sl, ν, T, P, Pp, 25.0, nthreads=0)
    Plow = np.array([1e-5, 10.0])
    reflow = orc.xsec(orc.VOIGT, sl, ν[:1500], T[:2], Plow, 1e-3 * Plow, 25.0, nthreads=0)
    for env in ({}, {"CS_LINESUM_NO_SPLIT": "1"}, {"CS_LINESUM_NO_BAND": "1"}):
        for k in ("CS_LINESUM_NO_SPLIT", "CS_LINESUM_NO_BAND"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = cs.Context(0)                       # the switches are read when a context is created
        try:
            cs.device_lines(sl, ctx)              # upload (one launch of its own) before counting
            n0 = ctx.launches()
            got = cs.xsec("voigt", ν, sl, T, P, Pp, 25.0, ctx=ctx)
            nlaunch = ctx.launches() - n0
            assert relerr(got, ref, 1e-290) < XSEC_TOL, env
            assert nlaunch == (5 if not env else 4), (env, nlaunch)      # Qref/Q + prep + ranges + (cold, far fold | one line-sum kernel)
            assert relerr(cs.xsec("lorentz", ν, sl, T, P, Pp, 25.0, ctx=ctx), lref, 1e-290) < XSEC_TOL, env
            assert relerr(cs.xsec("voigt", ν[:1500], sl, T[:2], Plow, 1e-3 * Plow, 25.0, ctx=ctx), reflow, 1e-290) < XSEC_TOL, env
        finally:
            sl.__dict__.pop("_dev", None)         # the uploaded copy goes before its context
            ctx.close()


def test_fold_guard_coincident_lines(cs, orc):
    """the far-wing fold keeps a running product of G values q = dnu^2 + gamma^2; 40 lines sitting on the same wavenumber with
    an evaluation point exactly there (q = gamma^2 for all of them, 1e-5 Pa .. 1 bar) must not underflow it: the host bounds
    the product from the narrowest spans of 4 / 8 / 16 consecutive lines and falls back to shorter groups / the per-tile path"""
    sl = synthetic_lines(cs, 400, seed=5, νmax=60.0)
    ν0 = 30.0
    sl.ν[100:140] = ν0
    sl.ν[:] = np.sort(sl.ν)
    ν = np.unique(np.concatenate([ν0 + 0.01 * np.arange(-300, 301), [ν0]]))
    T = np.array([180.0, 296.0, 250.0])
    P = np.array([1e-5, 1e5, 50.0])
    Pp = 1e-3 * P
    for shape, sid in (("lorentz", orc.LORENTZ), ("voigt", orc.VOIGT)):
        got = cs.xsec(shape, ν, sl, T, P, Pp, 25.0)
        ref = orc.xsec(sid, sl, ν, T, P, Pp, 25.0, nthreads=0)
        assert np.all(np.isfinite(got)), shape
        assert relerr(got, ref, 1e-290) < XSEC_TOL, shape


def test_xsec_cutoff_is_inclusive(cs, orc):
    """a point exactly Δνcut away from a line is included (line_shapes.jl:10); the strict prefilter
    (line_shapes.jl:18-22) drops a line sitting exactly at min(ν) - cut"""
    sl = synthetic_lines(cs, 50, seed=3, νmax=200.0)
    sl.ν[:] = np.sort(np.round(sl.ν, 2))
    sl.ν[10] = 100.0
    sl.ν[:] = np.sort(sl.ν)
    ν = np.array([75.0, 100.0, 125.0, 125.5, 140.0, 170.0])
    for shape, sid in (("lorentz", 1), ("voigt", 2)):
        got = cs.xsec(shape, ν, sl, [260.0], [1e4], [10.0], 25.0)
        ref = orc.xsec(sid, sl, ν, [260.0], [1e4], [10.0], 25.0)
        assert relerr(got, ref, 1e-290) < XSEC_TOL
    ν2 = np.array([sl.ν[0] + 25.0, sl.ν[0] + 30.0])    # first line sits exactly at min(ν) - cut: prefiltered out
    got = cs.xsec("lorentz", ν2, sl, [260.0], [1e4], [10.0], 25.0)
    ref = orc.xsec(1, sl, ν2, [260.0], [1e4], [10.0], 25.0)
    assert relerr(got, ref, 1e-290) < XSEC_TOL


def test_xsec_edge_shapes(cs, orc, co2):
    """single point, two points, no line in range (zeros), gaps in the grid"""
    for ν in (np.array([667.5]), np.array([100.0, 5000.0]), np.array([20000.0, 20001.0]),
              np.concatenate([np.linspace(500, 510, 77), np.linspace(2300, 2310, 1031)])):
        got = cs.xsec("voigt", ν, co2, [220.0, 288.0], [1e3, 1e5], [0.4, 40.0], 25.0)
        ref = orc.xsec(orc.VOIGT, co2, ν, [220.0, 288.0], [1e3, 1e5], [0.4, 40.0], 25.0)
        assert relerr(got, ref, 1e-290) < XSEC_TOL
    assert cs.device_lines(co2).count_evals(np.array([667.5]), 25.0) == \
        orc.count_evals(np.array([667.5]), orc.included_lines(np.array([667.5]), co2.ν, 25.0), 25.0)


def test_xsec_errors(cs, co2):
    """reference pre-conditions surface as errors, not garbage (line_shapes.jl:29,59)"""
    with pytest.raises(cs.ClearSkyError):
        cs.xsec("voigt", np.array([3.0, 2.0, 1.0]), co2, [250.0], [1e4], [1.0], 25.0)
    with pytest.raises(cs.ClearSkyError):
        cs.xsec("voigt", np.array([1.0, 2.0]), co2, [20.0], [1e4], [1.0], 25.0)
    with pytest.raises(cs.ClearSkyError):
        cs.xsec("voigt", np.array([1.0, 2.0]), co2, [1200.0], [1e4], [1.0], 25.0)


def test_linearity_property(cs):
    """size-independent property at a larger size: σ is linear in S and additive over disjoint line sets"""
    sl = synthetic_lines(cs, 40000, seed=11, νmax=400.0)
    ν = 100.0 + 0.01 * np.arange(20000)
    T, P, Pp = [230.0, 280.0], [2e3, 8e4], [1.0, 30.0]
    full = cs.xsec("voigt", ν, sl, T, P, Pp, 25.0)
    import copy
    a, b = copy.copy(sl), copy.copy(sl)
    a.__dict__.pop("_dev", None); b.__dict__.pop("_dev", None)
    a.S = sl.S.copy(); b.S = sl.S.copy()
    a.S[::2] = 0.0
    b.S[1::2] = 0.0
    part = cs.xsec("voigt", ν, a, T, P, Pp, 25.0) + cs.xsec("voigt", ν, b, T, P, Pp, 25.0)
    assert relerr(part, full, 1e-290) < 1e-12
    c = copy.copy(sl); c.__dict__.pop("_dev", None); c.S = 2 * sl.S
    assert relerr(cs.xsec("voigt", ν, c, T, P, Pp, 25.0), 2 * full, 1e-290) < 1e-13


def test_bake_and_table(cs, orc, co2):
    """bake -> OpacityTable fit -> evaluation (gases.jl:75-145) vs the oracle, incl. the golden table run"""
    ν, P, Γ = c1_problem(cs)
    Ω = cs.AtmosphericDomain((140, 300), 12, (5, 1.1e5), 24)
    gas = cs.Gas(co2, 400e-6, ν, Ω, "voigt", 25.0, keep_block=True)
    Cg = np.full((Ω.nP, Ω.nT), 400e-6)
    block, nz = orc.bake(orc.VOIGT, co2, ν, Ω.T, Ω.P, Cg, 25.0, nthreads=0)
    assert gas.nzeroed == nz
    assert relerr(gas.σblock(), block, 1e-290) < XSEC_TOL
    A = orc.table_fit(block)
    T = Γ(P)
    ref = orc.gas_nodes(A, Ω.T, Ω.P, T, P, np.ones(len(P)))
    got = gas.rawσ(T, P)
    assert relerr(got, ref, 1e-290) < XSEC_TOL
    G = np.load(os.path.join(GOLDEN, "c1_co2.npz"))
    assert relerr(400e-6 * got[::5], G["tab_sigma_lev"], 1e-290) < XSEC_TOL
    # random interior (T, P)
    rng = np.random.default_rng(5)
    Tq = rng.uniform(141, 299, 9)
    Pq = np.exp(rng.uniform(np.log(6), np.log(1e5), 9))
    assert relerr(gas.rawσ(Tq, Pq), orc.gas_nodes(A, Ω.T, Ω.P, Tq, Pq, np.ones(9)), 1e-290) < XSEC_TOL
    # StrictBoundaries: outside the domain is an error
    with pytest.raises(cs.ClearSkyError):
        gas.rawσ(np.array([305.0]), np.array([1e4]))
    with pytest.raises(cs.ClearSkyError):
        gas.rawσ(np.array([250.0]), np.array([2e5]))


@pytest.mark.parametrize("nν,nq", [(777, 3), (778, 150), (64, 101), (1301, 129)])
def test_table_eval_variants(cs, orc, co2, nν, nq):
    """K4 on both kernels (TMA-fed for even nν, register-prefetch for odd nν), ragged wavenumber tiles, level counts
    around the 8/12-level groups and above the 128-level block; the Gas functor accumulates C·σ (gases.jl:278)"""
    ν = np.linspace(600.0, 760.0, nν)
    Ω = cs.AtmosphericDomain((150, 310), 9, (8, 1.05e5), 11)
    gas = cs.Gas(co2, lambda T, P: 300e-6 + 1e-9 * T, ν, Ω, "voigt", 25.0, keep_block=True)
    A = orc.table_fit(gas.σblock())
    rng = np.random.default_rng(nq)
    Tq = rng.uniform(151, 309, nq)
    Pq = np.exp(rng.uniform(np.log(9), np.log(1e5), nq))
    ref = orc.gas_nodes(A, Ω.T, Ω.P, Tq, Pq, np.ones(nq))
    assert relerr(gas.rawσ(Tq, Pq), ref, 1e-290) < XSEC_TOL
    ws = cs.SigmaWorkspace(ν, nq)
    gas.add_to(ws, Tq, Pq)
    gas.add_to(ws, Tq, Pq)
    C = 300e-6 + 1e-9 * Tq
    assert relerr(ws.read(), 2 * C[:, None] * ref, 1e-290) < XSEC_TOL


def test_zero_mixing_repair(cs, orc):
    """a wavenumber whose σ underflows to 0 at some nodes only is zeroed at all nodes (gases.jl:131-142) and its
    table becomes log(floatmin) everywhere (gases.jl:77-79)"""
    sl = synthetic_lines(cs, 5, seed=2, νmax=50.0)
    ν = np.array([10.0, 20.0, 2000.0, 2500.0])
    Ω = cs.AtmosphericDomain((100, 300), 5, (1, 1e5), 6)
    gas = cs.Gas(sl, 1e-3, ν, Ω, "doppler", 5000.0, keep_block=True)
    block, nz = orc.bake(orc.DOPPLER, sl, ν, Ω.T, Ω.P, np.full((Ω.nP, Ω.nT), 1e-3), 5000.0)
    assert gas.nzeroed == nz
    assert relerr(gas.σblock(), block, 1e-300) < XSEC_TOL
    A = orc.table_fit(block)
    ref = orc.gas_nodes(A, Ω.T, Ω.P, [200.0], [1e3], [1.0])
    assert relerr(gas.rawσ(200.0, 1e3), ref[0], 0) < 1e-300 or relerr(gas.rawσ(200.0, 1e3), ref[0], 1e-320) < XSEC_TOL


def test_cia(cs, orc, co2):
    ν = np.linspace(1.0, 3000.0, 1234)
    raw = cs.readcia(os.path.join(DATA, "CO2-CO2_2018.cia.gz"))
    T = np.array([120.0, 200.0, 233.3, 300.0, 350.0, 800.0])
    P = np.array([10.0, 1e3, 1e4, 5e4, 1e5, 2e5])
    C1 = np.full(6, 0.95)
    for extrapolate, singles in ((False, False), (True, False), (True, True), (False, True)):
        x = cs.CIATables(raw, extrapolate=extrapolate, singles=singles)
        ws = cs.SigmaWorkspace(ν, len(T))
        from clearsky_b200._lib import check, lib, ptr, f64
        check(lib().cs_sigma_add_cia(ws.h, x.handle(), ptr(f64(T)), ptr(f64(P)), ptr(f64(C1)), ptr(f64(C1))))
        ref = orc.cia_nodes(x, ν, T, P, C1, C1)
        assert relerr(ws.read(), ref, 1e-300) < XSEC_TOL, (extrapolate, singles)
    x = cs.CIATables(cs.readcia(os.path.join(DATA, "CO2-CH4_2018.cia.gz")), extrapolate=True)
    ws = cs.SigmaWorkspace(ν, len(T))
    check(lib().cs_sigma_add_cia(ws.h, x.handle(), ptr(f64(T)), ptr(f64(P)), ptr(f64(C1)), ptr(f64(1 - C1))))
    assert relerr(ws.read(), orc.cia_nodes(x, ν, T, P, C1, 1 - C1), 1e-300) < XSEC_TOL


def _oracle_fluxes(orc, cs, ν, P, fT, μ, σnodes, g, nstream, nlob, fS=None, fa=None, θs=0.841):
    m, W = cs.streamnodes(nstream)
    x, w = cs.lobattonodes(nlob)
    μn = np.full((len(P) - 1, nlob), μ)
    Tlev = np.array([float(fT(p)) for p in P])
    return orc.fluxes(ν, P, nlob, w, μn, Tlev, σnodes, g, fS, fa, θs, nstream, m, W, nthreads=0)


def test_fluxes_c1_linegas(cs, orc, co2):
    """BASELINE configs[0] with the exact line-by-line gas: OLR, F±, M±, τ vs golden (oracle) -- <= 1e-8"""
    G = np.load(os.path.join(GOLDEN, "c1_co2.npz"))
    ν, P, Γ = c1_problem(cs)
    gas = cs.LineGas(co2, 400e-6, ν, "voigt", 25.0)
    F = cs.radiate(P, 9.8, Γ, 0.029, None, None, gas, core=cs.Discretized(5, 2))
    assert relerr(F.Fup, G["lbl_Fup"]) < FLUX_TOL
    assert relerr(F.Fdn[1:], G["lbl_Fdn"][1:]) < FLUX_TOL
    assert relerr(F.τ, G["lbl_tau"]) < FLUX_TOL
    assert relerr(F.Mup[:, 0], G["lbl_Mup_toa"], 1e-300) < FLUX_TOL
    assert relerr(F.Fnet, G["lbl_Fup"] - G["lbl_Fdn"]) < FLUX_TOL
    Fup, Fdn = cs.fluxes(P, 9.8, Γ, 0.029, None, None, gas)
    assert relerr(Fup, G["lbl_Fup"]) < FLUX_TOL and abs(Fup[0] - 378.8388262928894) < 1e-5


def test_fluxes_c1_table(cs, orc, co2):
    """BASELINE configs[0] through Gas + OpacityTable, the reference's own route (fluxes.jl:311-340)"""
    G = np.load(os.path.join(GOLDEN, "c1_co2.npz"))
    ν, P, Γ = c1_problem(cs)
    Ω = cs.AtmosphericDomain((140, 300), 12, (5, 1.1e5), 24)
    gas = cs.Gas(co2, 400e-6, ν, Ω)
    Fup, Fdn = cs.fluxes(P, 9.8, Γ, 0.029, None, None, gas)
    assert relerr(Fup, G["tab_Fup"]) < FLUX_TOL
    assert relerr(Fdn[1:], G["tab_Fdn"][1:]) < FLUX_TOL
    net = cs.netfluxes(P, 9.8, Γ, 0.029, None, None, gas)
    assert relerr(net, G["tab_Fup"] - G["tab_Fdn"]) < FLUX_TOL


@pytest.mark.parametrize("nstream,nlob", [(1, 2), (3, 3), (5, 4), (8, 2), (11, 5)])
def test_fluxes_streams_lobatto_sun_albedo(cs, orc, co2, h2o, nstream, nlob):
    """two gases + CIA-free mix, stellar beam and albedo, ragged ν count, every stream/Lobatto template"""
    ν = np.linspace(50.0, 2400.0, 1177)
    P = cs.pressuregrid(20.0, 9e4, 14)
    Γ = cs.DryAdiabat(280.0, 9e4, 1040.0, 0.029, Ptropo=1.2e4)
    Ω = cs.AtmosphericDomain((150, 300), 8, (10, 1e5), 12)
    g1 = cs.Gas(co2, 400e-6, ν, Ω)
    g2 = cs.Gas(h2o, lambda T, P: min(1e-2, 1e-6 + 1e-7 * P / 1e3), ν, Ω, "lorentz", 25.0)
    gray = cs.SemiGrayGas(1e-27, ν, 1000.0)
    fun = lambda x, T, P: 1e-28 * (1 + x / 1e3) * (T / 250.0)
    fS = lambda x: 1.5 * np.exp(-((x - 1500) / 600.0) ** 2)
    fa = lambda x: 0.2 + 0.1 * np.sin(x / 300.0)
    F = cs.radiate(P, 9.8, Γ, 0.029, fS, fa, g1, g2, gray, fun, core=cs.Discretized(nstream, nlob), θs=0.6)
    # oracle: same Σ assembled on the CPU
    from clearsky_b200.fluxes import lobattoevaluations, _unique_nodes, formprofile
    Tl, μl, Pn = lobattoevaluations(P, Γ, formprofile(P, 0.029), nlob)
    Tn, Pq = _unique_nodes(P, Tl, Pn, nlob)
    σ = np.zeros((len(Tn), len(ν)))
    for gas, shp, fC in ((co2, orc.VOIGT, g1.fC), (h2o, orc.LORENTZ, g2.fC)):
        Cg = np.array([[fC(T, Pj) for T in Ω.T] for Pj in Ω.P])
        blk, _ = orc.bake(shp, gas, ν, Ω.T, Ω.P, Cg, 25.0, nthreads=0)
        σ += orc.gas_nodes(orc.table_fit(blk), Ω.T, Ω.P, Tn, Pq, np.array([fC(t, p) for t, p in zip(Tn, Pq)]))
    σ += np.where(ν <= 1000.0, 1e-27, 0.0)[None, :]
    σ += np.array([fun(ν, t, p) for t, p in zip(Tn, Pq)])
    ref = _oracle_fluxes(orc, cs, ν, P, Γ, 0.029, σ, 9.8, nstream, nlob, fS(ν), fa(ν), 0.6)
    assert relerr(F.Fup, ref["Fup"]) < FLUX_TOL
    assert relerr(F.Fdn, ref["Fdn"]) < FLUX_TOL
    assert relerr(F.Mup, ref["Mup"], 1e-300) < FLUX_TOL
    assert relerr(F.Mdn, ref["Mdn"], 1e-300) < FLUX_TOL
    assert relerr(F.τ, ref["τ"]) < FLUX_TOL


def test_gray_olr_analytic_gpu(cs):
    """GrayGas on a dry adiabat vs the analytic solution (test/test_gray.jl), rel. err < 1 %"""
    from test_oracle_pins import _gray_analytic
    g, μ, cp, Ps, Ts = 10.0, 0.01, 1e3, 1e5, 300.0
    ν = np.concatenate([np.linspace(1e-3, 10, 60)[:-1], np.linspace(10, 6000, 3000)])
    P = np.exp(np.linspace(np.log(1e-2), np.log(Ps), 601))
    Γ = cs.DryAdiabat(Ts, Ps, cp, μ)
    m, W = cs.streamnodes(5)
    for σ in (1e-27, 1e-26, 1e-25):
        Fup, _ = cs.fluxes(P, g, Γ, μ, None, None, cs.GrayGas(σ, ν))
        assert abs(Fup[0] / _gray_analytic(σ, g, μ, cp, Ps, Ts, m, W) - 1) < 0.01


def test_accelerated_absorber_and_cia_mars(cs, orc, co2):
    """config-3 flavour: pure CO2, PHCO2 shape, CO2-CO2 CIA with extrapolation, through an AcceleratedAbsorber
    built at the cell edges and used at finer radiative levels (radiative_convective.jl:72-89)"""
    ν = np.linspace(20.0, 2000.0, 700)
    Pe = cs.pressuregrid(50.0, 2e5, 9)
    Γ = cs.DryAdiabat(250.0, 2e5, 770.0, 0.044, Ptropo=1e4)
    Te = Γ(Pe)
    Ω = cs.AtmosphericDomain((100, 300), 7, (10, 2.1e5), 10)
    gas = cs.Gas(co2, 1.0, ν, Ω, "PHCO2", 500.0)
    x = cs.CIATables(cs.readcia(os.path.join(DATA, "CO2-CO2_2018.cia.gz")), extrapolate=True)
    A = cs.AcceleratedAbsorber(Te, Pe, gas, x)
    Pr = np.sort(np.concatenate([Pe, (Pe[:-1] + Pe[1:]) / 2]))
    prof = cs.AtmosphericProfile(Pe, Te)
    F = cs.radiate(Pr, 3.71, prof, 0.044, None, None, A)
    # oracle
    blk, _ = orc.bake(orc.PHCO2, co2, ν, Ω.T, Ω.P, np.ones((Ω.nP, Ω.nT)), 500.0, nthreads=0)
    Ac = orc.table_fit(blk)
    σe = orc.gas_nodes(Ac, Ω.T, Ω.P, Te, Pe, np.ones(len(Pe))) + orc.cia_nodes(x, ν, Te, Pe, np.ones(len(Pe)), np.ones(len(Pe)))
    lnσ = np.maximum(np.log(σe), np.log(np.finfo(float).tiny))
    σr = orc.accel_nodes(np.log(Pe), lnσ, Pr)
    ref = _oracle_fluxes(orc, cs, ν, Pr, prof, 0.044, σr, 3.71, 5, 2)
    assert relerr(F.Fup, ref["Fup"]) < FLUX_TOL and relerr(F.Fdn[1:], ref["Fdn"][1:]) < FLUX_TOL
    # update! with a new temperature profile
    A.update(Te + 5.0)
    σe = orc.gas_nodes(Ac, Ω.T, Ω.P, Te + 5, Pe, np.ones(len(Pe))) + orc.cia_nodes(x, ν, Te + 5, Pe, np.ones(len(Pe)), np.ones(len(Pe)))
    σr = orc.accel_nodes(np.log(Pe), np.maximum(np.log(σe), np.log(np.finfo(float).tiny)), Pr)
    ref = _oracle_fluxes(orc, cs, ν, Pr, prof, 0.044, σr, 3.71, 5, 2)
    F = cs.radiate(Pr, 3.71, prof, 0.044, None, None, A)
    assert relerr(F.Fup, ref["Fup"]) < FLUX_TOL


def test_opticaldepth(cs, orc, co2):
    """opticaldepth(P::Vector, ...; nlobatto=4) (fluxes.jl:68-97): unsorted P accepted, no τ floor"""
    ν, P, Γ = c1_problem(cs, nν=300)
    gas = cs.LineGas(co2, 400e-6, ν, "voigt", 25.0)
    τ = cs.opticaldepth(P[::-1], 9.8, Γ, 0.029, 0.3, gas, nlobatto=4)
    from clearsky_b200.fluxes import lobattoevaluations, _unique_nodes, formprofile
    Tl, μl, Pn = lobattoevaluations(P, Γ, formprofile(P, 0.029), 4)
    Tn, Pq = _unique_nodes(P, Tl, Pn, 4)
    σ = 400e-6 * orc.xsec(orc.VOIGT, co2, ν, Tn, Pq, 400e-6 * Pq, 25.0, nthreads=0)
    x, w = cs.lobattonodes(4)
    ref = orc.opticaldepth(P, 4, w, μl, σ, 9.8, 0.3)
    assert relerr(τ, ref, 1e-300) < FLUX_TOL


def test_radau_equivalents(cs, co2):
    """Radau-core entry points on the GPU (fluxes.jl:39-66, 133-158, 160-192, 197-236): convergence to tol against
    closed forms and against much finer Discretized solves (they are not 1e-8 parity targets, see radau.py)"""
    from test_oracle_pins import _gray_analytic
    from clearsky_b200 import constants as K
    # opticaldepth(P1, P2, ...) for a gray gas is closed form: tau = 1e-4 Na sigma (P1 - P2) / (g mu cos(theta))
    g, μ, cp, Ps, Ts = 10.0, 0.01, 1e3, 1e5, 300.0
    ν = np.concatenate([np.linspace(1e-3, 10, 60)[:-1], np.linspace(10, 6000, 1500)])
    Γ = cs.DryAdiabat(Ts, Ps, cp, μ)
    gray = cs.GrayGas(1e-26, ν)
    τ = cs.opticaldepth(Ps, 10.0, g, Γ, μ, 0.5, gray)
    assert relerr(τ, np.full(len(ν), 1e-4 * K.Na * 1e-26 * (Ps - 10.0) / (g * μ * np.cos(0.5))), 1e-300) < 1e-12
    assert np.allclose(cs.transmittance(Ps, 10.0, g, Γ, μ, 0.5, gray), np.exp(-τ))
    # outgoing(Ps, ...) integrates to the analytic gray OLR (test/test_gray.jl)
    m, W = cs.streamnodes(5)
    for σ in (1e-27, 1e-25):
        olr = cs.outgoing(Ps, g, Γ, μ, cs.GrayGas(σ, ν), Ptop=1e-2)
        assert abs(cs.trapz(ν, olr) / _gray_analytic(σ, g, μ, cp, Ps, Ts, m, W) - 1) < 0.01
    # real gas: table-based CO2 on the C1 atmosphere
    ν = np.linspace(500.0, 800.0, 1201)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    Ω = cs.AtmosphericDomain((140, 300), 10, (5, 1.1e5), 16)
    gas = cs.Gas(co2, 400e-6, ν, Ω, "voigt", 25.0)
    τ = cs.opticaldepth(1e5, 10.0, 9.8, Γ, 0.029, 0.0, gas, tol=1e-6)
    Pf = np.exp(np.linspace(np.log(10.0), np.log(1e5), 1025))
    assert relerr(τ, cs.opticaldepth(Pf, 9.8, Γ, 0.029, 0.0, gas, nlobatto=4), 1e-300) < 1e-5
    olr = cs.outgoing(1e5, 9.8, Γ, 0.029, gas, Ptop=10.0, tol=1e-5)
    ctx = cs.default_context()
    ctx.set_tau_floor(1e-9)
    try:    # independent reference: Richardson extrapolation of two fine Discretized solves
        Mf, _ = cs.monochromaticfluxes(Pf, 9.8, Γ, 0.029, None, None, gas, core=cs.Discretized(5, 4))
        Mc, _ = cs.monochromaticfluxes(Pf[::2], 9.8, Γ, 0.029, None, None, gas, core=cs.Discretized(5, 4))
    finally:
        ctx.set_tau_floor(1e-6)
    ref = (4 * Mf[:, 0] - Mc[:, 0]) / 3
    assert relerr(olr, ref, 1e-3 * ref.max()) < 5e-5
    # outgoing(P::Vector) == M+ at the top level of the Discretized core with nlobatto = 3, no sun, black surface
    P = cs.pressuregrid(10.0, 1e5, 31)
    Mu, _ = cs.monochromaticfluxes(P, 9.8, Γ, 0.029, None, None, gas, core=cs.Discretized(5, 3))
    assert np.array_equal(cs.outgoing(P[::-1], 9.8, Γ, 0.029, gas), Mu[:, 0])
    # monochromaticfluxes!(…, core::Radau, …): levels of P from a refined solve; tau is NaN like the reference's
    F = cs.FluxPack(len(P), len(ν))
    cs.radiate_(F, cs.Radau(5, 1e-5), P, 9.8, Γ, 0.029, lambda x: 0.1 + 0 * x, 0.2, gas)
    assert np.all(np.isnan(F.τ))
    k = 32
    lnP = np.log(P)
    fine = np.concatenate([np.exp(lnP[:-1, None] + (lnP[1:] - lnP[:-1])[:, None] * (np.arange(k) / k)[None, :]).ravel(), P[-1:]])
    fine[::k] = P
    ctx.set_tau_floor(1e-9)
    try:
        Mu, Md = cs.monochromaticfluxes(fine, 9.8, Γ, 0.029, lambda x: 0.1 + 0 * x, 0.2, gas, core=cs.Discretized(5, 4))
    finally:
        ctx.set_tau_floor(1e-6)
    assert relerr(F.Mup, Mu[:, ::k], 1e-3 * Mu.max()) < 5e-5 and relerr(F.Mdn, Md[:, ::k], 1e-3 * Md.max()) < 5e-5
    assert abs(F.Fup[0] / cs.trapz(ν, Mu[:, 0]) - 1) < 5e-5


def test_flux_errors(cs, co2):
    ν, P, Γ = c1_problem(cs, nν=64)
    gas = cs.GrayGas(1e-26, ν)
    with pytest.raises(AssertionError):
        cs.fluxes(P[::-1], 9.8, Γ, 0.029, None, None, gas)        # issorted(P) (fluxes.jl:257)
    with pytest.raises(AssertionError):
        cs.fluxes(P, 9.8, Γ, 0.029, None, None, gas, θs=2.0)       # checkazimuth (fluxes.jl:4-6)
    with pytest.raises(AssertionError):
        cs.Gas(co2, 1.5, ν, cs.AtmosphericDomain((150, 300), 4, (10, 1e5), 4))   # gases.jl:124


def test_rcm_heating_and_step(cs, orc, co2):
    """RCM heating!/step! (radiative_convective.jl:109-151) on the GPU vs the same host arithmetic driven by the
    oracle's fluxes: AcceleratedAbsorber frozen at the initial edge temperatures, radmul = 2"""
    ν = np.linspace(50.0, 2000.0, 500)
    Pe = cs.pressuregrid(50.0, 1e5, 11)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1.5e4)
    Te = Γ(Pe)
    Ω = cs.AtmosphericDomain((120, 320), 8, (20, 1.1e5), 12)
    gas = cs.Gas(co2, 400e-6, ν, Ω)
    fS = lambda x: 0.3 * np.exp(-((x - 1200.0) / 500.0) ** 2)
    rcm = cs.RCM(Pe, Te, 9.8, 0.029, fS, 0.25, 1040.0, 1e7, gas, radmul=2)
    assert len(rcm.Pr) == 21 and np.all(np.diff(rcm.Pr) > 0)
    # oracle-driven twin
    blk, _ = orc.bake(orc.VOIGT, co2, ν, Ω.T, Ω.P, np.full((Ω.nP, Ω.nT), 400e-6), 25.0, nthreads=0)
    σe = orc.gas_nodes(orc.table_fit(blk), Ω.T, Ω.P, Te, Pe, np.full(len(Pe), 400e-6))
    lnσ = np.maximum(np.log(σe), np.log(np.finfo(float).tiny))
    σr = orc.accel_nodes(np.log(Pe), lnσ, rcm.Pr)
    m, W = cs.streamnodes(5)
    x, w = cs.lobattonodes(2)
    P, T = rcm.P.copy(), rcm.T.copy()

    def heating(T):
        fT = cs.AtmosphericProfile(P, T)
        Tlev = fT(rcm.Pr)
        f = orc.fluxes(ν, rcm.Pr, 2, w, np.full((len(rcm.Pr) - 1, 2), 0.029), Tlev, σr, 9.8, fS(ν), np.full(len(ν), 0.25),
                       0.841, 5, m, W, nthreads=0, full=False)
        R = -cs.AtmosphericProfile(rcm.Pr, f["Fnet"])(Pe)
        H = np.empty(len(Pe))
        H[:-1] = (9.8 / 1040.0) * (R[:-1] - R[1:]) / (Pe[1:] - Pe[:-1])
        H[-1] = R[-1] / 1e7
        return H

    for _ in range(3):
        Href = heating(T)
        rcm.step_(3600.0)
        T = T + 3600.0 * Href
        scale = np.max(np.abs(Href))
        assert np.max(np.abs(rcm.H - Href)) < 1e-8 * scale
        assert np.max(np.abs(rcm.T - T)) < 1e-9 * np.max(T)
    # ---- device-resident loop (cs_rcm_step: Fnet -> edges, heating rates, T += dt*H and the level temperatures on the
    # device, three kernels per step replayed from a CUDA graph) against the same oracle-driven twin, continuing the run
    for nsteps in (1, 4):
        for _ in range(nsteps):
            Href = heating(T)
            T = T + 3600.0 * Href
        rcm.steps_(3600.0, nsteps)
        assert np.max(np.abs(rcm.H - Href)) < 1e-8 * np.max(np.abs(Href))
        assert np.max(np.abs(rcm.T - T)) < 1e-9 * np.max(T)
    # heating! without moving T (dt = 0), and the fluxes of the last step against a plain flux call
    Tb = rcm.T.copy()
    rcm.steps_(0.0, 1)
    assert np.array_equal(rcm.T, Tb)
    Fup, Fdn = cs.fluxes(rcm.Pr, 9.8, cs.AtmosphericProfile(rcm.P, rcm.T), 0.029, fS, 0.25, rcm.A)
    assert relerr(rcm.F.Fup, Fup) < 1e-12 and relerr(rcm.F.Fdn, Fdn, 1e-3) < 1e-12
    # host loop and device loop are interchangeable step by step
    rcm2 = cs.RCM(Pe, Te, 9.8, 0.029, fS, 0.25, 1040.0, 1e7, gas, radmul=2)
    rcm3 = cs.RCM(Pe, Te, 9.8, 0.029, fS, 0.25, 1040.0, 1e7, gas, radmul=2)
    for _ in range(6):
        rcm2.step_(1800.0)
    rcm3.steps_(1800.0, 6)
    assert relerr(rcm2.T, rcm3.T) < 1e-13 and np.max(np.abs(rcm2.H - rcm3.H)) < 1e-9 * np.max(np.abs(rcm2.H))
    rcm.close(); rcm3.close()


def test_rcm_device_loop_variants(cs, co2):
    """cs_rcm_step against the host loop for the other kernel variants: no sun / no albedo (null pointers), nlobatto = 3
    (intermediate quadrature nodes), generic stream count, radmul = 4"""
    ν = np.linspace(50.0, 2000.0, 333)
    Pe = cs.pressuregrid(50.0, 1e5, 9)
    Te = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1.5e4)(Pe)
    Ω = cs.AtmosphericDomain((120, 330), 8, (20, 1.1e5), 12)
    gas = cs.Gas(co2, 400e-6, ν, Ω)
    fS = lambda x: 0.3 * np.exp(-((x - 1200.0) / 500.0) ** 2)
    for core, sun, alb, radmul in ((cs.Discretized(5, 2), None, None, 2), (cs.Discretized(4, 3), fS, 0.1, 2),
                                   (cs.Discretized(7, 2), fS, None, 4)):
        a = cs.RCM(Pe, Te, 9.8, 0.029, sun, alb, lambda T, P: 1000.0 + 0.1 * (T - 250.0), 1e7, gas, radmul=radmul, core=core)
        b = cs.RCM(Pe, Te, 9.8, 0.029, sun, alb, 1.0, 1e7, gas, radmul=radmul, core=core)
        cp0 = np.array([1000.0 + 0.1 * (a.T[i] - 250.0) for i in range(a.np - 1)])     # held at the first call's T
        b.fcp = lambda T, P, _c=cp0, _P=b.P: float(_c[int(np.argmin(np.abs(_P[:-1] - P)))])
        a.steps_(900.0, 5)
        for _ in range(5):
            b.step_(900.0)
        assert relerr(a.T, b.T) < 1e-13 and np.max(np.abs(a.H - b.H)) < 1e-9 * np.max(np.abs(b.H))
        assert np.max(np.abs(a.F.Fnet - b.F.Fnet)) < 1e-10 * np.max(np.abs(b.F.Fnet))
        a.close()


def test_rcm_jacobian_batched(cs, co2):
    """jacobian! (radiative_convective.jl:154-171): the np+1 flux solves as one batched call (cs_fluxes_batch, shared
    layer depths and transmittances) equal the reference's loop of heating! calls; 11 profiles = two launches (8 + 3);
    with sun and albedo, nstream 5 (specialised kernel) and 3 (generic kernel)"""
    ν = np.linspace(50.0, 2000.0, 777)
    Pe = cs.pressuregrid(50.0, 1e5, 10)
    Te = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1.5e4)(Pe)
    Ω = cs.AtmosphericDomain((120, 330), 8, (20, 1.1e5), 12)
    gas = cs.Gas(co2, 400e-6, ν, Ω)
    fS = lambda x: 0.3 * np.exp(-((x - 1200.0) / 500.0) ** 2)
    for core in (cs.Discretized(5, 2), cs.Discretized(3, 3)):
        a = cs.RCM(Pe, Te, 9.8, 0.029, fS, 0.25, 1040.0, 1e7, gas, radmul=2, core=core)
        b = cs.RCM(Pe, Te, 9.8, 0.029, fS, 0.25, 1040.0, 1e7, gas, radmul=2, core=core)
        a.jacobian_(0.5)
        b.jacobian_(0.5, batched=False)
        b.heating_()          # the loop leaves H and F of the last perturbed profile behind; the batch leaves the base state
        assert np.max(np.abs(a.H - b.H)) < 1e-10 * np.max(np.abs(b.H)) and relerr(a.F.Fnet, b.F.Fnet) < 1e-11
        assert np.max(np.abs(a.J - b.J)) < 1e-9 * np.max(np.abs(b.J))
    # the batched entry point on its own: every profile equals a plain fluxes() call
    A = cs.AcceleratedAbsorber(Te, Pe, gas)
    P = cs.pressuregrid(50.0, 1e5, 19)
    Ts = [cs.AtmosphericProfile(Pe, Te + d) for d in (0.0, 3.0, -7.0)]
    Fup, Fdn = cs.fluxes_batch(P, 9.8, Ts, 0.029, fS, 0.25, A)
    for k, fT in enumerate(Ts):
        u, d = cs.fluxes(P, 9.8, fT, 0.029, fS, 0.25, A)
        assert relerr(Fup[k], u) < 1e-11 and relerr(Fdn[k], d) < 1e-11


def test_full_size_properties_c2(cs):
    """BASELINE configs[1] at FULL size (500k lines, 300k ν, 101 levels): size-independent properties.
    (1) additivity over a partition of the line list, (2) ν-sharding invariance of the spectrally integrated fluxes
    with global trapezoid weights (the multi-GPU decomposition), (3) OLR within physical bounds."""
    import bench
    from clearsky_b200._lib import check, f64, lib, ptr
    wl = bench.make_workload(cs, "c2")
    ν, P, T = wl["ν"], wl["P"], wl["T"]
    nlev = len(P)
    (co2, Cc), (h2o, Ch) = wl["gases"]
    ws = cs.SigmaWorkspace(ν, nlev)
    Tn, Pn = f64(T), f64(P)

    def add(sl, C):
        dl = cs.DeviceLines(sl)
        check(lib().cs_sigma_add_lines(ws.h, dl.h, 2, ptr(Tn), ptr(Pn), ptr(f64(np.full(nlev, C))), 25.0))

    def sub(sl, mask):
        return cs.SpectralLines(sl.name, sl.formula, int(mask.sum()), sl.M, sl.I[mask], sl.μ[mask], sl.A[mask], sl.ν[mask],
                                sl.S[mask], sl.γa[mask], sl.γs[mask], sl.Epp[mask], sl.na[mask])

    add(co2, Cc)
    full = ws.read()
    ws.zero()
    m = np.zeros(co2.N, dtype=bool)
    m[::3] = True
    add(sub(co2, m), Cc)
    add(sub(co2, ~m), Cc)
    part = ws.read()
    assert relerr(part, full, 1e-300) < 1e-12
    del part
    # full Σ and fluxes
    add(h2o, Ch)
    m_, W = cs.streamnodes(5)
    x, w = cs.lobattonodes(2)
    μn = f64(np.full((nlev - 1, 2), 0.029))
    Fu, Fd, Fn = np.empty(nlev), np.empty(nlev), np.empty(nlev)
    check(lib().cs_fluxes(ws.h, nlev, ptr(Pn), 2, ptr(f64(w)), ptr(μn), ptr(Tn), 9.8, None, None, 0.841, 5, ptr(f64(m_)),
                          ptr(f64(W)), None, None, None, None, ptr(Fu), ptr(Fd), ptr(Fn)))
    σT4 = 5.67037442e-8 * 288.0 ** 4
    assert 0 < Fu[0] < Fu[-1] < σT4 and np.all(np.diff(Fd) >= -1e-9) and Fd[0] == 0.0
    # ν-sharded evaluation: 3 contiguous slices balanced by evaluations, global weights, partial sums added
    σ = ws.read()
    cnt = bench.per_point_counts(ν, co2.ν, 25.0) + bench.per_point_counts(ν, h2o.ν, 25.0)
    e = bench.balanced_slices(cnt, 3)
    wg = bench.trapz_weights(ν)
    tot = np.zeros(2 * nlev)
    for a, b in zip(e, e[1:]):
        wsl = cs.SigmaWorkspace(ν[a:b], nlev)
        wsl.add_host(np.ascontiguousarray(σ[:, a:b]))
        fu, fd, fn = np.empty(nlev), np.empty(nlev), np.empty(nlev)
        check(lib().cs_fluxes(wsl.h, nlev, ptr(Pn), 2, ptr(f64(w)), ptr(μn), ptr(Tn), 9.8, None, None, 0.841, 5,
                              ptr(f64(m_)), ptr(f64(W)), ptr(f64(wg[a:b])), None, None, None, ptr(fu), ptr(fd), ptr(fn)))
        tot += np.concatenate([fu, fd])
    assert relerr(tot[:nlev], Fu) < 1e-12 and relerr(tot[nlev + 1:], Fd[1:]) < 1e-12


def test_table_export_import_roundtrip(cs, orc, co2):
    """cs_table_block -> cs_table_from_block (persistence path) against the ORACLE, not against
    itself: the exported block equals orc.bake (what the reference hands to OpacityTable, gases.jl:109-144, zero-mixing
    repair included), and a table re-imported from that block evaluates like orc.table_fit + orc.gas_nodes on it"""
    import ctypes as C
    from clearsky_b200._lib import check, f64, lib, ptr
    ν, P, Γ = c1_problem(cs, nν=400)
    Ω = cs.AtmosphericDomain((140, 300), 9, (5, 1.1e5), 11)
    gas = cs.Gas(co2, 400e-6, ν, Ω, keep_block=True)
    blk = gas.σblock()
    oblk, onz = orc.bake(orc.VOIGT, co2, ν, Ω.T, Ω.P, np.full((Ω.nP, Ω.nT), 400e-6), 25.0, nthreads=0)
    assert np.asarray(blk).reshape(oblk.shape).shape == oblk.shape
    assert relerr(np.asarray(blk).reshape(oblk.shape), oblk, 1e-290) < XSEC_TOL
    assert np.array_equal(np.asarray(blk).reshape(oblk.shape) == 0, oblk == 0) and gas.nzeroed == onz
    Tq, Pq = Γ(P), P
    want = orc.gas_nodes(orc.table_fit(oblk), Ω.T, Ω.P, Tq, Pq, np.ones(len(Pq)))
    # a block that did not come from this library at all: the oracle's, handed to cs_table_from_block like a Julia caller would
    h2 = C.c_void_p()
    check(lib().cs_table_from_block(gas.ctx.h, len(ν), Ω.nT, ptr(f64(Ω.T)), Ω.nP, ptr(f64(Ω.P)), ptr(f64(oblk)), C.byref(h2)))
    out2 = np.empty((len(P), len(ν)))
    check(lib().cs_table_eval(h2, len(P), ptr(f64(Tq)), ptr(f64(Pq)), ptr(out2)))
    lib().cs_table_free(h2)
    assert relerr(out2, want, 1e-290) < XSEC_TOL
    h = C.c_void_p()
    check(lib().cs_table_from_block(gas.ctx.h, len(ν), Ω.nT, ptr(f64(Ω.T)), Ω.nP, ptr(f64(Ω.P)), ptr(f64(blk)), C.byref(h)))
    T = Γ(P)
    out = np.empty((len(P), len(ν)))
    check(lib().cs_table_eval(h, len(P), ptr(f64(T)), ptr(f64(P)), ptr(out)))
    lib().cs_table_free(h)
    assert relerr(out, gas.rawσ(T, P), 1e-300) < 1e-13
    assert relerr(out, want, 1e-290) < XSEC_TOL


def test_phco2_all_chi_classes(cs, orc):
    """PHCO2 with the default 500 cm^-1 cut-off on a fine grid: every chi class (|dnu| < 3, < 30, < 120, >= 120),
    lines straddling the class borders, both sides of the tile, several temperatures (B1, B2 depend on T)"""
    sl = synthetic_lines(cs, 6000, seed=23, νmax=1300.0)
    ν = 640.0 + 0.01 * np.arange(2500)
    T = np.array([140.0, 220.0, 300.0])
    P = np.array([2e2, 2e4, 2e5])
    got = cs.xsec("PHCO2", ν, sl, T, P, P, 500.0)
    ref = orc.xsec(orc.PHCO2, sl, ν, T, P, P, 500.0, nthreads=0)
    assert relerr(got, ref, 1e-290) < XSEC_TOL
    # irregular grid with a tile wider than 3 cm^-1 (no plain class) and one wider than 27 (no class at all)
    ν2 = np.unique(np.concatenate([np.linspace(300, 300.9, 100), np.linspace(301, 330, 120), np.linspace(331, 400, 30)]))
    got = cs.xsec("PHCO2", ν2, sl, T, P, 0.5 * P, 500.0)
    ref = orc.xsec(orc.PHCO2, sl, ν2, T, P, 0.5 * P, 500.0, nthreads=0)
    assert relerr(got, ref, 1e-290) < XSEC_TOL


def test_phco2_farfield_expansion(cs, orc, expansion):
    """expansion mode for PHCO2: the F3/F4 chi classes (|dnu| >= 30 for the whole tile) go through power-law expansions
    with the chi*gamma correction series; parity against the oracle at 1e-9 and against the direct mode at 1e-10, over
    pressures from 10 Pa to 2 bar, at 50 bar (series not allowed by the host bound: direct sum), on a grid with wide
    tiles (no expansion) and with cut-offs that remove some classes"""
    sl = synthetic_lines(cs, 8000, seed=29, νmax=1500.0)
    ν = 700.0 + 0.01 * np.arange(2900)
    T = np.array([120.0, 200.0, 250.0, 310.0, 250.0])
    P = np.array([10.0, 3e3, 2e5, 2e5, 5e6])
    for cut, Pp in ((500.0, P), (500.0, 0.3 * P), (100.0, P), (25.0, P)):
        got = cs.xsec("PHCO2", ν, sl, T, P, Pp, cut)
        assert relerr(got, orc.xsec(orc.PHCO2, sl, ν, T, P, Pp, cut, nthreads=0), 1e-290) < XSEC_TOL
        expansion.set_farfield("direct")
        direct = cs.xsec("PHCO2", ν, sl, T, P, Pp, cut)
        expansion.set_farfield("expansion")
        assert relerr(got, direct, 1e-290) < 1e-10
        if cut == 500.0:
            assert not np.array_equal(got[2], direct[2])      # the expansion really ran at 2 bar
            assert np.array_equal(got[4], direct[4])          # ... and did not at 50 bar
    ν2 = np.unique(np.concatenate([np.linspace(300, 300.9, 100), np.linspace(301, 330, 120), np.linspace(331, 900, 130)]))
    got = cs.xsec("PHCO2", ν2, sl, T, P, 0.5 * P, 500.0)
    assert relerr(got, orc.xsec(orc.PHCO2, sl, ν2, T, P, 0.5 * P, 500.0, nthreads=0), 1e-290) < XSEC_TOL


def test_single_process_device_group(cs, orc, co2):
    """cs_group: ν-sharded fluxes from ONE process over all visible GPUs (NCCL all-reduce inside the library);
    with one GPU the group degenerates to a single slice.  Result must equal the unsharded run to 1e-12."""
    ν = np.linspace(500.0, 900.0, 4001)
    P = cs.pressuregrid(10.0, 1e5, 16)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    ndev = cs.device_count()
    grp = cs.DeviceGroup(list(range(ndev)))
    fS = lambda x: 0.2 + 0 * x
    Fup, Fdn, Fnet = cs.sharded_fluxes(grp, P, 9.8, Γ, 0.029, fS, 0.3, [(co2, 400e-6, "voigt", 25.0)], ν)
    gas = cs.LineGas(co2, 400e-6, ν, "voigt", 25.0)
    F = cs.radiate(P, 9.8, Γ, 0.029, fS, 0.3, gas)
    assert relerr(Fup, F.Fup) < 1e-12 and relerr(Fdn, F.Fdn) < 1e-12
    # emulate a 3-way split on whatever devices exist (contexts may share a device)
    grp3 = cs.DeviceGroup([i % ndev for i in range(3)]) if ndev >= 3 else None
    if grp3 is not None:
        Fup3, Fdn3, _ = cs.sharded_fluxes(grp3, P, 9.8, Γ, 0.029, fS, 0.3, [(co2, 400e-6, "voigt", 25.0)], ν)
        assert relerr(Fup3, F.Fup) < 1e-12
        grp3.close()
    grp.close()


def test_single_process_device_group_with_cia(cs, co2):
    """config-3 flavour through cs_group: PHCO2 line gas + CO2-CO2 CIA on every slice; equals the unsharded run"""
    ν = np.linspace(20.0, 1500.0, 3001)
    P = cs.pressuregrid(10.0, 2e5, 12)
    Γ = cs.DryAdiabat(250.0, 2e5, 770.0, 0.044, Ptropo=1e4)
    x = cs.CIATables(os.path.join(DATA, "CO2-CO2_2018.cia.gz"), extrapolate=True)
    ndev = cs.device_count()
    grp = cs.DeviceGroup(list(range(ndev)))      # NCCL needs distinct devices: one slice per visible GPU
    sh = cs.ShardedLineByLine(grp, [(co2, 1.0, "PHCO2", 500.0)], ν, cia=[(x, 0, 0)])
    Fup, Fdn, Fnet = sh.fluxes(P, 3.71, Γ, 0.044)
    gas = cs.LineGas(co2, 1.0, ν, "PHCO2", 500.0)
    F = cs.radiate(P, 3.71, Γ, 0.044, None, None, gas, x)
    assert relerr(Fup, F.Fup) < 1e-12 and relerr(Fdn[1:], F.Fdn[1:]) < 1e-12
    nocia = cs.radiate(P, 3.71, Γ, 0.044, None, None, gas)
    assert relerr(nocia.Fup[:1], F.Fup[:1]) > 1e-4          # the CIA term is not a no-op
    with pytest.raises(AssertionError):
        cs.ShardedLineByLine(grp, [(co2, 1.0, "PHCO2", 500.0)], ν,
                             cia=[(cs.CIATables(os.path.join(DATA, "CO2-CH4_2018.cia.gz")), 0, 0)])
    grp.close()


def test_sharded_tables_and_rcm(cs, co2, h2o):
    """ShardedAbsorber: opacity tables baked per ν slice on every device of the group, AcceleratedAbsorber per slice, and the
    RCM loop on top (BASELINE configs[4]); equals the single-context run"""
    ν = np.linspace(400.0, 1000.0, 2401)
    Ω = cs.AtmosphericDomain((140, 320), 8, (5, 1.1e5), 12)
    Pe = cs.pressuregrid(10.0, 1e5, 13)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    Te = Γ(Pe)
    grp = cs.DeviceGroup(list(range(cs.device_count())))
    sh = cs.ShardedAbsorber(grp, ν, lambda νs, ctx: (cs.Gas(co2, 400e-6, νs, Ω, ctx=ctx), cs.Gas(h2o, 1e-3, νs, Ω, ctx=ctx)))
    g1, g2 = cs.Gas(co2, 400e-6, ν, Ω), cs.Gas(h2o, 1e-3, ν, Ω)
    P = cs.pressuregrid(10.0, 1e5, 25)
    Fup, Fdn, Fnet = sh.fluxes(P, 9.8, Γ, 0.029, lambda x: 0.3 + 0 * x, 0.25)
    F = cs.radiate(P, 9.8, Γ, 0.029, lambda x: 0.3 + 0 * x, 0.25, g1, g2)
    assert relerr(Fup, F.Fup) < 1e-12 and relerr(Fdn, F.Fdn) < 1e-12
    with pytest.raises(AssertionError):
        sh.fluxes(cs.pressuregrid(1.0, 1e5, 9), 9.8, Γ, 0.029)       # below the table domain (checkpressures)
    a = cs.RCM(Pe, Te, 9.8, 0.029, None, None, 1040.0, 1e7, sh, radmul=2)
    b = cs.RCM(Pe, Te, 9.8, 0.029, None, None, 1040.0, 1e7, g1, g2, radmul=2)
    for _ in range(3):
        a.step_(3600.0)
        b.step_(3600.0)
    assert relerr(a.T, b.T) < 1e-12 and relerr(a.H, b.H, 1e-30) < 1e-9
    sh.update(a.Te + 1.0)                                               # update! re-snapshots every slice
    b.A.update(b.Te + 1.0)
    a.heating_(); b.heating_()
    assert relerr(a.H, b.H, 1e-30) < 1e-9
    # device-resident loop over the group (cs_group_rcm_step: partial fluxes -> ncclAllReduce -> column update per step)
    a.steps_(1800.0, 4)
    for _ in range(4):
        b.step_(1800.0)
    assert relerr(a.T, b.T) < 1e-12 and np.max(np.abs(a.H - b.H)) < 1e-9 * np.max(np.abs(b.H))
    a.close()
    del a, sh
    grp.close()


def test_rcm_fused_peer_step(cs, co2):
    """nu-sharded RCM step with the exchange fused into the step's tail kernel (cs_rcm_enqueue_step_peer: peer-memory mailboxes,
    flags, rank-ordered sum, column update) -- two "ranks" as two contexts on ONE device so that it runs on a one-GPU box (their
    mailboxes are then plain device pointers; between processes they are cs_ipc_export / cs_ipc_open mappings, bench.py c5).
    Both ranks must hold bit-identical columns, equal to the unsharded device loop."""
    import ctypes as C

    import bench
    from clearsky_b200._lib import check, f64, lib, ptr
    ν = np.linspace(500.0, 900.0, 2001)
    wts = bench.trapz_weights(ν)
    wl = dict(Pe=cs.pressuregrid(10.0, 1e5, 13), radmul=2)
    wl["Te"] = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)(wl["Pe"])
    col = bench._HostColumn(cs, wl)
    Pe, Pr, nrad, npc = f64(wl["Pe"]), f64(col.Pr), len(col.Pr), len(wl["Pe"])
    m, W = (f64(x) for x in cs.streamnodes(5))
    wlob = f64(cs.lobattonodes(2)[1])
    μn = f64(np.full((nrad - 1, 2), 0.029))
    cp = f64(np.full(npc - 1, 1040.0))
    Tlev = f64(cs.AtmosphericProfile(col.P, col.T)(col.Pr))

    def make(ctx, i0, i1):
        νs = np.ascontiguousarray(ν[i0:i1])
        ws = cs.SigmaWorkspace(νs, nrad, ctx)
        cs.UnifiedAbsorber(cs.LineGas(co2, 400e-6, νs, "voigt", 25.0, ctx=ctx)).sigma_nodes(ws, Tlev, Pr)
        h = C.c_void_p()
        check(lib().cs_rcm_create(ws.h, npc, ptr(Pe), ptr(f64(col.P)), ptr(f64(col.T)), ptr(cp), 1e7, nrad, ptr(Pr), 2, ptr(wlob),
                                  ptr(μn), 9.8, None, None, 0.841, 5, ptr(m), ptr(W), ptr(f64(wts[i0:i1])), C.byref(h)))
        return h, ws

    def state(h):
        T, H, Fu, Fd = np.empty(npc), np.empty(npc), np.empty(nrad), np.empty(nrad)
        check(lib().cs_rcm_state(h, ptr(T), ptr(H), None, ptr(Fu), ptr(Fd), None))
        return T, H, Fu, Fd

    ctxs = [cs.Context(0), cs.Context(0), cs.Context(0)]
    full, _w0, parts = None, None, []
    try:
        full, _w0 = make(ctxs[0], 0, len(ν))
        parts = [make(ctxs[1], 0, 1000), make(ctxs[2], 1000, len(ν))]
        boxes = (C.c_void_p * 2)()
        for q, (h, _) in enumerate(parts):
            b = C.c_void_p()
            check(lib().cs_rcm_peer_mailbox(h, 2, C.byref(b), None))
            boxes[q] = b.value
        for q, (h, _) in enumerate(parts):
            check(lib().cs_rcm_peer_connect(h, q, 2, boxes))
        nsteps = 5
        for _ in range(nsteps):                        # enqueue only: the two tails meet on the device
            for h, _ in parts:
                check(lib().cs_rcm_enqueue_step_peer(h, 1800.0))
        check(lib().cs_rcm_step(full, 1800.0, nsteps))
        S = [state(h) for h, _ in parts]
        for h, _ in parts:
            n, late = C.c_int64(0), C.c_int32(0)
            check(lib().cs_rcm_peer_status(h, C.byref(n), C.byref(late)))
            assert n.value == nsteps and late.value == 0
        for x, y in zip(S[0], S[1]):
            assert np.array_equal(x, y)                # rank-ordered sums: the same bits on both ranks
        T, H, Fu, Fd = state(full)
        assert relerr(S[0][0], T) < 1e-12 and np.max(np.abs(S[0][1] - H)) < 1e-9 * np.max(np.abs(H))
        assert relerr(S[0][2], Fu) < 1e-12 and relerr(S[0][3][1:], Fd[1:]) < 1e-12
        with pytest.raises(cs.ClearSkyError):
            check(lib().cs_rcm_peer_mailbox(parts[0][0], 2, C.byref(C.c_void_p()), None))      # one mailbox per column
        with pytest.raises(cs.ClearSkyError):
            check(lib().cs_rcm_enqueue_step_peer(full, 1800.0))                                 # not connected
    finally:
        for h in [full] + [h for h, _ in parts]:
            if h is not None:
                lib().cs_rcm_free(h)
        del _w0, parts
        co2.__dict__.pop("_dev", None)
        for c in ctxs:
            c.close()


def test_group_spans_distinct_gpus(cs, co2):
    """driver-visible proof of the library's own NCCL path (cs_group.cu: ncclAllReduce over the group's communicator): needs two
    DISTINCT devices, so it is skipped on a one-GPU box (`gpurun --gpus 2 -- python -m pytest tests -m gpu -k distinct_gpus`).
    nu-sharded fluxes through the group equal the one-device run, and after the call EVERY device's buffer holds the same reduced
    sums (each started from its own slice's partial integrals)."""
    ndev = cs.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    grp = cs.DeviceGroup(list(range(ndev)))
    assert len(set(grp.devices)) == ndev
    ν = np.linspace(500.0, 900.0, 4001)
    P = cs.pressuregrid(10.0, 1e5, 16)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    Fup, Fdn, Fnet = cs.sharded_fluxes(grp, P, 9.8, Γ, 0.029, None, None, [(co2, 400e-6, "voigt", 25.0)], ν)
    F = cs.radiate(P, 9.8, Γ, 0.029, None, None, cs.LineGas(co2, 400e-6, ν, "voigt", 25.0))
    assert relerr(Fup, F.Fup) < 1e-12 and relerr(Fdn[1:], F.Fdn[1:]) < 1e-12
    # every device's reduced buffer holds the same sums
    got = [grp.read(i, 2 * len(P)) for i in range(ndev)]
    for g in got[1:]:
        assert np.array_equal(g, got[0])
    assert relerr(got[0][:len(P)], F.Fup) < 1e-12
    grp.close()


def test_line_params_and_functors(cs, orc, co2):
    """vector forms of scaleintensity / αdoppler / γlorentz (line_shapes.jl:125-132,146-148,259-261), the Gas and
    UnifiedAbsorber functors (gases.jl:256-281, absorbers.jl:97-99) and the CIATables functor / cia()"""
    import ctypes as C
    T, P, Pp = 233.0, 4e4, 16.0
    S, α, γ = cs.scaleintensity(co2, T), cs.αdoppler(co2, T), cs.γlorentz(co2, T, P, Pp)
    niso, ncheb, cheb, has = co2.cheb_table()
    idx = np.arange(0, co2.N, 37)
    Sref = np.array([orc.scalar("orc_scaleintensity", co2.S[j], co2.ν[j], co2.Epp[j], T, C.c_int(int(ncheb[co2.I[j] - 1])),
                                np.ascontiguousarray(cheb[co2.I[j] - 1]).ctypes.data_as(C.POINTER(C.c_double))) for j in idx])
    αref = np.array([orc.scalar("orc_alpha_doppler", co2.ν[j], co2.μ[j], T) for j in idx])
    γref = np.array([orc.scalar("orc_gamma_lorentz", co2.γa[j], co2.γs[j], co2.na[j], T, P, Pp) for j in idx])
    assert relerr(S[idx], Sref, 1e-300) < 1e-12 and relerr(α[idx], αref) < 1e-14 and relerr(γ[idx], γref) < 1e-13
    ν = np.linspace(600.0, 750.0, 301)
    Ω = cs.AtmosphericDomain((150, 300), 6, (10, 1e5), 8)
    gas = cs.Gas(co2, 400e-6, ν, Ω)
    r = gas.rawσ(250.0, 2e4)
    assert gas.rawσ(17, 250.0, 2e4) == r[17] and abs(gas(17, 250.0, 2e4) - 400e-6 * r[17]) < 1e-40
    gray = cs.GrayGas(1e-27, ν)
    U = cs.UnifiedAbsorber(gas, gray)
    assert relerr(U(250.0, 2e4), 400e-6 * r + 1e-27, 1e-300) < 1e-13
    g2 = gas.reconcentrate(1e-3)
    assert relerr(g2(250.0, 2e4), 1e-3 * r, 1e-300) < 1e-15
    x = cs.CIATables(os.path.join(DATA, "CO2-CO2_2018.cia.gz"), extrapolate=True)
    νc = np.linspace(5.0, 700.0, 97)
    assert relerr(x(νc, 250.0), orc.cia_k(x, νc, np.full(len(νc), 250.0)), 1e-300) < 1e-12
    σ = x.cia(νc, 250.0, 1e5, 9e4, 9e4)
    assert relerr(σ, orc.cia_nodes(x, νc, [250.0], [1e5], [0.9], [0.9])[0], 1e-300) < 1e-12


def test_par_ingestion_bit_exact(cs, tmp_path):
    """GPU .par parser (cs_par_parse) vs the host parser: bit-identical on the reference's three fixtures, on an
    exhaustive sweep of the E10.3 intensity format (all 4-digit mantissas x exponents -60..+20, i.e. through the
    double-double path), on random F12.6 / F10.4 / F5.4 fields, and on a written-out synthetic line list"""
    import gzip
    for name in ("CO2", "H2O", "CH4"):
        fn = os.path.join(DATA, f"{name}.par.gz")
        a, b = cs.readpar(fn), cs.readpar_b200(fn)
        for k in ("M", "I", "ν", "S", "A", "γa", "γs", "Epp", "na", "δa"):
            assert np.array_equal(a[k], b[k]), (name, k)
    a, b = cs.readpar(fn, νmin=1000, νmax=3000, Scut=1e-26, I=[1, 2], maxlines=500), \
        cs.readpar_b200(fn, νmin=1000, νmax=3000, Scut=1e-26, I=[1, 2], maxlines=500)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    # device-side filters / maxlines / sort (cs_par_read) against the host readpar over the keyword combinations, ties in S
    # and ν included (stable order), and the reference's pre-filter comparison of maxlines
    for kw in (dict(νmin=600.0, νmax=800.0), dict(Scut=1e-24), dict(I=["1"]), dict(I=[2, "3"], Scut=1e-27),
               dict(maxlines=100), dict(maxlines=100, νmin=2000.0), dict(maxlines=10**9), dict(νmin=600.0, νmax=2400.0, I=[1], maxlines=7)):
        for name in ("CO2", "H2O"):
            fn = os.path.join(DATA, f"{name}.par.gz")
            a, b = cs.readpar(fn, **kw), cs.readpar_b200(fn, **kw)
            assert all(a[k].shape == b[k].shape and np.array_equal(a[k], b[k]) for k in a), (name, kw)
    rec = cs.readpar_b200(fn, νmin=1000.0, maxlines=50, index=True)
    full, _, _ = cs.parse_records_b200(gzip.open(fn, "rb").read())           # file order
    assert np.array_equal(full["ν"][rec["record"]], rec["ν"]) and np.array_equal(full["S"][rec["record"]], rec["S"])
    with pytest.raises(AssertionError, match="filtered to nothing"):
        cs.readpar_b200(fn, νmin=1e6)
    # exhaustive E10.3 sweep
    recs, vals = [], []
    for e in range(-60, 21):
        for m in range(1000, 10000, 1):
            s = f"{m / 1000:.3f}E{e:+03d}"
            recs.append(s)
            vals.append(float(s))
    rng = np.random.default_rng(0)
    nu = [f"{x:12.6f}" for x in rng.uniform(0, 99999, len(recs))]
    ep = [f"{x:10.4f}" for x in rng.uniform(-1, 99999, len(recs))]
    ga = [("%.4f" % x)[1:] for x in rng.uniform(0, 0.9999, len(recs))]
    text = "".join(f" 21{n}{s.rjust(10)}{s.rjust(10)}{g}{g}{e}0.75-.001234".ljust(160) + "\n" for n, s, g, e in zip(nu, recs, ga, ep))
    par, flags, reclen = cs.parse_records_b200(text.encode("ascii"))
    assert reclen == 161 and not flags.any()
    assert np.array_equal(par["S"], np.array(vals)) and np.array_equal(par["A"], np.array(vals))
    assert np.array_equal(par["ν"], np.array([float(x) for x in nu]))
    assert np.array_equal(par["Epp"], np.array([float(x) for x in ep]))
    assert np.array_equal(par["γa"], np.array([float(x) for x in ga]))
    assert np.all(par["δa"] == -0.001234) and np.all(par["na"] == 0.75)
    # malformed field is flagged, not silently converted
    bad = text[:161 * 3].replace("E", "Q", 1).encode("ascii")
    _, fl, _ = cs.parse_records_b200(bad)
    assert fl[0] == 1 and not fl[1:].any()
    # write -> read round trip of a synthetic list (what tools hand to the Julia reference)
    sl = synthetic_lines(cs, 5000, seed=9)
    fn = str(tmp_path / "syn.par")
    cs.writepar(fn, sl)
    back = cs.SpectralLines.from_par(cs.readpar_b200(fn))
    for k in ("ν", "S", "γa", "γs", "Epp", "na"):
        assert np.array_equal(getattr(back, k), getattr(sl, k)), k
    assert np.array_equal(getattr(cs.SpectralLines.from_file(fn), "S"), sl.S)


def test_extreme_cutoffs_and_sizes(cs, orc, co2):
    """cut-off 0 (only exact coincidences count), a cut-off wider than the whole line list (every line is an interior
    line for every tile), many levels in one call, a one-line list, many pressure levels in the flux solver"""
    νl = co2.ν[(co2.ν > 660) & (co2.ν < 670)]
    ν = np.unique(np.concatenate([np.linspace(660, 670, 500), νl[:40]]))
    T, P = [250.0], [3e4]
    for cut in (0.0, 1e-3, 1e6):
        for shape, sid in (("voigt", 2), ("lorentz", 1), ("doppler", 0), ("PHCO2", 3)):
            got = cs.xsec(shape, ν, co2, T, P, [12.0], cut)
            ref = orc.xsec(sid, co2, ν, T, P, [12.0], cut)
            assert relerr(got, ref, 1e-290) < XSEC_TOL, (cut, shape)
    # 257 levels in one call
    Pn = cs.pressuregrid(10.0, 1e5, 257)
    Tn = np.linspace(150.0, 300.0, 257)
    ν2 = np.linspace(600.0, 760.0, 333)
    got = cs.xsec("voigt", ν2, co2, Tn, Pn, 400e-6 * Pn, 25.0)
    ref = orc.xsec(orc.VOIGT, co2, ν2, Tn, Pn, 400e-6 * Pn, 25.0, nthreads=0)
    assert relerr(got, ref, 1e-290) < XSEC_TOL
    # one line
    one = cs.SpectralLines(co2.name, co2.formula, 1, co2.M, co2.I[3000:3001], co2.μ[3000:3001], co2.A[3000:3001],
                           co2.ν[3000:3001], co2.S[3000:3001], co2.γa[3000:3001], co2.γs[3000:3001], co2.Epp[3000:3001],
                           co2.na[3000:3001])
    ν3 = co2.ν[3000] + np.linspace(-30, 30, 1201)
    assert relerr(cs.xsec("voigt", ν3, one, [200.0], [100.0], [0.04], 25.0),
                  orc.xsec(orc.VOIGT, one, ν3, [200.0], [100.0], [0.04], 25.0), 1e-300) < XSEC_TOL
    # 401 pressure levels through the flux solver
    P4 = cs.pressuregrid(10.0, 1e5, 401)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    gray = cs.GrayGas(3e-26, ν2)
    Fup, Fdn = cs.fluxes(P4, 9.8, Γ, 0.029, None, None, gray)
    m, W = cs.streamnodes(5)
    x, w = cs.lobattonodes(2)
    ref = orc.fluxes(ν2, P4, 2, w, np.full((400, 2), 0.029), Γ(P4), np.full((401, len(ν2)), 3e-26), 9.8, None, None, 0.841, 5, m, W,
                     full=False)
    assert relerr(Fup, ref["Fup"]) < FLUX_TOL and relerr(Fdn[1:], ref["Fdn"][1:]) < FLUX_TOL


def test_thousands_of_levels_vs_oracle(cs, orc, co2):
    """the flux solver beyond the former 1025-level cap (per-warp partial sums in global memory): 2049 levels with 4-point
    Lobatto layers on a real-gas table, monochromatic and integrated fluxes against the oracle's Discretized solve of the
    same refined grid (the grid the Radau-equivalent entry points refine to)"""
    ν = np.linspace(600.0, 760.0, 161)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    Ω = cs.AtmosphericDomain((140, 300), 8, (5, 1.1e5), 12)
    gas = cs.Gas(co2, 400e-6, ν, Ω, "voigt", 25.0)
    P = np.exp(np.linspace(np.log(10.0), np.log(1e5), 2049))
    fS, fa = (lambda x: 0.05 + 0 * x), 0.2
    Mup, Mdn = cs.monochromaticfluxes(P, 9.8, Γ, 0.029, fS, fa, gas, core=cs.Discretized(5, 4))
    Fup, Fdn = cs.fluxes(P, 9.8, Γ, 0.029, fS, fa, gas, core=cs.Discretized(5, 4))
    m, W = cs.streamnodes(5)
    x, w = cs.lobattonodes(4)
    Pn = P[:-1, None] + np.diff(P)[:, None] * x[None, :]
    nodesP = np.concatenate([Pn[:, :-1].ravel(), P[-1:]])
    ws = cs.SigmaWorkspace(ν, len(nodesP))
    cs.UnifiedAbsorber(gas).sigma_nodes(ws, Γ(nodesP), nodesP)
    ref = orc.fluxes(ν, P, 4, w, np.full((len(P) - 1, 4), 0.029), Γ(P), ws.read(), 9.8, np.full(len(ν), 0.05), np.full(len(ν), 0.2),
                     0.841, 5, m, W, nthreads=0)
    assert relerr(Mup, ref["Mup"], 1e-300) < FLUX_TOL and relerr(Mdn, ref["Mdn"], 1e-300) < FLUX_TOL
    assert relerr(Fup, ref["Fup"]) < FLUX_TOL and relerr(Fdn, ref["Fdn"]) < FLUX_TOL
    # and the refinement of the Radau-equivalent `outgoing` now reaches its tolerance instead of stopping at the cap
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        cs.outgoing(1e5, 9.8, Γ, 0.029, gas, Ptop=10.0, tol=1e-5)


def test_bench_line_contract(cs):
    """bench.py prints ONE JSON line with the keys the driver reads (small workload, N = 1)"""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c2small", "--steps", "2", "--warmup", "3",
                          "--cpu-evals", "2e8"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["dtype"] == "f64" and d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["value"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert "workload" in d["config"]
