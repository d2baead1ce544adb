#!/usr/bin/env python
"""GPU-vs-oracle parity on EVERY configuration of BASELINE.json at its own parameters, on a bounded ν slice.

    python tools/config_parity.py [c2 c3 c4 c5 ...] [--points N]

The full configurations are hours of CPU work for the oracle, so each check keeps the configuration's line lists, shapes,
cut-offs, level / node grids and absorbers and restricts only the wavenumber axis to a contiguous slice (a slice sees
exactly the lines of the full run inside its window: the line sum is independent point by point).  Tolerances are the
north-star ones: 1e-9 on cross-sections, 1e-8 on fluxes.  tests/test_gpu_configs.py asserts the same functions under
`-m gpu`; tools/bench_configs.py prints them beside the timings.  TEST INFRASTRUCTURE: imports oracle/.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200"))
sys.path.insert(0, ROOT)

TOL_SIGMA, TOL_FLUX = 1e-9, 1e-8


def _rel(a, b, floor=1e-290):
    a, b = np.asarray(a), np.asarray(b)
    keep = np.abs(b) > floor
    return float(np.max(np.abs(a[keep] - b[keep]) / np.abs(b[keep]))) if keep.any() else 0.0


def _threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _slice(ν, n, where=0.5):
    n = min(n, len(ν))
    i0 = int((len(ν) - n) * where)
    return i0, np.ascontiguousarray(ν[i0:i0 + n])


def _modes(ctx, fn):
    """run fn() in both far-field modes -> {mode: result}"""
    out = {}
    try:
        for mode in ("direct", "expansion"):
            ctx.set_farfield(mode)
            out[mode] = fn()
    finally:
        ctx.set_farfield("direct")
    return out


def c2(points=1500, where=0.5):
    """configs[1]: Earth-like clear sky, 2 x 250k synthetic Voigt lines, 0.01 cm^-1 grid, 101 levels, cut-off 25"""
    import bench
    import clearsky_b200 as cs
    from oracle import oracle as orc
    orc.build()
    nth = _threads()
    wl = bench.make_workload(cs, "c2")
    i0, νs = _slice(wl["ν"], points, where)
    P, T = wl["P"], wl["T"]
    nlev = len(P)
    m, W = cs.streamnodes(wl["nstream"])
    x, w = cs.lobattonodes(wl["nlob"])
    μn = np.full((nlev - 1, wl["nlob"]), wl["μ"])
    σo = np.zeros((nlev, len(νs)))
    for sl, C in wl["gases"]:
        σo += C * orc.xsec(orc.VOIGT, sl, νs, T, P, C * P, wl["cut"], nthreads=nth)
    Fo = orc.fluxes(νs, P, wl["nlob"], w, μn, T, σo, wl["g"], None, None, 0.841, wl["nstream"], m, W, nthreads=nth, full=False)
    ctx = cs.default_context()
    gases = [cs.LineGas(sl, C, νs, "voigt", wl["cut"]) for sl, C in wl["gases"]]

    def run():
        ws = cs.SigmaWorkspace(νs, nlev)
        cs.UnifiedAbsorber(*gases).sigma_nodes(ws, T, P)
        Fup, Fdn = cs.fluxes(P, wl["g"], cs.AtmosphericProfile(P, T), wl["μ"], None, None, *gases)
        return ws.read(), Fup, Fdn

    res = {"config": "c2", "n_nu": len(νs), "first_index": i0, "levels": nlev}
    for mode, (Σ, Fup, Fdn) in _modes(ctx, run).items():
        res[f"max_rel_sigma_{mode}"] = _rel(Σ, σo)
        res[f"max_rel_flux_{mode}"] = max(_rel(Fup, Fo["Fup"]), _rel(Fdn[1:], Fo["Fdn"][1:]))
    return _verdict(res)


def c3(points=256, where=0.22):
    """configs[2]: early-Mars 2 bar pure CO2, PHCO2 (cut-off 500 cm^-1) + CO2-CO2 CIA with extrapolate=true, 101 levels"""
    import bench
    import clearsky_b200 as cs
    from oracle import oracle as orc
    orc.build()
    nth = _threads()
    co2 = bench.synthetic_lines(cs, 500_000, 20261019 + 1, 2, (0.06, 0.13))
    ν = 0.01 * np.arange(1, 300_001)
    i0, νs = _slice(ν, points, where)          # around 660 cm^-1: inside the CIA tables and the 15 micron band region
    P = cs.pressuregrid(10.0, 2e5, 101)
    Γ = cs.DryAdiabat(250.0, 2e5, 770.0, 0.044, Ptropo=1e4)
    T = Γ(P)
    nlev = len(P)
    x = cs.CIATables(os.path.join(ROOT, "tests", "data", "CO2-CO2_2018.cia.gz"), extrapolate=True)
    one = np.ones(nlev)
    σo = orc.xsec(orc.PHCO2, co2, νs, T, P, 1.0 * P, 500.0, nthreads=nth) + orc.cia_nodes(x, νs, T, P, one, one)
    m, W = cs.streamnodes(5)
    _, w = cs.lobattonodes(2)
    Fo = orc.fluxes(νs, P, 2, w, np.full((nlev - 1, 2), 0.044), T, σo, 3.71, None, None, 0.841, 5, m, W, nthreads=nth, full=False)
    ctx = cs.default_context()
    gas = cs.LineGas(co2, 1.0, νs, "PHCO2", 500.0)

    def run():
        ws = cs.SigmaWorkspace(νs, nlev)
        cs.UnifiedAbsorber(gas, x).sigma_nodes(ws, T, P)
        Fup, Fdn = cs.fluxes(P, 3.71, Γ, 0.044, None, None, gas, x)
        return ws.read(), Fup, Fdn

    res = {"config": "c3", "n_nu": len(νs), "first_index": i0, "levels": nlev, "shape": "PHCO2", "cutoff": 500.0,
           "cia": "CO2-CO2 extrapolate=true"}
    for mode, (Σ, Fup, Fdn) in _modes(ctx, run).items():
        res[f"max_rel_sigma_{mode}"] = _rel(Σ, σo)
        res[f"max_rel_flux_{mode}"] = max(_rel(Fup, Fo["Fup"]), _rel(Fdn[1:], Fo["Fdn"][1:]))
    return _verdict(res)


def c4(points=256, where=0.4):
    """configs[3]: OpacityTable build on 50 T x 50 P nodes (CO2 and H2O, 250k lines each, nu step 0.003), the Bichebyshev
    fit, and the interpolated sweep at the 101 levels of the C2 atmosphere"""
    import bench
    import clearsky_b200 as cs
    from oracle import oracle as orc
    orc.build()
    nth = _threads()
    nT = nP = 50
    ν = 0.003 * np.arange(1, 1_000_001)
    i0, νs = _slice(ν, points, where)
    Ω = cs.AtmosphericDomain((150, 320), nT, (5, 1.1e5), nP)
    P = cs.pressuregrid(10.0, 1e5, 101)
    T = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)(P)
    ctx = cs.default_context()
    res = {"config": "c4", "n_nu": len(νs), "first_index": i0, "nT": nT, "nP": nP, "levels": len(P), "gases": {}}
    worst = {}
    for M, C, seed, rng in ((2, 400e-6, 20261018, (0.06, 0.13)), (1, 1e-3, 20261019, (0.10, 0.50))):
        sl = bench.synthetic_lines(cs, 250_000, seed, M, rng)
        blk, nz = orc.bake(orc.VOIGT, sl, νs, Ω.T, Ω.P, np.full((nP, nT), C), 25.0, nthreads=nth)
        A = orc.table_fit(blk)
        σo = orc.gas_nodes(A, Ω.T, Ω.P, T, P, np.ones(len(P)), nthreads=nth)

        def run():
            gas = cs.Gas(sl, C, νs, Ω, keep_block=True)
            return gas.σblock(), gas.rawσ(T, P), gas.nzeroed

        g = {}
        for mode, (b, σ, nzg) in _modes(ctx, run).items():
            g[f"max_rel_block_{mode}"] = _rel(b, blk)
            g[f"max_rel_sigma_{mode}"] = _rel(σ, σo)
            g[f"nzeroed_equal_{mode}"] = bool(nzg == nz)
            for k in ("block", "sigma"):
                worst[f"max_rel_{k}_{mode}"] = max(worst.get(f"max_rel_{k}_{mode}", 0.0), g[f"max_rel_{k}_{mode}"])
        res["gases"][sl.formula] = g
    res.update(worst)
    return _verdict(res)


def c5(points=1200, where=0.23, steps=4):
    """configs[4]: radiative-convective loop on the C2 atmosphere: 12 x 24 tables on C2's nu grid, AcceleratedAbsorber at the
    51 cell edges, radmul = 2 -> 101 radiative levels, device-resident steps against the oracle-driven host twin"""
    import bench
    import clearsky_b200 as cs
    from oracle import oracle as orc
    orc.build()
    nth = _threads()
    ν = 0.01 * np.arange(1, 300_001)
    i0, νs = _slice(ν, points, where)
    Ω = cs.AtmosphericDomain((140, 320), 12, (5, 1.1e5), 24)
    l1 = bench.synthetic_lines(cs, 250_000, 20261018, 2, (0.06, 0.13))
    l2 = bench.synthetic_lines(cs, 250_000, 20261019, 1, (0.10, 0.50))
    Pe = cs.pressuregrid(10.0, 1e5, 51)
    Te = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)(Pe)
    # oracle: tables -> Sigma at the edges -> accelerated absorber at the radiative levels
    σe = np.zeros((len(Pe), len(νs)))
    for sl, C in ((l1, 400e-6), (l2, 1e-3)):
        blk, _ = orc.bake(orc.VOIGT, sl, νs, Ω.T, Ω.P, np.full((Ω.nP, Ω.nT), C), 25.0, nthreads=nth)
        σe += orc.gas_nodes(orc.table_fit(blk), Ω.T, Ω.P, Te, Pe, np.full(len(Pe), C), nthreads=nth)
    lnσ = np.maximum(np.log(σe), np.log(np.finfo(float).tiny))
    co2, h2o = cs.Gas(l1, 400e-6, νs, Ω), cs.Gas(l2, 1e-3, νs, Ω)
    rcm = cs.RCM(Pe, Te, 9.8, 0.029, None, None, 1040.0, 1e7, co2, h2o, radmul=2)
    σr = orc.accel_nodes(np.log(Pe), lnσ, rcm.Pr)
    nr = len(rcm.Pr)
    ws = cs.SigmaWorkspace(νs, nr)
    rcm.A.sigma_nodes(ws, rcm.Pr * 0 + 250.0, rcm.Pr)
    res = {"config": "c5", "n_nu": len(νs), "first_index": i0, "nrad": nr, "steps": steps,
           "max_rel_sigma_direct": _rel(ws.read(), σr)}
    m, W = cs.streamnodes(5)
    _, w = cs.lobattonodes(2)
    P, T = rcm.P.copy(), rcm.T.copy()
    eH = eF = 0.0
    for _ in range(steps):
        Tlev = cs.AtmosphericProfile(P, T)(rcm.Pr)
        f = orc.fluxes(νs, rcm.Pr, 2, w, np.full((nr - 1, 2), 0.029), Tlev, σr, 9.8, None, None, 0.841, 5, m, W, nthreads=nth, full=False)
        R = -cs.AtmosphericProfile(rcm.Pr, f["Fnet"])(Pe)
        H = np.empty(len(Pe))
        H[:-1] = (9.8 / 1040.0) * (R[:-1] - R[1:]) / (Pe[1:] - Pe[:-1])
        H[-1] = R[-1] / 1e7
        T = T + 600.0 * H
        rcm.steps_(600.0, 1)
        eH = max(eH, float(np.max(np.abs(rcm.H - H)) / np.max(np.abs(H))))
        eF = max(eF, _rel(rcm.F.Fup, f["Fup"]), _rel(rcm.F.Fdn[1:], f["Fdn"][1:]))
    res["max_rel_flux_direct"] = eF
    res["max_rel_heating"] = eH
    res["max_rel_T"] = _rel(rcm.T, T)
    rcm.close()
    return _verdict(res)


def _verdict(res):
    ok = True
    for k, v in res.items():
        if k.startswith("max_rel_sigma") or k.startswith("max_rel_block"):
            ok &= v <= TOL_SIGMA
        elif k.startswith("max_rel_flux") or k in ("max_rel_heating", "max_rel_T"):
            ok &= v <= TOL_FLUX
        elif k.startswith("nzeroed_equal"):
            ok &= bool(v)
    res["tol_sigma"], res["tol_flux"], res["ok"] = TOL_SIGMA, TOL_FLUX, bool(ok)
    return res


ALL = {"c2": c2, "c3": c3, "c4": c4, "c5": c5}

if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or list(ALL)
    kw = {}
    if "--points" in sys.argv:
        kw["points"] = int(sys.argv[sys.argv.index("--points") + 1])
    rc = 0
    for name in which:
        r = ALL[name](**kw)
        print(json.dumps(r))
        rc |= 0 if r["ok"] else 1
    sys.exit(rc)
