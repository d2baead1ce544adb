#!/usr/bin/env python
"""Secondary configurations of BASELINE.json (configs[2..4]) on ONE GPU -- not the headline bench line.

    python tools/bench_configs.py c3|c4|c5 [--small]

Prints one JSON line per configuration (kept under profiles/).  All timings are wall-clock around synchronous
C-ABI calls plus the library's own CUDA-event kernel timers.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import clearsky_b200 as cs  # noqa: E402


def c3(small):
    """Early-Mars 2 bar CO2, PHCO2 sub-Lorentzian shape (cut-off 500 cm^-1) + CO2-CO2 CIA, 100 layers"""
    n, nν = (20_000, 12_000) if small else (500_000, 300_000)
    co2 = bench.synthetic_lines(cs, n, 20261019 + 1, 2, (0.06, 0.13))
    ν = 0.01 * np.arange(1, nν + 1) * (300_000 / nν)
    P = cs.pressuregrid(10.0, 2e5, 101)
    Γ = cs.DryAdiabat(250.0, 2e5, 770.0, 0.044, Ptropo=1e4)
    x = cs.CIATables(os.path.join(ROOT, "tests", "data", "CO2-CO2_2018.cia.gz"), extrapolate=True)
    gas = cs.LineGas(co2, 1.0, ν, "PHCO2", 500.0)
    evals = gas.evals_per_node() * len(P)
    ctx = cs.default_context()
    out = []
    for it in range(2):
        t0 = time.perf_counter()
        tm0 = ctx.timers()
        Fup, Fdn = cs.fluxes(P, 3.71, Γ, 0.044, None, None, gas, x)
        dt = time.perf_counter() - t0
        tm1 = ctx.timers()
        out.append((dt, tm1["linesum"] - tm0["linesum"], tm1["cia"] - tm0["cia"]))
    dt, ls, cia = out[-1]
    return {"config": "c3", "lines": n, "n_nu": nν, "levels": len(P), "shape": "PHCO2", "cutoff": 500.0, "evals": evals,
            "s_per_spectrum": dt, "evals_per_s": evals / dt, "linesum_ms": ls, "cia_ms": cia, "olr": float(Fup[0]),
            "algorithmic_tflops": 16.0 * evals / (ls * 1e-3) / 1e12 if ls else None}


def c3sharded(small):
    """C3 nu-sharded over all visible GPUs from one process (cs_group + NCCL all-reduce)"""
    n, nν = (20_000, 12_000) if small else (500_000, 300_000)
    co2 = bench.synthetic_lines(cs, n, 20261019 + 1, 2, (0.06, 0.13))
    ν = 0.01 * np.arange(1, nν + 1) * (300_000 / nν)
    P = cs.pressuregrid(10.0, 2e5, 101)
    Γ = cs.DryAdiabat(250.0, 2e5, 770.0, 0.044, Ptropo=1e4)
    x = cs.CIATables(os.path.join(ROOT, "tests", "data", "CO2-CO2_2018.cia.gz"), extrapolate=True)
    grp = cs.DeviceGroup()
    sh = cs.ShardedLineByLine(grp, [(co2, 1.0, "PHCO2", 500.0)], ν, cia=[(x, 0, 0)])
    evals = sum(p["gases"][0].evals_per_node() for p in sh.parts if p is not None) * len(P)
    ts = []
    for it in range(3):
        t0 = time.perf_counter()
        Fup, Fdn, Fnet = sh.fluxes(P, 3.71, Γ, 0.044)
        ts.append(time.perf_counter() - t0)
    return {"config": "c3-sharded", "n_gpus": len(grp), "lines": n, "n_nu": nν, "levels": len(P), "evals": evals,
            "s_per_spectrum": min(ts[1:]), "evals_per_s": evals / min(ts[1:]), "olr": float(Fup[0])}


def c4(small):
    """OpacityTable build: 50 T x 50 P x 1e6 nu for CO2 and H2O, then interpolated sweep at 101 levels"""
    n, nν, nT, nP = (20_000, 40_000, 12, 12) if small else (250_000, 1_000_000, 50, 50)
    res = []
    ctx = cs.default_context()
    ν = (3000.0 / nν) * np.arange(1, nν + 1)
    Ω = cs.AtmosphericDomain((150, 320), nT, (5, 1.1e5), nP)
    P = cs.pressuregrid(10.0, 1e5, 101)
    T = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)(P)
    for M, C, seed, rng in ((2, 400e-6, 20261018, (0.06, 0.13)), (1, 1e-3, 20261019, (0.10, 0.50))):
        sl = bench.synthetic_lines(cs, n, seed, M, rng)
        evals = cs.device_lines(sl).count_evals(ν, 25.0) * nT * nP
        t0 = time.perf_counter()
        gas = cs.Gas(sl, C, ν, Ω)
        dt = time.perf_counter() - t0
        tm = gas.timers
        t0 = time.perf_counter()
        σ = gas.rawσ(T, P)
        dte = time.perf_counter() - t0
        tme = ctx.timers()
        nk = nT * nP
        res.append({"gas": sl.formula, "bake_s": dt, "evals": evals, "evals_per_s": evals / dt, "linesum_ms": tm["linesum"],
                    "prep_ms": tm["prep"], "fit_ms": tm["table_fit"],
                    "fit_algorithmic_GBps": 2 * nν * nk * 8 / (tm["table_fit"] * 1e-3) / 1e9,      # block read once + coefficients written once
                    "fit_tflops": 2.0 * nν * nk * (nT + nP) / (tm["table_fit"] * 1e-3) / 1e12,
                    "eval_101_levels_s": dte, "eval_kernel_ms": tme["table_eval"],
                    "eval_tflops": 2.0 * nν * nk * len(P) / (tme["table_eval"] * 1e-3) / 1e12, "nzeroed": gas.nzeroed,
                    "sigma_check": float(σ[50, nν // 2])})
        del gas
    return {"config": "c4", "nT": nT, "nP": nP, "n_nu": nν, "lines_per_gas": n, "gases": res}


def c5(small):
    """radiative-convective loop: repeated heating! (= radiate! with the AcceleratedAbsorber) on one GPU"""
    n, nν, steps = (20_000, 30_000, 20) if small else (250_000, 300_000, 100)
    ν = 0.01 * np.arange(1, nν + 1) * (300_000 / nν)
    Ω = cs.AtmosphericDomain((140, 320), 12, (5, 1.1e5), 24)
    co2 = cs.Gas(bench.synthetic_lines(cs, n, 20261018, 2, (0.06, 0.13)), 400e-6, ν, Ω)
    h2o = cs.Gas(bench.synthetic_lines(cs, n, 20261019, 1, (0.10, 0.50)), 1e-3, ν, Ω)
    Pe = cs.pressuregrid(10.0, 1e5, 51)
    Te = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)(Pe)
    t0 = time.perf_counter()
    rcm = cs.RCM(Pe, Te, 9.8, 0.029, None, None, 1040.0, 1e7, co2, h2o, radmul=2)
    tsetup = time.perf_counter() - t0
    rcm.step_(600.0)
    t0 = time.perf_counter()
    for _ in range(steps):
        rcm.step_(600.0)
    dt = time.perf_counter() - t0
    tm = cs.default_context().timers()
    # jacobian!: np+1 = 52 flux solves, as one batched call and as the reference's loop
    rcm.jacobian_(1.0)
    t0 = time.perf_counter()
    rcm.jacobian_(1.0)
    tjb = time.perf_counter() - t0
    tjk = cs.default_context().timers()["rt"]
    t0 = time.perf_counter()
    rcm.jacobian_(1.0, batched=False)
    tjl = time.perf_counter() - t0
    return {"config": "c5", "n_nu": nν, "nrad": len(rcm.Pr), "steps": steps, "steps_per_s": steps / dt, "ms_per_step": dt / steps * 1e3,
            "rt_kernel_ms": tm["rt"], "reduce_kernel_ms": tm["reduce"], "accelerated_absorber_setup_s": tsetup,
            "OLR": float(rcm.F.Fup[0]), "Tsurf": float(rcm.T[-1]),
            "jacobian_profiles": rcm.np + 1, "jacobian_batched_ms": tjb * 1e3, "jacobian_batched_kernels_ms": tjk,
            "jacobian_loop_ms": tjl * 1e3}


def c5sharded(small):
    """config 5 as BASELINE.json words it: the radiative-convective loop on all visible GPUs.  One process (cs_group):
    opacity tables are baked per ν slice on the GPU that owns it, the AcceleratedAbsorber of every slice stays resident
    there, and every heating! is Σ + K6/K7 per device and one all-reduce of the 2·np fluxes"""
    n, nν, steps = (20_000, 30_000, 20) if small else (250_000, 300_000, 100)
    ν = 0.01 * np.arange(1, nν + 1) * (300_000 / nν)
    Ω = cs.AtmosphericDomain((140, 320), 12, (5, 1.1e5), 24)
    l1 = bench.synthetic_lines(cs, n, 20261018, 2, (0.06, 0.13))
    l2 = bench.synthetic_lines(cs, n, 20261019, 1, (0.10, 0.50))
    grp = cs.DeviceGroup()
    t0 = time.perf_counter()
    sh = cs.ShardedAbsorber(grp, ν, lambda νs, ctx: (cs.Gas(l1, 400e-6, νs, Ω, ctx=ctx), cs.Gas(l2, 1e-3, νs, Ω, ctx=ctx)))
    tbake = time.perf_counter() - t0
    Pe = cs.pressuregrid(10.0, 1e5, 51)
    Te = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)(Pe)
    t0 = time.perf_counter()
    rcm = cs.RCM(Pe, Te, 9.8, 0.029, None, None, 1040.0, 1e7, sh, radmul=2)
    tsetup = time.perf_counter() - t0
    rcm.step_(600.0)
    t0 = time.perf_counter()
    for _ in range(steps):
        rcm.step_(600.0)
    dt = time.perf_counter() - t0
    return {"config": "c5-sharded", "n_gpus": len(grp), "n_nu": nν, "nrad": len(rcm.Pr), "steps": steps, "steps_per_s": steps / dt,
            "ms_per_step": dt / steps * 1e3, "bake_s": tbake, "accelerated_absorber_setup_s": tsetup,
            "OLR": float(rcm.F.Fup[0]), "Tsurf": float(rcm.T[-1])}


def par(small):
    """.par ingestion: 250k synthetic CO2 lines written as 160-column records (40 MB): host readpar vs the one-call device
    readpar (cs_par_read: pinned double-buffered H2D of the text, parse, filter, maxlines, stable sort by nu on the GPU)"""
    import tempfile
    n = 20_000 if small else 250_000
    sl = bench.synthetic_lines(cs, n, 20261018, 2, (0.06, 0.13))
    fn = os.path.join(tempfile.mkdtemp(), "syn_co2.par")
    cs.writepar(fn, sl)
    nbytes = os.path.getsize(fn)
    ctx = cs.default_context()
    out = {"config": "par", "lines": n, "bytes": nbytes}
    for tag, kw in (("all", {}), ("filtered", dict(νmin=500.0, νmax=2500.0, Scut=1e-27, maxlines=100_000))):
        t0 = time.perf_counter(); a = cs.readpar(fn, **kw); th = time.perf_counter() - t0
        cs.readpar_b200(fn, **kw)
        best = None
        for _ in range(3):
            tm = {}
            t0 = time.perf_counter(); b = cs.readpar_b200(fn, timing=tm, **kw); tg = time.perf_counter() - t0
            tm["kernels_ms"] = ctx.timers()["total"]
            tm["wall_s"] = tg
            if best is None or tm["call_s"] < best["call_s"]:
                best = tm
        out[tag] = {"rows_out": int(len(b["ν"])), "bit_identical": bool(all(np.array_equal(a[k], b[k]) for k in a)), "host_readpar_s": th,
                    "gpu_readpar_s": best["wall_s"], "file_read_s": best["read_s"], "device_call_s": best["call_s"],
                    "kernels_ms": best["kernels_ms"], "call_GBps_text": nbytes / best["call_s"] / 1e9}
    buf = open(fn, "rb").read()
    cs.parse_records_b200(buf)
    t0 = time.perf_counter(); cs.parse_records_b200(buf); tp = time.perf_counter() - t0
    out["parse_only_call_s"] = tp
    out["parse_kernel_ms"] = ctx.timers()["total"]
    return out


if __name__ == "__main__":
    which = sys.argv[1]
    small = "--small" in sys.argv
    print(json.dumps({"c3": c3, "c3sharded": c3sharded, "c4": c4, "c5": c5, "c5sharded": c5sharded, "par": par}[which](small)))
