#!/usr/bin/env python
"""host wall-clock breakdown of one bench step (C2), per C-ABI call"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200")); sys.path.insert(0, ROOT)
import clearsky_b200 as cs
from clearsky_b200._lib import check, f64, lib, ptr
import bench

wl = bench.make_workload(cs, sys.argv[1] if len(sys.argv) > 1 else "c2")
ctx = cs.Context(0)
ν, P, T = wl["ν"], wl["P"], wl["T"]
if os.environ.get("NLEV"):
    k = np.linspace(0, len(P) - 1, int(os.environ["NLEV"])).astype(int)
    P, T = P[k], T[k]
nlev = len(P)
dls = [cs.DeviceLines(sl, ctx) for sl, _ in wl["gases"]]
ws = cs.SigmaWorkspace(ν, nlev, ctx)
Cs = [f64(np.full(nlev, C)) for _, C in wl["gases"]]
m, W = cs.streamnodes(5); x, w = cs.lobattonodes(2)
μn = f64(np.full((nlev - 1, 2), 0.029)); Tn, Pn = f64(T), f64(P)
Fu, Fd, Fn = np.empty(nlev), np.empty(nlev), np.empty(nlev)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t = [time.perf_counter()]
    ws.zero(); ctx.synchronize(); t.append(time.perf_counter())
    for dl, C in zip(dls, Cs):
        check(lib().cs_sigma_add_lines(ws.h, dl.h, 2, ptr(Tn), ptr(Pn), ptr(C), 25.0)); t.append(time.perf_counter())
    check(lib().cs_fluxes(ws.h, nlev, ptr(Pn), 2, ptr(f64(w)), ptr(μn), ptr(Tn), 9.8, None, None, 0.841, 5, ptr(f64(m)), ptr(f64(W)),
                          None, None, None, None, ptr(Fu), ptr(Fd), ptr(Fn))); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(f"iter {it}: zero {d[0]:.2f} ms, add_lines {d[1]:.2f} + {d[2]:.2f} ms, fluxes {d[3]:.2f} ms, total {sum(d):.2f}; timers {ctx.timers()}")
print("OLR", Fu[0])

# ---- e2e pieces (host buffers -> result)
import copy
for it in range(4):
    t = [time.perf_counter()]
    dls2 = []
    for sl, _ in wl["gases"]:
        c = copy.copy(sl); c.__dict__.pop("_dev", None)
        dls2.append(cs.DeviceLines(c, ctx))
    t.append(time.perf_counter())
    ws2 = cs.SigmaWorkspace(ν, nlev, ctx); t.append(time.perf_counter())
    tm0 = ctx.timers()
    for dl, C in zip(dls2, Cs):
        check(lib().cs_sigma_add_lines(ws2.h, dl.h, 2, ptr(Tn), ptr(Pn), ptr(C), 25.0))
    t.append(time.perf_counter())
    tm1 = ctx.timers()
    print("   kernel deltas: linesum", tm1["linesum"] - tm0["linesum"], "prep", tm1["prep"] - tm0["prep"])
    check(lib().cs_fluxes(ws2.h, nlev, ptr(Pn), 2, ptr(f64(w)), ptr(μn), ptr(Tn), 9.8, None, None, 0.841, 5, ptr(f64(m)), ptr(f64(W)),
                          None, None, None, None, ptr(Fu), ptr(Fd), ptr(Fn))); t.append(time.perf_counter())
    del dls2, ws2; t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(f"e2e {it}: upload lines {d[0]:.2f} ms, workspace {d[1]:.2f} ms, add_lines {d[2]:.2f} ms, fluxes {d[3]:.2f} ms, free {d[4]:.2f} ms")
