#!/usr/bin/env python
"""host wall-clock breakdown of bench.py's end-to-end step (C2, pinned host buffers): where the milliseconds between the
device-timed step and the e2e step go"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200")); sys.path.insert(0, ROOT)
import torch
import bench
import clearsky_b200 as cs
from clearsky_b200._lib import check, f64, lib, ptr

wl = bench.make_workload(cs, "c2")
ctx = cs.default_context()
ν, P, T = wl["ν"], wl["P"], wl["T"]
nlev = len(P)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
gases = [(cs.SpectralLines(sl.name, sl.formula, sl.N, sl.M, pin(sl.I), pin(sl.μ), pin(sl.A), pin(sl.ν), pin(sl.S), pin(sl.γa), pin(sl.γs),
                           pin(sl.Epp), pin(sl.na)), C) for sl, C in wl["gases"]]
νp, wts = pin(ν), pin(bench.trapz_weights(ν))
m, W = cs.streamnodes(5); x, w = cs.lobattonodes(2)
m, W, w = f64(m), f64(W), f64(w)
μn = f64(np.full((nlev - 1, 2), 0.029)); Tn, Pn = f64(T), f64(P)
dF = torch.zeros(2 * nlev, dtype=torch.float64, device="cuda:0")
Fh = torch.empty(2 * nlev, dtype=torch.float64).pin_memory()
for it in range(5):
    t = [time.perf_counter()]
    for sl, _ in gases:
        sl.__dict__.pop("_dev", None)
    lg = [cs.LineGas(sl, C, νp, "voigt", 25.0, ctx=ctx) for sl, C in gases]; t.append(time.perf_counter())
    A = cs.UnifiedAbsorber(*lg); t.append(time.perf_counter())
    ws = cs.SigmaWorkspace(νp, nlev, ctx); t.append(time.perf_counter())
    A.sigma_nodes(ws, Tn, Pn); t.append(time.perf_counter())
    check(lib().cs_fluxes_device(ws.h, nlev, ptr(Pn), 2, ptr(w), ptr(μn), ptr(Tn), 9.8, None, None, 0.841, 5, ptr(m), ptr(W), ptr(wts),
                                 dF.data_ptr())); t.append(time.perf_counter())
    Fh.copy_(dF); t.append(time.perf_counter())
    del ws, A, lg; t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(f"it {it}: LineGas x2 {d[0]:.2f}  Unified {d[1]:.2f}  workspace {d[2]:.2f}  sigma_nodes(enqueue) {d[3]:.2f}  fluxes(enqueue) {d[4]:.2f}  "
          f"D2H+wait {d[5]:.2f}  free {d[6]:.2f}  total {sum(d):.2f} ms")
