#!/usr/bin/env python
"""per-slice line-sum kernel time of the N-way ν sharding, measured sequentially on one GPU"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200")); sys.path.insert(0, ROOT)
import clearsky_b200 as cs
from clearsky_b200._lib import check, f64, lib, ptr
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = bench.make_workload(cs, "c2")
ν, P, T = wl["ν"], wl["P"], wl["T"]; nlev = len(P); cut = 25.0
ctx = cs.default_context()
from clearsky_b200 import sharding
counts = (sharding.slice_cost(ν, [sl.ν for sl, _ in wl["gases"]], cut, farfield=ctx.get_farfield()) if os.environ.get("COSTMODEL", "1") == "1"
          else sum(bench.per_point_counts(ν, sl.ν, cut) for sl, _ in wl["gases"]))
print("far-field mode:", ctx.get_farfield(), "cost model:", os.environ.get("COSTMODEL", "1"))
edges = bench.balanced_slices(counts, N)
Tn, Pn = f64(T), f64(P)
res = []
for r in range(N):
    a, b = edges[r], edges[r + 1]
    νs = np.ascontiguousarray(ν[a:b])
    ws = cs.SigmaWorkspace(νs, nlev, ctx)
    tot = 0.0
    for sl, C in wl["gases"]:
        keep = (sl.ν >= νs[0] - cut - 1e-9) & (sl.ν <= νs[-1] + cut + 1e-9)
        s2 = cs.SpectralLines(sl.name, sl.formula, int(keep.sum()), sl.M, sl.I[keep], sl.μ[keep], sl.A[keep], sl.ν[keep], sl.S[keep], sl.γa[keep], sl.γs[keep], sl.Epp[keep], sl.na[keep])
        dl = cs.DeviceLines(s2, ctx)
        for it in range(2):
            t0 = ctx.timers()["linesum"]
            check(lib().cs_sigma_add_lines(ws.h, dl.h, 2, ptr(Tn), ptr(Pn), ptr(f64(np.full(nlev, C))), cut))
            dt = ctx.timers()["linesum"] - t0
        tot += dt
    res.append(tot)
    print(f"slice {r}: ν {νs[0]:.2f}-{νs[-1]:.2f} ({b-a} pts) linesum {tot:.2f} ms")
print("max/mean", max(res) / np.mean(res), "sum", sum(res))
