#!/usr/bin/env python
"""ncu target: one C2 line sum (Voigt, both gases) at NLEV levels, direct or expansion mode.
    ncu --set full --import-source on -k regex:line_sum_kernel -c 1 -o gpurun_out/x python tools/k2_profile_target.py 11 direct"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200")); sys.path.insert(0, ROOT)
import bench
import clearsky_b200 as cs
from clearsky_b200._lib import check, f64, lib, ptr
nlev = int(sys.argv[1]) if len(sys.argv) > 1 else 11
mode = sys.argv[2] if len(sys.argv) > 2 else "direct"
wl = bench.make_workload(cs, "c2")
ctx = cs.default_context()
ctx.set_farfield(mode)
ν, P, T = wl["ν"], wl["P"], wl["T"]
k = np.linspace(0, len(P) - 1, nlev).astype(int)
P, T = f64(P[k]), f64(T[k])
ws = cs.SigmaWorkspace(ν, nlev, ctx)
for sl, C in wl["gases"]:
    dl = cs.DeviceLines(sl, ctx)
    check(lib().cs_sigma_add_lines(ws.h, dl.h, 2, ptr(T), ptr(P), ptr(f64(np.full(nlev, C))), 25.0))
ctx.synchronize()
print("linesum ms", ctx.timers_total()["linesum"])
