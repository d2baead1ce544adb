#!/usr/bin/env python
"""ncu target: one 50 x 50 table fit on NNU wavenumbers (expansion mode keeps the bake's line sum short).
    ncu --set full --import-source on -k regex:table_fit_fused -c 1 -o gpurun_out/x python tools/k3_profile_target.py 200000"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200")); sys.path.insert(0, ROOT)
import bench
import clearsky_b200 as cs
nν = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
ctx = cs.default_context()
ctx.set_farfield("expansion")
sl = bench.synthetic_lines(cs, 250_000, 20261018, 2, (0.06, 0.13))
ν = (3000.0 / nν) * np.arange(1, nν + 1)
Ω = cs.AtmosphericDomain((150, 320), 50, (5, 1.1e5), 50)
gas = cs.Gas(sl, 400e-6, ν, Ω)
print("fit ms", gas.timers["table_fit"], "linesum ms", gas.timers["linesum"])
