#!/usr/bin/env python
"""host-side profile (cProfile) of the RCM step loop on config 5: where the time outside the kernels goes"""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200")); sys.path.insert(0, ROOT)
import numpy as np
import clearsky_b200 as cs
import bench
n, nν = 250_000, 300_000
ν = 0.01 * np.arange(1, nν + 1)
Ω = cs.AtmosphericDomain((140, 320), 12, (5, 1.1e5), 24)
co2 = cs.Gas(bench.synthetic_lines(cs, n, 20261018, 2, (0.06, 0.13)), 400e-6, ν, Ω)
h2o = cs.Gas(bench.synthetic_lines(cs, n, 20261019, 1, (0.10, 0.50)), 1e-3, ν, Ω)
Pe = cs.pressuregrid(10.0, 1e5, 51)
Te = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)(Pe)
rcm = cs.RCM(Pe, Te, 9.8, 0.029, None, None, 1040.0, 1e7, co2, h2o, radmul=2)
for _ in range(5):
    rcm.step_(600.0)
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    rcm.step_(600.0)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
