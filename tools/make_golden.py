#!/usr/bin/env python
"""Generate tests/golden/*.npz from the CPU oracle.

The reference (Julia) cannot run in this image and its tests hold no golden vectors for the hot path, so
these fixtures are ORACLE outputs ("parity unpinned", see oracle/oracle.c).  They freeze the oracle against
drift and give the GPU tests a fixed target that does not depend on rebuilding the oracle.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import clearsky_b200 as cs  # noqa: E402  (host-side readers / profiles only; no GPU call is made)
from helpers import c1_problem  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    co2 = cs.SpectralLines.from_file(os.path.join(ROOT, "tests", "data", "CO2.par.gz"))
    ν, P, Γ = c1_problem(cs)
    T = Γ(P)
    C = 400e-6
    g = {}
    g["nu"], g["P"], g["T"] = ν, P, T
    g["sigma_voigt"] = orc.xsec(orc.VOIGT, co2, ν, T[::5], P[::5], C * P[::5], 25.0)
    g["sigma_lorentz"] = orc.xsec(orc.LORENTZ, co2, ν, T[::5], P[::5], C * P[::5], 25.0)
    g["sigma_doppler"] = orc.xsec(orc.DOPPLER, co2, ν, T[::5], P[::5], C * P[::5], 25.0)
    g["sigma_phco2"] = orc.xsec(orc.PHCO2, co2, ν, T[::10], P[::10], P[::10], 500.0)
    # config 1, exact line-by-line gas at the levels (nlobatto = 2 -> nodes are the levels)
    σ = orc.xsec(orc.VOIGT, co2, ν, T, P, C * P, 25.0)
    m, W = cs.streamnodes(5)
    x, w = cs.lobattonodes(2)
    L = len(P) - 1
    μn = np.full((L, 2), 0.029)
    f = orc.fluxes(ν, P, 2, w, μn, T, C * σ, 9.8, None, None, 0.841, 5, m, W)
    g["lbl_Fup"], g["lbl_Fdn"], g["lbl_tau"] = f["Fup"], f["Fdn"], f["τ"]
    g["lbl_Mup_toa"] = f["Mup"][:, 0]
    # config 1 through a 12 x 24 opacity table
    Ω = cs.AtmosphericDomain((140, 300), 12, (5, 1.1e5), 24)
    Cg = np.full((Ω.nP, Ω.nT), C)
    block, nz = orc.bake(orc.VOIGT, co2, ν, Ω.T, Ω.P, Cg, 25.0, nthreads=0)
    A = orc.table_fit(block)
    σt = orc.gas_nodes(A, Ω.T, Ω.P, T, P, np.full(len(P), C))
    f = orc.fluxes(ν, P, 2, w, μn, T, σt, 9.8, None, None, 0.841, 5, m, W)
    g["tab_Fup"], g["tab_Fdn"] = f["Fup"], f["Fdn"]
    g["tab_sigma_lev"] = σt[::5]
    g["tab_nzeroed"] = np.array([nz])
    np.savez_compressed(os.path.join(out, "c1_co2.npz"), **g)
    print("wrote", os.path.join(out, "c1_co2.npz"), "OLR lbl/table:", g["lbl_Fup"][0], g["tab_Fup"][0], "zeroed", nz)


if __name__ == "__main__":
    main()
