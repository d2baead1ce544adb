#!/usr/bin/env python
"""Harness self-test for tests/test_reference_golden.py: write the SAME files tools/julia_golden.jl writes (names, shapes,
column-major layout, hand-written .npy 1.0 header), but from the ORACLE instead of the reference.

    python tools/mock_reference_golden.py OUTDIR

The result pins nothing (oracle against itself); it only proves that the consuming test reads the files the Julia script
will produce the way it intends to (orientation of every block, node order, keyword conventions), so that a failure on
real reference vectors means a numerical difference and not a plumbing mistake.  Used by
tests/test_reference_golden.py::test_harness_on_mock_vectors.  TEST INFRASTRUCTURE: imports oracle/.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def writenpy(out, name, A):
    """byte-for-byte the writer of tools/julia_golden.jl (format 1.0, fortran_order, 64-byte aligned header)"""
    A = np.asarray(A, dtype=np.float64)
    shape = f"({A.shape[0]},)" if A.ndim == 1 else "(" + ", ".join(str(s) for s in A.shape) + ")"
    hdr = f"{{'descr': '<f8', 'fortran_order': True, 'shape': {shape}, }}"
    pad = (64 - (10 + len(hdr) + 1) % 64) % 64
    hdr = hdr + " " * pad + "\n"
    with open(os.path.join(out, name + ".npy"), "wb") as f:
        f.write(bytes([0x93, 0x4E, 0x55, 0x4D, 0x50, 0x59, 0x01, 0x00]))
        f.write(np.uint16(len(hdr)).astype("<u2").tobytes())
        f.write(hdr.encode("ascii"))
        f.write(np.asfortranarray(A).tobytes(order="F"))


def main(out):
    import clearsky_b200 as cs
    from helpers import c1_problem
    from oracle import oracle as orc
    orc.build()
    os.makedirs(out, exist_ok=True)
    DATA = os.path.join(ROOT, "tests", "data")
    W = lambda n, a: writenpy(out, n, a)
    # 1. faddeyeva
    xs, ys = [], []
    offs = [-1e-1, -1e-3, -1e-6, -1e-9, -1e-12, 0.0, 1e-12, 1e-9, 1e-6, 1e-3, 1e-1]
    for s in (2.5, 3.5, 28.5, 30.0, 62.0, 107.0, 160.0, 256.0, 1.6e4, 3.8e4):
        for f in offs:
            for φ in np.linspace(0.0, np.pi / 2, 33):
                r = np.sqrt(s * (1 + f))
                xs.append(r * np.cos(φ)); ys.append(r * np.sin(φ))
    for y2 in (6e-14, 1e-13, 0.026, 0.072):
        for f in offs:
            for x in np.concatenate([[0.0], 10.0 ** np.linspace(-3, 5, 49)]):
                xs.append(x); ys.append(np.sqrt(y2 * (1 + f)))
    for x in np.concatenate([[0.0], 10.0 ** np.linspace(-6, 5, 56)]):
        for y in 10.0 ** np.linspace(-30, 5, 71):
            xs.append(x); ys.append(y)
    xs, ys = np.array(xs), np.array(ys)
    W("fad_x", xs); W("fad_y", ys); W("fad_w", orc.faddeyeva985(xs, ys))
    # 2. C1
    co2 = cs.SpectralLines.from_file(os.path.join(DATA, "CO2.par.gz"))
    ν, P, Γ = c1_problem(cs)
    T = Γ(P)
    Cc = 400e-6
    W("c1_nu", ν); W("c1_P", P); W("c1_T", T)
    niso, ncheb, cheb, has = co2.cheb_table()
    p = lambda row: np.ascontiguousarray(row).ctypes.data_as(C.POINTER(C.c_double))
    W("c1_line_S", [orc.scalar("orc_scaleintensity", co2.S[j], co2.ν[j], co2.Epp[j], 250.0, C.c_int(int(ncheb[co2.I[j] - 1])),
                               p(cheb[co2.I[j] - 1])) for j in range(co2.N)])
    W("c1_line_alpha", [orc.scalar("orc_alpha_doppler", co2.ν[j], co2.μ[j], 250.0) for j in range(co2.N)])
    W("c1_line_gamma", [orc.scalar("orc_gamma_lorentz", co2.γa[j], co2.γs[j], co2.na[j], 250.0, 5e4, Cc * 5e4) for j in range(co2.N)])
    for name, sid, cut, step, pure in (("voigt", orc.VOIGT, 25.0, 5, False), ("lorentz", orc.LORENTZ, 25.0, 5, False),
                                       ("doppler", orc.DOPPLER, 25.0, 5, False), ("phco2", orc.PHCO2, 500.0, 10, True)):
        Pl = P[::step]
        W(f"c1_sigma_{name}", orc.xsec(sid, co2, ν, T[::step], Pl, Pl if pure else Cc * Pl, cut).T)       # [nν, nlev]
    νq = np.linspace(667.0, 668.0, 2001)
    W("c1q_nu", νq)
    W("c1q_sigma_voigt", orc.xsec(orc.VOIGT, co2, νq, T[::5], P[::5], Cc * P[::5], 25.0).T)
    Ω = cs.AtmosphericDomain((140, 300), 6, (5, 1.1e5), 8)
    block, _ = orc.bake(orc.VOIGT, co2, ν, Ω.T, Ω.P, np.full((Ω.nP, Ω.nT), Cc), 25.0)
    A = orc.table_fit(block)
    W("c1_tab_Tnodes", Ω.T); W("c1_tab_Pnodes", Ω.P)
    nodes = np.stack([orc.gas_nodes(A, Ω.T, Ω.P, Ω.T, np.full(Ω.nT, Ω.P[j]), np.ones(Ω.nT)) for j in range(Ω.nP)])
    W("c1_tab_nodes", np.transpose(nodes, (2, 1, 0)))                                                    # [nν, nT, nP]
    W("c1_tab_levels", orc.gas_nodes(A, Ω.T, Ω.P, T, P, np.ones(len(P))).T)
    Ω = cs.AtmosphericDomain((140, 300), 12, (5, 1.1e5), 24)
    block, _ = orc.bake(orc.VOIGT, co2, ν, Ω.T, Ω.P, np.full((Ω.nP, Ω.nT), Cc), 25.0, nthreads=0)
    A = orc.table_fit(block)
    L = len(P) - 1

    def solve(ns, nl, fS, fa, θ):
        m, Wt = cs.streamnodes(ns)
        x, w = cs.lobattonodes(nl)
        Pn = P[:-1, None] + np.diff(P)[:, None] * x[None, :]
        nodesP = np.concatenate([Pn[:, :-1].ravel(), P[-1:]])
        σn = orc.gas_nodes(A, Ω.T, Ω.P, Γ(nodesP), nodesP, np.full(len(nodesP), Cc))
        return orc.fluxes(ν, P, nl, w, np.full((L, nl), 0.029), T, σn, 9.8, fS, fa, θ, ns, m, Wt), σn, w

    f, σn, _ = solve(5, 2, None, None, 0.841)
    W("c1_tab_Fup", f["Fup"]); W("c1_tab_Fdn", f["Fdn"])
    W("c1_tab12_levels", (σn / Cc).T)
    W("c1_tab_Mup", f["Mup"].T); W("c1_tab_Mdn", f["Mdn"].T)                                             # [np, nν]
    f, _, _ = solve(4, 3, np.full(len(ν), 1e-3), np.full(len(ν), 0.3), 0.5)
    W("c1_tab_Fup_sun", f["Fup"]); W("c1_tab_Fdn_sun", f["Fdn"])
    x, w = cs.lobattonodes(4)
    Pn = P[:-1, None] + np.diff(P)[:, None] * x[None, :]
    nodesP = np.concatenate([Pn[:, :-1].ravel(), P[-1:]])
    σn = orc.gas_nodes(A, Ω.T, Ω.P, Γ(nodesP), nodesP, np.full(len(nodesP), Cc))
    W("c1_tab_depth", orc.opticaldepth(P, 4, w, np.full((L, 4), 0.029), σn, 9.8, 0.0))
    νc = np.linspace(1.0, 3000.0, 1500)
    W("cia_nu", νc)
    one = np.ones(len(P))
    for tag, ex in (("ex", True), ("noex", False)):
        x = cs.CIATables(os.path.join(DATA, "CO2-CO2_2018.cia.gz"), extrapolate=ex)
        W("cia_sigma_" + tag, orc.cia_nodes(x, νc, T, P, one, one).T)
    # 3. C2 slice
    i0, n = 149250, 1500
    ν2 = 0.01 * np.arange(i0 + 1, i0 + n + 1)
    P2 = cs.pressuregrid(10.0, 1e5, 101)
    T2 = Γ(P2)
    W("c2_nu", ν2); W("c2_P", P2); W("c2_T", T2)
    for name, Cg in (("CO2", 400e-6), ("H2O", 1e-3)):
        sl = cs.SpectralLines.from_file(os.path.join(ROOT, "tests", "golden", "ref_inputs", f"c2slice_{name}.par.gz"))
        W("c2_lines_nu_" + name, sl.ν); W("c2_lines_S_" + name, sl.S)
        W("c2_sigma_voigt_" + name, orc.xsec(orc.VOIGT, sl, ν2, T2[::10], P2[::10], Cg * P2[::10], 25.0, nthreads=0).T)
    with open(os.path.join(out, "MANIFEST.txt"), "w") as fh:
        fh.write("MOCK vectors written by tools/mock_reference_golden.py from the oracle -- these pin nothing\n")


if __name__ == "__main__":
    main(sys.argv[1])
