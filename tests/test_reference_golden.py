"""The oracle against golden vectors of the UNMODIFIED reference (tools/julia_golden.jl -> tests/golden/ref/*.npy).

There is no Julia in the build image, so the vectors can only be produced elsewhere:

    julia --project=/path/to/ClearSky.jl tools/julia_golden.jl        # writes tests/golden/ref/
    python -m pytest tests/test_reference_golden.py -q                # pins (or refutes) oracle/oracle.c

While tests/golden/ref/ is absent every test here is an expected failure with the reason "PARITY UNPINNED"; with the
files present they are ordinary assertions at the north-star tolerances (1e-9 on cross-sections, 1e-8 on fluxes), and
`test_pin_status` then reports the generating package versions.  CPU only (-m "not gpu").
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import DATA, GOLDEN, ROOT, relerr

REF = os.environ.get("CS_REF_GOLDEN_DIR") or os.path.join(GOLDEN, "ref")          # the override is for the harness self-test
PINNED = os.path.exists(os.path.join(REF, "MANIFEST.txt"))
UNPINNED = "PARITY UNPINNED: tests/golden/ref/ is absent -- run tools/julia_golden.jl against the unmodified reference"
TOL_SIGMA, TOL_FLUX = 1e-9, 1e-8

needs_ref = pytest.mark.xfail(not PINNED, reason=UNPINNED, run=False, strict=True)


def ref(name):
    return np.load(os.path.join(REF, name + ".npy"))


def test_pin_status():
    """always runs: says which state the oracle is in, and that the recipe and its inputs are in the tree"""
    assert os.path.exists(os.path.join(ROOT, "tools", "julia_golden.jl"))
    for f in ("c2slice_CO2.par.gz", "c2slice_H2O.par.gz", "c2slice.json"):
        assert os.path.exists(os.path.join(GOLDEN, "ref_inputs", f)), f
    if PINNED:
        print("\nparity pinned by:", open(os.path.join(REF, "MANIFEST.txt")).read())
    else:
        print("\n" + UNPINNED)


def test_recipe_inputs_match_generator(cs):
    """the committed .par slices are exactly what bench.py's generator makes for configs[1] (so a Julia run on them is a
    run on the benchmark's lines), and survive writepar -> readpar bit for bit"""
    import json

    import bench
    meta = json.load(open(os.path.join(GOLDEN, "ref_inputs", "c2slice.json")))
    wl = bench.make_workload(cs, "c2")
    i0, n = meta["first_index"], meta["n_nu"]
    νs = wl["ν"][i0:i0 + n]
    assert νs[0] == meta["nu_first"] and np.array_equal(νs, 0.01 * np.arange(i0 + 1, i0 + n + 1))
    for (sl, Cg), name in zip(wl["gases"], ("CO2", "H2O")):
        sub = cs.SpectralLines.from_file(os.path.join(GOLDEN, "ref_inputs", f"c2slice_{name}.par.gz"))
        assert sub.N == meta["gases"][name]["lines"] and Cg == meta["gases"][name]["C"]
        k = (sl.ν >= sub.ν[0]) & (sl.ν <= sub.ν[-1])
        for a in ("ν", "S", "γa", "γs", "Epp", "na", "I"):
            assert np.array_equal(getattr(sub, a), getattr(sl, a)[k]), (name, a)
        # every line the slice can see (strict prefilter, line_shapes.jl:21) is in the file
        seen = (sl.ν > νs[0] - meta["cut"]) & (sl.ν < νs[-1] + meta["cut"])
        assert np.all(k[seen])


@pytest.mark.skipif(bool(os.environ.get("CS_REF_GOLDEN_DIR")), reason="inner run of the harness self-test")
def test_harness_on_mock_vectors(tmp_path):
    """the consuming tests below, run on files laid out exactly like the Julia script's but computed by the oracle
    (tools/mock_reference_golden.py): proves orientation / node order / keyword plumbing, pins nothing"""
    import subprocess
    import sys
    out = str(tmp_path / "mock")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "mock_reference_golden.py"), out])
    env = dict(os.environ, CS_REF_GOLDEN_DIR=out)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and "xfailed" not in r.stdout and "passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@needs_ref
def test_faddeyeva_region_map(orc):
    """Faddeyeva985.faddeyeva(x, y) on points straddling every border of BOTH candidate region maps
    (oracle.c, orc_faddeyeva985 <-> line_shapes.jl:375).  The oracle's default map must be the package's; when the other map is
    the one that matches, the message says so (orc_set_w985_map(0) / rebuild the CUDA library with -DCS_W985_MAP=0)."""
    x, y, w = ref("fad_x"), ref("fad_y"), ref("fad_w")
    errs = {}
    try:
        for m in (1, 0):
            orc.set_w985_map(m)
            got = orc.faddeyeva985(x, y)
            e = np.abs(got - w) / np.maximum(np.abs(w), 1e-300)   # y = 0 in a continued-fraction region gives exactly 0 on both sides
            worst = int(np.argmax(e))
            errs[m] = (float(e[worst]), float(x[worst]), float(y[worst]))
    finally:
        orc.set_w985_map(1)
    default = orc.get_w985_map()
    assert errs[default][0] < TOL_SIGMA, (
        f"the reference's faddeyeva does not follow the oracle's default region map {default}: max rel. error "
        f"{errs[default][0]:.2e} at x = {errs[default][1]:.6g}, y = {errs[default][2]:.6g}; the other map gives {errs[1 - default][0]:.2e}"
        + (" -> it is the package's: switch the default (orc_w985_map in oracle.c, CS_W985_MAP in cs_internal.cuh)"
           if errs[1 - default][0] < TOL_SIGMA else " -> neither map is the package's"))


@needs_ref
def test_c1_inputs_and_line_parameters(orc, cs, co2):
    from helpers import c1_problem
    ν, P, Γ = c1_problem(cs)
    assert np.array_equal(ν, ref("c1_nu"))
    assert relerr(P, ref("c1_P")) < 1e-14 and relerr(Γ(P), ref("c1_T")) < 1e-12
    niso, ncheb, cheb, has = co2.cheb_table()
    p = lambda row: np.ascontiguousarray(row).ctypes.data_as(C.POINTER(C.c_double))
    S = np.array([orc.scalar("orc_scaleintensity", co2.S[j], co2.ν[j], co2.Epp[j], 250.0, C.c_int(int(ncheb[co2.I[j] - 1])),
                             p(cheb[co2.I[j] - 1])) for j in range(co2.N)])
    α = np.array([orc.scalar("orc_alpha_doppler", co2.ν[j], co2.μ[j], 250.0) for j in range(co2.N)])
    γ = np.array([orc.scalar("orc_gamma_lorentz", co2.γa[j], co2.γs[j], co2.na[j], 250.0, 5e4, 400e-6 * 5e4) for j in range(co2.N)])
    assert relerr(S, ref("c1_line_S")) < 1e-13
    assert relerr(α, ref("c1_line_alpha")) < 1e-14
    assert relerr(γ, ref("c1_line_gamma")) < 1e-14


@needs_ref
@pytest.mark.parametrize("name,cut,step,pure", [("voigt", 25.0, 5, False), ("lorentz", 25.0, 5, False),
                                                ("doppler", 25.0, 5, False), ("phco2", 500.0, 10, True)])
def test_c1_shapes(orc, co2, name, cut, step, pure):
    """voigt!/lorentz!/doppler!/PHCO2! on configs[0] (line_shapes.jl:200, 313, 412, 527)"""
    ν, P, T = ref("c1_nu"), ref("c1_P")[::step], ref("c1_T")[::step]
    want = ref(f"c1_sigma_{name}").T                      # Julia [nν, nlev] -> [nlev, nν]
    got = orc.xsec(getattr(orc, name.upper()), co2, ν, T, P, P if pure else 400e-6 * P, cut)
    assert relerr(got, want, 1e-290) < TOL_SIGMA


@needs_ref
def test_c1_q_branch_fine_grid(orc, co2):
    """0.0005 cm^-1 grid across the 15 micron Q branch: near-centre Faddeyeva branches at 10 Pa ... 1 bar"""
    got = orc.xsec(orc.VOIGT, co2, ref("c1q_nu"), ref("c1_T")[::5], ref("c1_P")[::5], 400e-6 * ref("c1_P")[::5], 25.0)
    assert relerr(got, ref("c1q_sigma_voigt").T, 1e-290) < TOL_SIGMA


@needs_ref
def test_c1_bake_and_opacity_table(orc, cs, co2):
    """bake on a 6 x 8 AtmosphericDomain and OpacityTable evaluation at nodes and off nodes (gases.jl:26-145)"""
    ν, P, T = ref("c1_nu"), ref("c1_P"), ref("c1_T")
    Ω = cs.AtmosphericDomain((140, 300), 6, (5, 1.1e5), 8)
    assert relerr(Ω.T, ref("c1_tab_Tnodes")) < 1e-14 and relerr(Ω.P, ref("c1_tab_Pnodes")) < 1e-13
    block, _ = orc.bake(orc.VOIGT, co2, ν, Ω.T, Ω.P, np.full((Ω.nP, Ω.nT), 400e-6), 25.0)
    want = np.transpose(ref("c1_tab_nodes"), (2, 1, 0))     # [nν, nT, nP] -> [nP, nT, nν]
    A = orc.table_fit(block)
    nodes = np.empty_like(block)
    for j in range(Ω.nP):
        nodes[j] = orc.gas_nodes(A, Ω.T, Ω.P, Ω.T, np.full(Ω.nT, Ω.P[j]), np.ones(Ω.nT))
    assert relerr(nodes, want, 1e-290) < TOL_SIGMA
    lev = orc.gas_nodes(A, Ω.T, Ω.P, T, P, np.ones(len(P)))
    assert relerr(lev, ref("c1_tab_levels").T, 1e-290) < TOL_SIGMA


def _c1_table_sigma(orc, cs, co2, ν, T, P):
    Ω = cs.AtmosphericDomain((140, 300), 12, (5, 1.1e5), 24)
    block, _ = orc.bake(orc.VOIGT, co2, ν, Ω.T, Ω.P, np.full((Ω.nP, Ω.nT), 400e-6), 25.0, nthreads=0)
    return orc.table_fit(block), Ω


@needs_ref
def test_c1_fluxes(orc, cs, co2):
    """fluxes / monochromaticfluxes / opticaldepth on configs[0] through its 12 x 24 table (fluxes.jl:68-97, 281-340)"""
    ν, P, T = ref("c1_nu"), ref("c1_P"), ref("c1_T")
    A, Ω = _c1_table_sigma(orc, cs, co2, ν, T, P)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    L = len(P) - 1
    for tag, ns, nl, fS, fa, θ in (("", 5, 2, None, None, 0.841), ("_sun", 4, 3, 1e-3, 0.3, 0.5)):
        m, W = cs.streamnodes(ns)
        x, w = cs.lobattonodes(nl)
        Pn = (P[:-1, None] + np.diff(P)[:, None] * x[None, :])          # discretized.jl:19-27
        nodesP = np.concatenate([Pn[:, :-1].ravel(), P[-1:]])
        σn = orc.gas_nodes(A, Ω.T, Ω.P, Γ(nodesP), nodesP, np.full(len(nodesP), 400e-6))
        f = orc.fluxes(ν, P, nl, w, np.full((L, nl), 0.029), Γ(P), σn, 9.8, None if fS is None else np.full(len(ν), fS),
                       None if fa is None else np.full(len(ν), fa), θ, ns, m, W)
        assert relerr(f["Fup"], ref("c1_tab_Fup" + tag)) < TOL_FLUX
        assert relerr(f["Fdn"][1:], ref("c1_tab_Fdn" + tag)[1:]) < TOL_FLUX
        if tag == "":
            assert relerr(σn, 400e-6 * ref("c1_tab12_levels").T, 1e-290) < TOL_SIGMA
            assert relerr(f["Mup"], ref("c1_tab_Mup").T, 1e-300) < TOL_FLUX
            assert relerr(f["Mdn"][:, 1:], ref("c1_tab_Mdn").T[:, 1:], 1e-300) < TOL_FLUX
    x, w = cs.lobattonodes(4)
    Pn = (P[:-1, None] + np.diff(P)[:, None] * x[None, :])
    nodesP = np.concatenate([Pn[:, :-1].ravel(), P[-1:]])
    σn = orc.gas_nodes(A, Ω.T, Ω.P, Γ(nodesP), nodesP, np.full(len(nodesP), 400e-6))
    τ = orc.opticaldepth(P, 4, w, np.full((L, 4), 0.029), σn, 9.8, 0.0)
    assert relerr(τ, ref("c1_tab_depth")) < TOL_SIGMA


@needs_ref
def test_cia(orc, cs):
    """cia(ν, x, T, Pa, P1, P2) on the reference's CO2-CO2 file, both extrapolate settings (cia.jl:251-303)"""
    ν, P, T = ref("cia_nu"), ref("c1_P"), ref("c1_T")
    one = np.ones(len(P))
    for tag, ex in (("ex", True), ("noex", False)):
        x = cs.CIATables(os.path.join(DATA, "CO2-CO2_2018.cia.gz"), extrapolate=ex)
        got = orc.cia_nodes(x, ν, T, P, one, one)
        assert relerr(got, ref("cia_sigma_" + tag).T, 1e-300) < TOL_SIGMA


@needs_ref
def test_c2_slice(orc, cs):
    """configs[1]: the two synthetic line lists through the reference's readpar + voigt! on the 1500-point slice"""
    ν, P, T = ref("c2_nu"), ref("c2_P")[::10], ref("c2_T")[::10]
    for name, Cg in (("CO2", 400e-6), ("H2O", 1e-3)):
        sl = cs.SpectralLines.from_file(os.path.join(GOLDEN, "ref_inputs", f"c2slice_{name}.par.gz"))
        assert np.array_equal(sl.ν, ref("c2_lines_nu_" + name)) and np.array_equal(sl.S, ref("c2_lines_S_" + name))
        got = orc.xsec(orc.VOIGT, sl, ν, T, P, Cg * P, 25.0, nthreads=0)
        assert relerr(got, ref("c2_sigma_voigt_" + name).T, 1e-290) < TOL_SIGMA
