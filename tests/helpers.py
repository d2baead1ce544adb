"""shared builders of test problems (inputs only; no oracle or product imports here)"""
import numpy as np


def synthetic_lines(cs, n, seed=20261018, M=2, νmax=3000.0, γs_rng=(0.06, 0.13)):
    """SURVEY.md section 8(d) config-2 style synthetic HITRAN-format lines, rounded to the .par field widths"""
    rng = np.random.default_rng(np.random.PCG64(seed))
    ν = np.sort(np.round(rng.uniform(0.0, νmax, n), 6))
    ν = np.maximum(ν, 1e-6)
    S = np.array([float(f"{x:.3E}") for x in 10.0 ** rng.uniform(-30, -19, n)])
    γa = np.round(rng.uniform(0.05, 0.10, n), 4)
    γs = np.round(rng.uniform(*γs_rng, n), 3)
    Epp = np.round(rng.uniform(0, 3000, n), 4)
    na = np.round(rng.uniform(0.5, 0.8, n), 2)
    mp = cs.MOLPARAM[M]
    I = np.ones(n, dtype=np.int16)
    return cs.SpectralLines(mp.name, mp.formula, n, M, I, np.full(n, mp.mu[0]), np.full(n, mp.A[0]),
                            ν, S, γa, γs, Epp, na)


def c1_problem(cs, nν=1000, nlayer=20):
    """BASELINE.json configs[0]: ν_i = 1 + 2.5 i, dry adiabat 1 bar / 288 K, 20 layers"""
    ν = 1.0 + 2.5 * np.arange(nν)
    P = cs.pressuregrid(10.0, 1e5, nlayer + 1)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    return ν, P, Γ
