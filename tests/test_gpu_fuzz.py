"""Randomised parity sweep (fixed seeds): small problems with random line densities, grids (uniform at a random
resolution, irregular, clustered with gaps), cut-offs, pressures and temperatures, every shape, in both far-field modes,
each against the CPU oracle at the 1e-9 bar.  The cases are built to land on the class borders the kernels special-case:
cut-off windows narrower than a tile, tiles wider than the cut-off, near-centre ranges that reach the window edges,
empty windows, single-line windows, lines exactly on grid points."""
import numpy as np
import pytest

from conftest import relerr
from helpers import synthetic_lines

pytestmark = pytest.mark.gpu

SHAPES = [("doppler", 0), ("lorentz", 1), ("voigt", 2), ("PHCO2", 3)]


def _grid(rng, kind, n):
    lo = rng.uniform(5.0, 900.0)
    if kind == 0:                       # uniform, resolution from 1e-4 to 3 cm^-1
        return lo + 10 ** rng.uniform(-4, 0.5) * np.arange(n)
    if kind == 1:                       # irregular
        return np.unique(lo + rng.uniform(0, rng.uniform(1.0, 400.0), n))
    parts, x = [], lo                   # clusters separated by gaps
    while sum(len(p) for p in parts) < n:
        m = int(rng.integers(1, 200))
        parts.append(x + 10 ** rng.uniform(-3, -0.5) * np.arange(m))
        x = parts[-1][-1] + rng.uniform(0.5, 120.0)
    return np.unique(np.concatenate(parts))[:n]


@pytest.mark.parametrize("seed", range(48))
def test_random_cases_both_modes(cs, orc, seed):
    rng = np.random.default_rng(1000 + seed)
    ctx = cs.default_context()
    prev = ctx.get_farfield()
    try:
        for case in range(3):
            nl = int(rng.choice([1, 7, 300, 3000]))
            sl = synthetic_lines(cs, nl, seed=seed * 10 + case, νmax=float(rng.choice([200.0, 1500.0])))
            ν = _grid(rng, int(rng.integers(0, 3)), int(rng.choice([1, 33, 128, 129, 1000, 2500])))
            if case == 2 and nl > 1:    # put some grid points exactly on line centres
                ν = np.unique(np.concatenate([ν, sl.ν[:: max(1, nl // 17)]]))
            nlev = int(rng.integers(1, 4))
            T = rng.uniform(100.0, 400.0, nlev)
            P = 10 ** rng.uniform(0.0, 6.5, nlev)
            Pp = P * rng.uniform(0.0, 1.0, nlev)
            for name, sid in SHAPES:
                cut = float(rng.choice([0.03, 1.0, 25.0, 140.0, 500.0])) if name == "PHCO2" else float(rng.choice([0.03, 1.0, 25.0, 300.0]))
                ref = orc.xsec(sid, sl, ν, T, P, Pp, cut, nthreads=0)
                for mode in ("direct", "expansion"):
                    ctx.set_farfield(mode)
                    got = cs.xsec(name, ν, sl, T, P, Pp, cut)
                    assert relerr(got, ref, 1e-290) < 1e-9, (seed, case, name, mode, cut, nl, len(ν))
    finally:
        ctx.set_farfield(prev)
