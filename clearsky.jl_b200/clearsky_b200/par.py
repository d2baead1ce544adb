"""HITRAN .par reader and SpectralLines (host side; stays host side in the drop-in too).

Reference: src/hitran/par.jl:6-13 (ISOINDEX), :91-193 (readpar), :224-286 (SpectralLines).
"""
import gzip
from dataclasses import dataclass

import numpy as np

from .molparam import MOLPARAM

# par.jl:6-13
ISOINDEX = {ch: i + 1 for i, ch in enumerate("1234567890ABCDEFGHIJKLMNOPQRSTUVWXYZ")}

# (key, start, stop) 0-based half-open slices of the 160-column record -- par.jl:131-149
_COLS = (("M", 0, 2), ("I", 2, 3), ("ν", 3, 15), ("S", 15, 25), ("A", 25, 35), ("γa", 35, 40),
         ("γs", 40, 45), ("Epp", 45, 55), ("na", 55, 59), ("δa", 59, 67))
_STR_COLS = (("Vp", 67, 82), ("Vpp", 82, 97), ("Qp", 97, 112), ("Qpp", 112, 127),
             ("Ierr", 127, 133), ("Iref", 133, 145), ("*", 145, 146), ("gp", 146, 153),
             ("gpp", 153, 160))


def _readlines(filename):
    op = gzip.open if filename.endswith(".gz") else open
    with op(filename, "rt") as f:
        return [ln.rstrip("\n").rstrip("\r") for ln in f if ln.strip("\r\n") != ""]


def readpar(filename, νmin=0, νmax=np.inf, Scut=0, I=(), maxlines=-1, progress=False, strings=False):
    """readpar(filename; νmin=0, νmax=Inf, Scut=0, I=[], maxlines=-1) -- par.jl:91-193.

    Returns a dict of numpy arrays keyed like the reference ("M","I","ν","S","A","γa","γs","Epp","na","δa").
    `.par.gz` is accepted in addition to `.par` (fixtures are stored compressed).  The quantum-number
    string columns are parsed only with strings=True (they are never used on the hot path).
    """
    base = filename[:-3] if filename.endswith(".gz") else filename
    assert base.endswith(".par"), "expected file with .par extension, downloaded from https://hitran.org/lbl/"
    lines = _readlines(filename)
    N = len(lines)
    par = {}
    # fixed 160-column records: slice the columns of all records at once (numpy), parse per column
    if N > 0 and all(len(ln) >= 160 for ln in lines[:1000:7]) and min(map(len, lines)) >= 67:
        width = 67
        raw = np.frombuffer("".join(ln[:width] for ln in lines).encode("ascii"), dtype=np.uint8).reshape(N, width)

        def col(a, b, dtype):
            return np.ascontiguousarray(raw[:, a:b]).view(f"S{b - a}").ravel().astype(dtype)

        par["M"] = col(0, 2, np.int64).astype(np.int16)
        par["I"] = col(2, 3, "U1")
        for key, a, b in _COLS[2:]:
            par[key] = col(a, b, np.float64)
    else:
        par["M"] = np.array([int(ln[0:2]) for ln in lines], dtype=np.int16)
        par["I"] = np.array([ln[2] for ln in lines], dtype="U1")
        for key, a, b in _COLS[2:]:
            par[key] = np.array([float(ln[a:b]) for ln in lines], dtype=np.float64)
    if strings:
        for key, a, b in _STR_COLS:
            par[key] = np.array([ln[a:b] for ln in lines], dtype=object)
    # filtering -- par.jl:154-169
    mask = np.ones(N, dtype=bool)
    mask &= par["ν"] >= νmin
    mask &= par["ν"] <= νmax
    mask &= par["S"] >= Scut
    if len(I) > 0:
        Iset = set(I)
        for j in range(N):
            ch = par["I"][j]
            if (ch not in Iset) and (ISOINDEX[ch] not in Iset):
                mask[j] = False
    assert mask.any(), "par information has been filtered to nothing!"
    for key in par:
        par[key] = par[key][mask]
    # strongest lines -- par.jl:177-186 (compares against the PRE-filter count N, like the reference)
    if maxlines > 0 and N > maxlines:
        idx = np.argsort(par["S"], kind="stable")[::-1][:maxlines]
        for key in par:
            par[key] = par[key][idx]
    idx = np.argsort(par["ν"], kind="stable")
    for key in par:
        par[key] = par[key][idx]
    return par


@dataclass
class SpectralLines:
    """par.jl:224-251: SoA of one molecule's lines, sorted by wavenumber."""
    name: str
    formula: str
    N: int
    M: int
    I: np.ndarray      # int16 local isotopologue numbers (1-based)
    μ: np.ndarray
    A: np.ndarray
    ν: np.ndarray
    S: np.ndarray
    γa: np.ndarray
    γs: np.ndarray
    Epp: np.ndarray
    na: np.ndarray

    @classmethod
    def from_par(cls, par):
        """SpectralLines(par::Dict) -- par.jl:253-284"""
        N = len(par["ν"])
        assert len(np.unique(par["M"])) == 1, "SpectralLines objects must contain only one molecule's lines"
        M = int(par["M"][0])
        mp = MOLPARAM[M]
        I = np.array([ISOINDEX[ch] for ch in par["I"]], dtype=np.int16)
        A = np.array([mp.A[i - 1] for i in I], dtype=np.float64)
        μ = np.array([mp.mu[i - 1] for i in I], dtype=np.float64)
        idx = np.argsort(par["ν"], kind="stable")
        f = lambda x: np.ascontiguousarray(np.asarray(x)[idx])
        return cls(mp.name, mp.formula, N, M, f(I), f(μ), f(A), f(par["ν"]).astype(np.float64),
                   f(par["S"]).astype(np.float64), f(par["γa"]).astype(np.float64),
                   f(par["γs"]).astype(np.float64), f(par["Epp"]).astype(np.float64),
                   f(par["na"]).astype(np.float64))

    @classmethod
    def from_file(cls, filename, **kwargs):
        """SpectralLines(filename; kwargs...) -- par.jl:286"""
        return cls.from_par(readpar(filename, **kwargs))

    def cheb_table(self):
        """(niso, ncheb[niso] int32, cheb[niso, MAXCHEB] float64, hascheb[niso] uint8) of MOLPARAM[M]"""
        from .molparam import MAXCHEB
        mp = MOLPARAM[self.M]
        niso = len(mp.A)
        ncheb = np.zeros(niso, dtype=np.int32)
        cheb = np.zeros((niso, MAXCHEB), dtype=np.float64)
        has = np.zeros(niso, dtype=np.uint8)
        for i in range(niso):
            has[i] = 1 if mp.hascheb[i] else 0
            ncheb[i] = mp.ncheb[i]
            assert mp.ncheb[i] <= MAXCHEB
            cheb[i, : mp.ncheb[i]] = mp.cheb[i]
        return niso, ncheb, cheb, has
