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


# ------------------------------------------------------------------------------------------------
# GPU ingestion: same result as readpar, parse loop on the device
def parse_records_b200(buf, ctx=None):
    """parse the numeric columns of fixed-width .par records held in `buf` (bytes) with cs_par_parse.
    Returns (dict of arrays in file order, flags) -- flags[i] != 0 marks a record with a malformed field."""
    import ctypes as C

    from . import _lib
    from ._lib import check, lib, ptr
    ctx = ctx or _lib.default_context()
    reclen, nrec = _record_geometry(buf)
    M, I = np.empty(nrec, np.int16), np.empty(nrec, np.int16)
    cols = {k: np.empty(nrec) for k in ("ν", "S", "A", "γa", "γs", "Epp", "na", "δa")}
    flags = np.empty(nrec, np.uint8)
    i16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int16))
    check(lib().cs_par_parse(ctx.h, len(buf), buf, reclen, nrec, i16(M), i16(I), *[ptr(cols[k]) for k in
                             ("ν", "S", "A", "γa", "γs", "Epp", "na", "δa")], flags.ctypes.data_as(C.POINTER(C.c_uint8))))
    par = {"M": M, "I": I}
    par.update(cols)
    return par, flags, reclen


def _record_geometry(buf):
    """(reclen, nrec) of a buffer of fixed-width records (a trailing blank line is not a record)"""
    nl = buf.find(b"\n")
    reclen = nl + 1 if nl >= 0 else len(buf)
    assert reclen >= 161 or (nl < 0 and reclen >= 160), "expected 160-column HITRAN records"
    nrec = (len(buf) + reclen - 1) // reclen if len(buf) % reclen else len(buf) // reclen
    if len(buf) % reclen and len(buf) - (nrec - 1) * reclen < 67:
        nrec -= 1
    return reclen, nrec


def readpar_b200(filename, νmin=0, νmax=np.inf, Scut=0, I=(), maxlines=-1, ctx=None, index=False, timing=None):
    """readpar (par.jl:91-193) in one device call (cs_par_read): parse loop (:127-152), filters (:154-170), the
    maxlines truncation (:178-185) and the final sort by ν (:187-191) all run on the GPU, so only the surviving records
    cross PCIe, already in output order.  "I" is returned as the isotopologue character like readpar does; every
    numeric column is bit-identical to the host parser.  index=True adds "record" = 0-based file record of each row
    (to gather the quantum-number string columns)."""
    import ctypes as C

    from . import _lib
    from ._lib import check, lib, ptr
    base = filename[:-3] if filename.endswith(".gz") else filename
    assert base.endswith(".par"), "expected file with .par extension, downloaded from https://hitran.org/lbl/"
    op = gzip.open if filename.endswith(".gz") else open
    import time
    t0 = time.perf_counter()
    with op(filename, "rb") as f:
        buf = f.read()
    t1 = time.perf_counter()
    ctx = ctx or _lib.default_context()
    reclen, nrec = _record_geometry(buf)
    Ilist = np.array([ISOINDEX[i] if isinstance(i, str) else int(i) for i in I], dtype=np.int16)
    M, Iv = np.empty(nrec, np.int16), np.empty(nrec, np.int16)
    keys = ("ν", "S", "A", "γa", "γs", "Epp", "na", "δa")
    cols = {k: np.empty(nrec) for k in keys}
    idx = np.empty(nrec, np.int64)
    nout, nbad = C.c_int64(0), C.c_int64(0)
    i16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int16))
    rc = lib().cs_par_read(ctx.h, len(buf), buf, reclen, nrec, float(νmin), float(νmax), float(Scut), len(Ilist),
                           i16(Ilist) if len(Ilist) else None, int(maxlines), i16(M), i16(Iv), *[ptr(cols[k]) for k in keys],
                           idx.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nout), C.byref(nbad))
    if timing is not None:      # where the wall clock of a call goes: file read, the device call (H2D + kernels + D2H)
        timing.update(read_s=t1 - t0, call_s=time.perf_counter() - t1, bytes=len(buf), records=nrec)
    if rc != 0 and b"filtered to nothing" in lib().cs_last_error():
        raise AssertionError("par information has been filtered to nothing!")          # par.jl:172 is an @assert
    check(rc)
    if nbad.value:
        raise ValueError(f"{nbad.value} malformed record(s) in {filename}")
    n = nout.value
    lut = np.array([" "] + [k for k, _ in sorted(ISOINDEX.items(), key=lambda kv: kv[1])], dtype="U1")     # number -> character
    par = {"M": M[:n].copy(), "I": lut[Iv[:n]]}
    for k in keys:
        par[k] = cols[k][:n].copy()
    if index:
        par["record"] = idx[:n].copy()
    return par


def _fortran_f(x, w, d):
    """Fw.d the way HITRAN files print it: drop the leading zero when the field is too narrow"""
    s = f"{x:{w}.{d}f}"
    if len(s) > w:
        s = s.replace("0.", ".", 1) if s.lstrip("-").startswith("0.") else s
    assert len(s) <= w, (x, w, d, s)
    return s.rjust(w)


def writepar(filename, sl, A=None, δa=None):
    """write a SpectralLines object as 160-column HITRAN records (column map of par.jl:131-149), so that the same
    synthetic line list can be fed to the reference's readpar.  Quantum-number columns are blank."""
    inv = {v: k for k, v in ISOINDEX.items()}
    op = gzip.open if filename.endswith(".gz") else open
    with op(filename, "wt") as f:
        for j in range(sl.N):
            rec = (f"{sl.M:2d}{inv[int(sl.I[j])]}{sl.ν[j]:12.6f}{sl.S[j]:10.3E}{(A[j] if A is not None else 0.0):10.3E}"
                   f"{_fortran_f(sl.γa[j], 5, 4)}{_fortran_f(sl.γs[j], 5, 3)}{sl.Epp[j]:10.4f}{sl.na[j]:4.2f}"
                   f"{_fortran_f(δa[j] if δa is not None else 0.0, 8, 6)}")
            assert len(rec) == 67, (len(rec), rec)
            f.write(rec.ljust(160) + "\n")
