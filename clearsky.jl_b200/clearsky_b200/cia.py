"""Collision-induced absorption: readcia, CIATables, CIA.

Reference: src/absorption/collision_induced_absorption.jl -- readcia :39-94, CIATables :145-242,
table evaluation :251-276, cia :295-303, CIA :431-465.  Parsing stays on the host; evaluation of
k(ν,T) and the amagat conversion run on the GPU inside cs_sigma_add_cia.
"""
import ctypes as C
import gzip

import numpy as np

from . import _lib
from ._lib import check, f64, i64ptr, lib, ptr


def readcia(filename):
    """readcia(filename) -- collision_induced_absorption.jl:39-94 (`.cia.gz` accepted as well)"""
    base = filename[:-3] if filename.endswith(".gz") else filename
    assert base.endswith(".cia"), "expected file with .cia extension downloaded from https://hitran.org/cia/"
    op = gzip.open if filename.endswith(".gz") else open
    with op(filename, "rt") as f:
        lines = [ln.rstrip("\n").rstrip("\r") for ln in f]
    while lines and lines[-1] == "":
        lines.pop()
    L = [len(ln) for ln in lines]
    assert max(L) == 100, f"unexpected maximum line length in cia file, expected 100 but got {max(L)}"
    hidx = [i for i, n in enumerate(L) if n == 100]
    hidx.append(len(lines))
    data = []
    for a, b in zip(hidx[:-1], hidx[1:]):
        h = lines[a]
        d = {
            "symbol": h[0:20].strip(), "νmin": float(h[20:30]), "νmax": float(h[30:40]),
            "npts": int(h[40:47]), "T": float(h[47:54]), "maxcia": float(h[54:64]),
            "res": float(h[64:70]), "comments": h[70:97].strip(), "reference": int(h[97:100]),
        }
        rows = [ln.split() for ln in lines[a + 1:b]]
        d["ν"] = np.array([float(r[0]) for r in rows])
        d["k"] = np.array([float(r[1]) for r in rows])
        data.append(d)
    return data


class CIATables:
    """CIATables(data | filename; extrapolate=false, singles=false) -- collision_induced_absorption.jl:161-242"""

    def __init__(self, data, extrapolate=False, singles=False, verbose=False, ctx=None):
        if isinstance(data, str):
            data = readcia(data)
        self.extrapolate, self.singles = bool(extrapolate), bool(singles)
        νranges = sorted(set((d["νmin"], d["νmax"]) for d in data), key=lambda x: x[0])
        self.grids, self.single_tables, self.T = [], [], []
        tiny = np.finfo(np.float64).tiny
        for νmin, νmax in νranges:
            grp = [d for d in data if np.isclose(d["νmin"], νmin, rtol=1.5e-8, atol=0)
                   and np.isclose(d["νmax"], νmax, rtol=1.5e-8, atol=0)]
            if len(grp) == 1:
                ν, k = grp[0]["ν"], grp[0]["k"].copy()
                k[k <= 0.0] = 0.0
                with np.errstate(divide="ignore"):
                    self.single_tables.append((ν, np.log(k)))
                self.T.append(grp[0]["T"])
            else:
                for g in grp[1:]:
                    assert np.isclose(np.sum(grp[0]["ν"] - g["ν"]), 0.0, atol=1e-8), \
                        "wavenumber sample within a wavenumber range appear to be different"
                grp = sorted(grp, key=lambda d: d["T"])
                ν = grp[0]["ν"]
                T = np.array([g["T"] for g in grp])
                k = np.stack([g["k"] for g in grp], axis=0)   # [nT, nν]  (Julia Z[iν, jT], ν fastest)
                k = np.where(k <= 0.0, tiny, k)
                self.grids.append((ν, T, np.log(k)))
        symbols = sorted(set(d["symbol"] for d in data))
        assert len(symbols) == 1
        self.name = symbols[0]
        self.formulae = tuple(self.name.split("-"))
        self._dev = {}

    def flat(self):
        """flattened arrays in the order the C ABI takes them (cs_cia_upload)"""
        g_nnu = np.array([len(g[0]) for g in self.grids], dtype=np.int64)
        g_nT = np.array([len(g[1]) for g in self.grids], dtype=np.int64)
        cat = lambda xs: f64(np.concatenate(xs)) if xs else np.zeros(0)
        g_nu = cat([g[0] for g in self.grids])
        g_T = cat([g[1] for g in self.grids])
        g_lnk = cat([g[2].ravel() for g in self.grids])
        s_n = np.array([len(s[0]) for s in self.single_tables], dtype=np.int64)
        s_nu = cat([s[0] for s in self.single_tables])
        s_lnk = cat([s[1] for s in self.single_tables])
        return g_nnu, g_nT, g_nu, g_T, g_lnk, s_n, s_nu, s_lnk

    def handle(self, ctx=None):
        ctx = ctx or _lib.default_context()
        if id(ctx) not in self._dev:
            g_nnu, g_nT, g_nu, g_T, g_lnk, s_n, s_nu, s_lnk = self.flat()
            h = C.c_void_p()
            z = np.zeros(1)
            zi = np.zeros(1, dtype=np.int64)
            check(lib().cs_cia_upload(
                ctx.h, len(g_nnu), i64ptr(g_nnu if len(g_nnu) else zi), i64ptr(g_nT if len(g_nT) else zi),
                ptr(g_nu if len(g_nu) else z), ptr(g_T if len(g_T) else z), ptr(g_lnk if len(g_lnk) else z),
                len(s_n), i64ptr(s_n if len(s_n) else zi), ptr(s_nu if len(s_nu) else z),
                ptr(s_lnk if len(s_lnk) else z), int(self.extrapolate), int(self.singles), C.byref(h)))
            self._dev[id(ctx)] = (h, ctx)
        return self._dev[id(ctx)][0]

    def cia(self, ν, T, Pa, P1, P2, ctx=None):
        """cia(ν, x::CIATables, T, Pₐ, P₁, P₂): cross-sections [cm²/molecule] for a vector of wavenumbers
        (collision_induced_absorption.jl:318-323), evaluated on the GPU"""
        ν = f64(np.atleast_1d(ν))
        ctx = ctx or _lib.default_context()
        h = C.c_void_p()
        check(lib().cs_sigma_create(ctx.h, len(ν), ptr(ν), 1, C.byref(h)))
        try:
            one = lambda v: f64(np.array([float(v)]))
            check(lib().cs_sigma_add_cia(h, self.handle(ctx), ptr(one(T)), ptr(one(Pa)), ptr(one(P1 / Pa)), ptr(one(P2 / Pa))))
            out = np.empty((1, len(ν)))
            check(lib().cs_sigma_read(h, ptr(out)))
        finally:
            lib().cs_sigma_free(h)
        return out[0]

    def __call__(self, ν, T):
        """(tables::CIATables)(ν, T): absorption coefficient k [cm⁵/molecule²] (collision_induced_absorption.jl:251-276),
        recovered from the cross-section at unit partial pressures"""
        from . import constants as K
        Pa = K.A
        ρ = (Pa / K.A) * (K.T0 / T)
        ρa = 1e-6 * Pa / (K.k * T)
        σ = self.cia(ν, T, Pa, Pa, Pa)
        k = σ * ρa / (K.Lo2 * ρ * ρ)
        return float(k[0]) if np.ndim(ν) == 0 else k

    def __del__(self):
        try:
            for h, ctx in self._dev.values():
                if ctx.h:
                    lib().cs_cia_free(h)
        except Exception:
            pass


class CIA:
    """CIA(ciatables, g₁, g₂) / CIA(ciatables, gases) -- collision_induced_absorption.jl:431-465"""

    def __init__(self, ciatables, *gases):
        if len(gases) == 1 and isinstance(gases[0], (tuple, list)):
            gases = tuple(gases[0])
        assert len(gases) > 0, "no Gas objects provided, cannot create CIA object"
        f1, f2 = ciatables.formulae

        def findgas(f):
            idx = [g for g in gases if g.formula == f]
            assert len(idx) > 0, f"pairing failed for {ciatables.name} CIA, gas {f} is missing"
            assert len(idx) == 1, f"pairing failed for {ciatables.name} CIA, duplicate {f} gases found"
            return idx[0]

        self.name, self.formulae, self.x = ciatables.name, ciatables.formulae, ciatables
        self.g1, self.g2 = findgas(f1), findgas(f2)

    def add_to(self, ws, T, P):
        C1 = f64(self.g1.concentration(T, P))
        C2 = f64(self.g2.concentration(T, P))
        check(lib().cs_sigma_add_cia(ws.h, self.x.handle(ws.ctx), ptr(T), ptr(P), ptr(C1), ptr(C2)))
