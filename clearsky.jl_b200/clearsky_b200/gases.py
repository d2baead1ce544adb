"""Gas objects: AtmosphericDomain, Gas (+ OpacityTable handle), GrayGas, SemiGrayGas, LineGas.

Reference: src/absorption/gases.jl -- AtmosphericDomain :26-61, OpacityTable :68-85, bake :97-145,
Gas :205-336, GrayGas :342-360, SemiGrayGas :366-386.  Seam S2 of SURVEY.md section 8b: the `Gas`
constructor evaluates fC on the (T,P) grid host-side and makes ONE cs_bake call; the table stays on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from . import constants as K
from ._lib import check, f64, lib, ptr
from .line_shapes import DEFAULT_CUT, device_lines, shape_id
from .par import SpectralLines
from .util import chebygrid


class AtmosphericDomain:
    """AtmosphericDomain(Trange, nT, Prange, nP) -- gases.jl:45-61"""

    def __init__(self, Trange, nT, Prange, nP):
        assert all(t > 0 for t in Trange), "temperature range must be positive"
        assert all(p > 0 for p in Prange), "pressure range must be positive"
        assert all(t >= K.TMIN for t in Trange), f"minimum temperature with Qref/Q accuracy is {K.TMIN} K"
        assert all(t <= K.TMAX for t in Trange), f"maximum temperature with Qref/Q accuracy is {K.TMAX} K"
        assert Trange[0] < Trange[1] and Prange[0] < Prange[1]
        self.T = chebygrid(float(Trange[0]), float(Trange[1]), int(nT))
        self.P = np.exp(chebygrid(np.log(float(Prange[0])), np.log(float(Prange[1])), int(nP)))
        self.Tmin, self.Tmax, self.nT = float(Trange[0]), float(Trange[1]), int(nT)
        self.Pmin, self.Pmax, self.nP = float(Prange[0]), float(Prange[1]), int(nP)


def _as_conc(fC):
    """accept a number (uniform concentration) or a callable fC(T,P)"""
    if callable(fC):
        return fC
    c = float(fC)
    return lambda T, P: c


def checkν(ν):
    """gases.jl:90-95 (one comparison pass; ascending + first element non-negative implies all non-negative)"""
    assert len(ν) < 2 or bool(np.all(ν[1:] > ν[:-1])), "wavenumbers must be unique and in ascending order"
    assert len(ν) == 0 or ν[0] >= 0, "wavenumbers must be positive"


def _meanmolarmass(sl):
    """mean molar mass of a line list, Σ(A·μ)/ΣA over the lines (gases.jl:233) -- informational only"""
    return float(np.dot(sl.A, sl.μ) / np.sum(sl.A))


class AbstractGas:
    pass


class Gas(AbstractGas):
    """Gas(sl, fC, ν, Ω, shape!=voigt!, Δνcut=25) -- gases.jl:225-238.  The OpacityTables live on the GPU."""

    def __init__(self, sl, fC, ν, Ω, shape="voigt", Δνcut=25.0, keep_block=False, ctx=None, **kwargs):
        if isinstance(sl, str):
            sl = SpectralLines.from_file(sl, **kwargs)   # gases.jl:240-249
        assert len(ν) > 0
        self.ctx = ctx or _lib.default_context()
        self.name, self.formula = sl.name, sl.formula
        self.μ = _meanmolarmass(sl)
        self.ν = f64(np.asarray(ν, dtype=np.float64))       # shared with the caller, never written
        checkν(self.ν)
        self.Ω = Ω
        self.fC = _as_conc(fC)
        sid = shape_id(shape)
        # C[i,j] = fC(T_i, P_j), Julia column-major -> index i + nT*j
        Cg = np.empty((Ω.nP, Ω.nT))
        for j, P in enumerate(Ω.P):
            for i, T in enumerate(Ω.T):
                c = float(self.fC(T, P))
                assert 0 <= c <= 1, \
                    f"gas molar concentrations must be in [0,1], not {c} (encountered @ {T} K, {P} Pa)"
                Cg[j, i] = c
        self.h = C.c_void_p()
        dl = device_lines(sl, self.ctx)
        check(lib().cs_bake(dl.h, sid, len(self.ν), ptr(self.ν), Ω.nT, ptr(f64(Ω.T)), Ω.nP, ptr(f64(Ω.P)),
                            ptr(f64(Cg)), float(Δνcut), 1 if keep_block else 0, C.byref(self.h)))
        self.timers = self.ctx.timers()
        nz = C.c_int64(0)
        check(lib().cs_table_info(self.h, None, None, None, C.byref(nz)))
        self.nzeroed = nz.value

    @classmethod
    def _from_handle(cls, other, fC):
        g = cls.__new__(cls)
        g.__dict__.update(other.__dict__)
        g._parent = other   # keeps the table alive; reconcentrate does not copy (gases.jl:302-303)
        g.fC = fC
        return g

    def rawσ(self, *args):
        """rawσ(g, T, P): cross-sections of all wavenumbers, without the concentration (gases.jl:263); T, P may be
        vectors (nodes) -> array [nnode, nν].  rawσ(g, i, T, P): single wavenumber index i (0-based) (gases.jl:256)."""
        if len(args) == 3:
            i, T, P = args
            return self.rawσ(T, P)[..., int(i)]
        T, P = args
        scalar = np.ndim(T) == 0
        T, P = f64(np.atleast_1d(T)), f64(np.atleast_1d(P))
        out = np.empty((len(T), len(self.ν)))
        check(lib().cs_table_eval(self.h, len(T), ptr(T), ptr(P), ptr(out)))
        return out[0] if scalar else out

    def concentration(self, T, P):
        """gases.jl:270-275"""
        if np.ndim(T) == 0:
            return self.fC(T, P)
        return np.array([self.fC(t, p) for t, p in zip(T, P)], dtype=np.float64)

    def __call__(self, *args):
        """(g::Gas)(T, P) = concentration * rawσ (gases.jl:281); (g::Gas)(i, T, P) for one wavenumber index (gases.jl:278)"""
        if len(args) == 3:
            i, T, P = args
            return self(T, P)[..., int(i)]
        T, P = args
        c = self.concentration(T, P)
        r = self.rawσ(T, P)
        return c * r if np.ndim(T) == 0 else np.asarray(c)[:, None] * r

    def σblock(self):
        """the baked σ[nν, nT, nP] block (Julia layout) to build stock OpacityTables host-side"""
        out = np.empty((self.Ω.nP, self.Ω.nT, len(self.ν)))
        check(lib().cs_table_block(self.h, ptr(out)))
        return out

    def reconcentrate(self, fC):
        """reconcentrate(g, fC) -- gases.jl:292-320 (shares the tables, new concentration)"""
        f = _as_conc(fC)
        for P in self.Ω.P:
            for T in self.Ω.T:
                c = f(T, P)
                assert 0 <= c <= 1.0, f"gas molar concentrations must be in [0,1], not {c}"
        return Gas._from_handle(self, f)

    # -- contribution to Σ at the quadrature nodes (absorbers.jl:84-95)
    def add_to(self, ws, T, P):
        Cn = f64(self.concentration(T, P))
        check(lib().cs_sigma_add_table(ws.h, self.h, ptr(T), ptr(P), ptr(Cn)))

    def __del__(self):
        try:
            if getattr(self, "_parent", None) is None and self.h and self.ctx.h:
                lib().cs_table_free(self.h)
        except Exception:
            pass


class LineGas(AbstractGas):
    """Engine-native exact gas: Σ is the line-by-line sum at each node's own (T, P, C·P) -- no table.
    Same constructor shape as Gas minus the domain.  Equivalent to a Gas whose table has no interpolation
    error; used for BASELINE.json configs[1] (line×ν×layer evaluations on the actual layers)."""

    def __init__(self, sl, fC, ν, shape="voigt", Δνcut=None, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.name, self.formula = sl.name, sl.formula
        self.μ = _meanmolarmass(sl)
        self.ν = f64(np.asarray(ν, dtype=np.float64))       # shared with the caller, never written
        checkν(self.ν)
        self.fC = _as_conc(fC)
        self.sid = shape_id(shape)
        self.Δνcut = DEFAULT_CUT[self.sid] if Δνcut is None else float(Δνcut)
        self._sl = sl          # uploaded on first use (cs_lines_upload runs on the context's copy stream: inside sigma_nodes the
        self.Ω = None          # upload of this gas overlaps the line sum of the previous one)

    @property
    def dl(self):
        return device_lines(self._sl, self.ctx)

    def concentration(self, T, P):
        if np.ndim(T) == 0:
            return self.fC(T, P)
        return np.array([self.fC(t, p) for t, p in zip(T, P)], dtype=np.float64)

    def evals_per_node(self):
        return self.dl.count_evals(self.ν, self.Δνcut)

    def add_to(self, ws, T, P):
        Cn = f64(self.concentration(T, P))
        check(lib().cs_sigma_add_lines(ws.h, self.dl.h, self.sid, ptr(T), ptr(P), ptr(Cn), self.Δνcut))


class GrayGas(AbstractGas):
    """GrayGas(σ, ν) -- gases.jl:342-360"""

    def __init__(self, σ, ν):
        self.name = self.formula = "Gray"
        self.μ = float("nan")
        self.ν = f64(np.array(ν, dtype=np.float64))
        self.σ = float(σ)

    def __call__(self, *_):
        return self.σ

    def add_to(self, ws, T, P):
        check(lib().cs_sigma_add_gray(ws.h, self.σ, float("inf")))


class SemiGrayGas(AbstractGas):
    """SemiGrayGas(σ, ν, νcut) -- gases.jl:366-386"""

    def __init__(self, σ, ν, νcut):
        self.name = self.formula = "SemiGray"
        self.μ = float("nan")
        self.ν = f64(np.array(ν, dtype=np.float64))
        self.νcut = float(νcut)
        self.σ = float(σ)

    def __call__(self, i, *_):
        return self.σ if self.ν[i] <= self.νcut else 0.0

    def add_to(self, ws, T, P):
        check(lib().cs_sigma_add_gray(ws.h, self.σ, self.νcut))
