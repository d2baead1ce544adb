"""Absorber containers: UnifiedAbsorber, AcceleratedAbsorber, and the device sigma workspace.

Reference: src/absorption/absorbers.jl -- UnifiedAbsorber :18-99, AcceleratedAbsorber :114-209,
unifyabsorbers :214-223, getwavenumbers :226-235, checkpressures :237-246, pressurelimits :248-256.
`Σ(𝒜, idx, T, P)` (absorbers.jl:95) is evaluated for ALL wavenumbers and ALL quadrature nodes at once
into a device workspace [node][ν]; arbitrary user functions σ(ν,T,P) cannot cross the C ABI and are
pre-evaluated on the host (cs_sigma_add_host).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, f64, lib, ptr
from .cia import CIA, CIATables
from .gases import AbstractGas, Gas, LineGas


class SigmaWorkspace:
    """cs_sigma: Σ at nnode (T,P) nodes for every wavenumber, resident on the GPU"""

    def __init__(self, ν, nnode, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.ν = f64(ν)
        self.nnode = int(nnode)
        self.h = C.c_void_p()
        check(lib().cs_sigma_create(self.ctx.h, len(self.ν), ptr(self.ν), self.nnode, C.byref(self.h)))

    def zero(self):
        check(lib().cs_sigma_zero(self.h))

    def add_host(self, σ):
        σ = f64(σ)
        assert σ.shape == (self.nnode, len(self.ν))
        check(lib().cs_sigma_add_host(self.h, ptr(σ)))

    def read(self):
        out = np.empty((self.nnode, len(self.ν)))
        check(lib().cs_sigma_read(self.h, ptr(out)))
        return out

    def close(self):
        # a workspace that outlives its context (a closed DeviceGroup) must not touch the freed cs_ctx
        if self.h and self.ctx.h:
            lib().cs_sigma_free(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AbstractAbsorber:
    pass


def getwavenumbers(gases):
    """absorbers.jl:226-229"""
    assert len(gases) > 0, "no gas objects found"
    ν0 = gases[0].ν
    assert all(len(g.ν) == len(ν0) and np.array_equal(g.ν, ν0) for g in gases), \
        "gases must have identical wavenumber vectors"
    return ν0


def pressurelimits(gases):
    """absorbers.jl:248-256"""
    g = [x for x in gases if isinstance(x, Gas)]
    if not g:
        return 0.0, np.inf
    return max(x.Ω.Pmin for x in g), min(x.Ω.Pmax for x in g)


def temperaturelimits(gases):
    """absorbers.jl:258-266"""
    g = [x for x in gases if isinstance(x, Gas)]
    if not g:
        return 0.0, np.inf
    return max(x.Ω.Tmin for x in g), min(x.Ω.Tmax for x in g)


class UnifiedAbsorber(AbstractAbsorber):
    """UnifiedAbsorber(absorbers...) -- absorbers.jl:47-77"""

    def __init__(self, *absorbers):
        if len(absorbers) == 1 and isinstance(absorbers[0], (tuple, list)):
            absorbers = tuple(absorbers[0])
        assert len(absorbers) > 0, "no absorbers... nothing to group"
        assert len(set(id(a) for a in absorbers)) == len(absorbers), "duplicate absorbers"
        for a in absorbers:
            if not (isinstance(a, (AbstractGas, CIATables)) or callable(a)):
                raise TypeError("absorbers must only be gases (<: Gas), CIA objects, or functions in the form σ(ν, T, P)")
        self.gas = tuple(a for a in absorbers if isinstance(a, AbstractGas))
        if not self.gas:
            raise ValueError("must have at least one Gas object, which specifies wavenumber samples")
        realgas = tuple(g for g in self.gas if isinstance(g, (Gas, LineGas)))
        self.cia = tuple(CIA(x, realgas) for x in absorbers if isinstance(x, CIATables))
        self.fun = tuple(a for a in absorbers if not isinstance(a, (AbstractGas, CIATables)))
        self.ν = getwavenumbers(self.gas)
        self.nν = len(self.ν)
        # the context the member gases live on (None: the process-wide default context)
        self.ctx = next((g.ctx for g in self.gas if getattr(g, "ctx", None) is not None), None)

    def update(self, T):   # update!(A::UnifiedAbsorber, T) is a no-op (absorbers.jl:80)
        return None

    def sigma_nodes(self, ws, T, P):
        """Σ(U, i, T, P) for all i at all nodes -> accumulated into ws (absorbers.jl:88-95)"""
        T, P = f64(T), f64(P)
        for g in self.gas:
            g.add_to(ws, T, P)
        for c in self.cia:
            c.add_to(ws, T, P)
        if self.fun:
            extra = np.zeros((len(T), self.nν))
            for f in self.fun:
                for m in range(len(T)):
                    try:
                        v = np.asarray(f(self.ν, T[m], P[m]), dtype=np.float64)
                        if v.shape != self.ν.shape:
                            raise ValueError
                    except Exception:
                        v = np.array([f(x, T[m], P[m]) for x in self.ν], dtype=np.float64)
                    extra[m] += v
            ws.add_host(extra)

    def __call__(self, T, P):
        """(U::UnifiedAbsorber)(T, P): Σ for every wavenumber at one (T, P) (absorbers.jl:97-99)"""
        ws = SigmaWorkspace(self.ν, 1, self.ctx)
        self.sigma_nodes(ws, np.array([float(T)]), np.array([float(P)]))
        return ws.read()[0]

    def checkpressures(self, Ps, Pt):
        """absorbers.jl:237-246"""
        assert Ps > Pt, "Pₛ must be greater than Pₜ"
        Pmin, Pmax = pressurelimits(self.gas)
        for P in (Ps, Pt):
            assert P >= Pmin, f"Pressure {P} Pa too low, domain minimum is {Pmin}"
            assert P <= Pmax, f"Pressure {P} Pa too high, domain maximum is {Pmax}"

    def temperaturelimits(self):
        return temperaturelimits(self.gas)


class AcceleratedAbsorber(AbstractAbsorber):
    """AcceleratedAbsorber(T, P, U) -- absorbers.jl:135-157: per-ν linear interpolation of ln σ in ln P at the
    given levels; Σ ignores T (absorbers.jl:203).  update!(A, T) re-evaluates the levels (absorbers.jl:173-200)."""

    def __init__(self, T, P, *absorbers, ctx=None):
        U = absorbers[0] if len(absorbers) == 1 and isinstance(absorbers[0], UnifiedAbsorber) \
            else UnifiedAbsorber(*absorbers)
        self.U = U
        self.ν, self.nν = U.ν, U.nν
        P = np.asarray(P, dtype=np.float64)
        T = np.asarray(T, dtype=np.float64)
        idx = np.argsort(P, kind="stable")
        self.P = f64(P[idx])
        self.T = f64(T[idx])
        self.h = C.c_void_p()
        self._ws = SigmaWorkspace(self.ν, len(self.P), ctx or U.ctx)
        self.update(self.T)

    def update(self, T):
        T = f64(T)
        assert len(T) == len(self.P)
        self._ws.zero()
        self.U.sigma_nodes(self._ws, T, self.P)
        if self.h:
            lib().cs_accel_free(self.h)
            self.h = C.c_void_p()
        check(lib().cs_accel_from_sigma(self._ws.h, ptr(self.P), C.byref(self.h)))
        self.T = T.copy()

    def sigma_nodes(self, ws, T, P):
        check(lib().cs_sigma_add_accel(ws.h, self.h, ptr(f64(P))))

    def checkpressures(self, Ps, Pt):
        self.U.checkpressures(Ps, Pt)

    def temperaturelimits(self):
        return self.U.temperaturelimits()

    def __del__(self):
        try:
            if self.h and self._ws.ctx.h:
                lib().cs_accel_free(self.h)
        except Exception:
            pass


def unifyabsorbers(x):
    """absorbers.jl:214-223 -> (𝒜, ν, nν)"""
    if len(x) == 0:
        raise ValueError("no absorbers")
    if len(x) == 1 and isinstance(x[0], (UnifiedAbsorber, AcceleratedAbsorber)):
        return x[0], x[0].ν, x[0].nν
    U = UnifiedAbsorber(*x)
    return U, U.ν, U.nν
