"""Line-shape operators backed by the CUDA library (seam S1 of SURVEY.md section 8b).

Reference: src/absorption/line_shapes.jl -- doppler[!] :188-235, lorentz[!] :301-348, voigt[!] :399-448,
PHCO2[!] :514-564.  Same argument order (ν, sl, T, P, Pₚ, Δνcut); the in-place forms fill `σ`.
All arithmetic happens on the GPU; there is no host fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CS_DOPPLER, CS_LORENTZ, CS_PHCO2, CS_VOIGT, check, f64, lib, ptr

SHAPES = {"doppler": CS_DOPPLER, "lorentz": CS_LORENTZ, "voigt": CS_VOIGT, "PHCO2": CS_PHCO2,
          "phco2": CS_PHCO2}
DEFAULT_CUT = {CS_DOPPLER: 25.0, CS_LORENTZ: 25.0, CS_VOIGT: 25.0, CS_PHCO2: 500.0}


def shape_id(shape):
    if isinstance(shape, str):
        return SHAPES[shape]
    if callable(shape) and getattr(shape, "shape_id", None) is not None:
        return shape.shape_id
    return int(shape)


class DeviceLines:
    """SpectralLines resident on one device (cs_lines)."""

    def __init__(self, sl, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.h = C.c_void_p()
        niso, ncheb, cheb, has = sl.cheb_table()
        self._keep = [f64(sl.ν), f64(sl.S), f64(sl.γa), f64(sl.γs), f64(sl.Epp), f64(sl.na), f64(sl.μ)]
        iso = np.ascontiguousarray(sl.I, dtype=np.int16)
        check(lib().cs_lines_upload(
            self.ctx.h, len(sl.ν), *[ptr(a) for a in self._keep],
            iso.ctypes.data_as(C.POINTER(C.c_int16)), niso,
            ncheb.ctypes.data_as(C.POINTER(C.c_int32)), ptr(np.ascontiguousarray(cheb)),
            has.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(self.h)))
        self.N = len(sl.ν)
        rng = getattr(sl, "grid_range", None)      # set by sharding.slice_lines on the line list of a ν slice
        if rng is not None:
            check(lib().cs_lines_set_grid_range(self.h, float(rng[0]), float(rng[1])))

    def count_evals(self, ν, Δνcut):
        ν = f64(ν)
        n = C.c_int64(0)
        check(lib().cs_count_evals(self.h, len(ν), ptr(ν), float(Δνcut), C.byref(n)))
        return n.value

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                lib().cs_lines_free(self.h)
        except Exception:
            pass


def device_lines(sl, ctx=None):
    """upload once per (SpectralLines, context) and cache on the object"""
    ctx = ctx or _lib.default_context()
    cache = sl.__dict__.setdefault("_dev", {})
    if id(ctx) not in cache:
        cache[id(ctx)] = DeviceLines(sl, ctx)
    return cache[id(ctx)]


def xsec(shape, ν, sl, T, P, Pp, Δνcut=None, ctx=None):
    """batched shape!: σ[k, :] = shape(ν, sl, T[k], P[k], Pp[k], Δνcut) for all nodes k in one call."""
    sid = shape_id(shape)
    ν = f64(np.atleast_1d(ν))
    T, P, Pp = (f64(np.atleast_1d(x)) for x in (T, P, Pp))
    assert len(T) == len(P) == len(Pp)
    cut = DEFAULT_CUT[sid] if Δνcut is None else float(Δνcut)
    dl = device_lines(sl, ctx)
    σ = np.empty((len(T), len(ν)), dtype=np.float64)
    check(lib().cs_xsec(dl.h, sid, len(ν), ptr(ν), len(T), ptr(T), ptr(P), ptr(Pp), cut, ptr(σ)))
    return σ


def _make(name, sid):
    def inplace(σ, ν, sl, T, P, Pp, Δνcut=DEFAULT_CUT[sid]):
        out = xsec(sid, ν, sl, [T], [P], [Pp], Δνcut)[0]
        σ[...] = out
        return None

    def vector(ν, sl, T, P, Pp, Δνcut=DEFAULT_CUT[sid]):
        scalar = np.ndim(ν) == 0
        out = xsec(sid, ν, sl, [T], [P], [Pp], Δνcut)[0]
        return float(out[0]) if scalar else out

    inplace.__name__ = name + "_b200_inplace"
    inplace.shape_id = sid
    inplace.__doc__ = f"{name}!(σ, ν, sl, T, P, Pₚ, Δνcut) on the GPU (line_shapes.jl)"
    vector.__name__ = name
    vector.shape_id = sid
    vector.__doc__ = f"{name}(ν, sl, T, P, Pₚ, Δνcut) on the GPU (line_shapes.jl)"
    return inplace, vector


doppler_b200_inplace, doppler = _make("doppler", CS_DOPPLER)
lorentz_b200_inplace, lorentz = _make("lorentz", CS_LORENTZ)
voigt_b200_inplace, voigt = _make("voigt", CS_VOIGT)
PHCO2_b200_inplace, PHCO2 = _make("PHCO2", CS_PHCO2)


def _line_params(sl, T, P, Pp, which, ctx=None):
    dl = device_lines(sl, ctx)
    out = np.empty(dl.N)
    args = [None, None, None]
    args[which] = ptr(out)
    check(lib().cs_line_params(dl.h, float(T), float(P), float(Pp), *args))
    return out


def scaleintensity(sl, T, ctx=None):
    """scaleintensity(sl, :, T): temperature-scaled intensity of every line (line_shapes.jl:107-132), on the GPU"""
    return _line_params(sl, T, 1.0, 0.0, 0, ctx)


def αdoppler(sl, T, ctx=None):
    """αdoppler(sl, :, T) (line_shapes.jl:144-148)"""
    return _line_params(sl, T, 1.0, 0.0, 1, ctx)


def γlorentz(sl, T, P, Pp, ctx=None):
    """γlorentz(sl, :, T, P, Pₚ) (line_shapes.jl:255-261)"""
    return _line_params(sl, T, P, Pp, 2, ctx)
