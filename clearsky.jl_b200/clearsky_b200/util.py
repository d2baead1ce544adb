"""Small host-side utilities mirrored from the reference.

Reference: src/util.jl:19-33 (pressuregrid, trapz); BasicInterpolators.chebygrid (third party; used
at src/absorption/gases.jl:57-58 and src/util.jl:22); src/atmospherics.jl:6-26 (AtmosphericProfile).
"""
import numpy as np


def chebygrid(*args):
    """chebygrid(n) = cos(pi*k/(n-1)), k = n-1..0 (ascending); chebygrid(a, b, n) = affine map."""
    if len(args) == 1:
        n = int(args[0])
        return np.cos(np.pi * np.arange(n - 1, -1, -1) / (n - 1))
    a, b, n = args
    return (chebygrid(n) + 1) * ((b - a) / 2) + a


def pressuregrid(Pt, Ps, n):
    """util.jl:19-23"""
    assert Ps > Pt
    assert n >= 3
    return np.exp(chebygrid(np.log(Pt), np.log(Ps), n))


def trapz(x, y):
    """util.jl:26-33 (sequential sum, same operation order)"""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    assert len(x) == len(y), "vectors must be equal length"
    s = 0.0
    terms = (x[1:] - x[:-1]) * (y[:-1] + y[1:]) / 2
    for t in terms:
        s += t
    return s


def findcell(q, x):
    """cell index i of [x_i, x_{i+1}] containing q, clamped to the end cells (NoBoundaries)."""
    n = len(x)
    i = int(np.searchsorted(x, q, side="right")) - 1
    return min(max(i, 0), n - 2)


class LinearInterpolator:
    """BasicInterpolators.LinearInterpolator(x, y, NoBoundaries()): linear extrapolation off the ends."""

    def __init__(self, x, y):
        self.x = np.array(x, dtype=np.float64)
        self.y = np.array(y, dtype=np.float64)
        assert len(self.x) == len(self.y) and len(self.x) >= 2
        assert np.all(np.diff(self.x) > 0)

    def __call__(self, q):
        q = np.asarray(q, dtype=np.float64)
        i = np.clip(np.searchsorted(self.x, q, side="right") - 1, 0, len(self.x) - 2)
        x, y = self.x, self.y
        return (q - x[i]) * (y[i + 1] - y[i]) / (x[i + 1] - x[i]) + y[i]


class AtmosphericProfile:
    """atmospherics.jl:6-26: piecewise linear in ln P with end-cell extrapolation."""

    def __init__(self, P, y):
        P = np.asarray(P, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        assert len(P) == len(y), "cannot form AtmosphericProfile with unequal numbers of points"
        idx = np.argsort(P, kind="stable")
        self.ϕ = LinearInterpolator(np.log(P[idx]), y[idx])

    def __call__(self, P, *_):
        return self.ϕ(np.log(P))
