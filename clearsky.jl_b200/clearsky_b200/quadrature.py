"""Quadrature nodes computed on the host and passed across the C ABI.

Reference: src/core/shared.jl:4-21 (streamnodes, Gauss-Legendre via FastGaussQuadrature),
src/core/discretized.jl:2-9 (lobattonodes, Gauss-Lobatto shifted to [0,1]).
"""
import functools

import numpy as np
from numpy.polynomial import legendre as _leg


def gausslobatto(n):
    """n-point Gauss-Lobatto nodes/weights on [-1,1] (FastGaussQuadrature.gausslobatto)."""
    assert n >= 2
    if n == 2:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    Pn1 = _leg.Legendre.basis(n - 1)
    dP = Pn1.deriv()
    xi = np.sort(dP.roots().real)
    d2P = dP.deriv()
    for _ in range(4):  # Newton polish
        xi = xi - dP(xi) / d2P(xi)
    x = np.concatenate(([-1.0], xi, [1.0]))
    w = 2.0 / (n * (n - 1) * Pn1(x) ** 2)
    x = (x - x[::-1]) / 2  # enforce antisymmetry
    w = (w + w[::-1]) / 2
    return x, w


@functools.lru_cache(maxsize=None)
def _lobattonodes(n):
    x, w = gausslobatto(n)
    return (x + 1) / 2, w / 2


def lobattonodes(n):
    """discretized.jl:2-9"""
    x, w = _lobattonodes(int(n))
    return x.copy(), w.copy()


@functools.lru_cache(maxsize=None)
def _streamnodes(n):
    x, w = _leg.leggauss(n)
    θ = (np.pi / 2) * (x + 1) / 2
    wi = (np.pi / 2) * w / 2
    m = 1 / np.cos(θ)
    W = 2 * np.pi * wi * np.cos(θ) * np.sin(θ)
    return m, W


def streamnodes(n):
    """shared.jl:4-21: (𝓂 = 1/cosθ, 𝒲 = 2π w cosθ sinθ) for θ ∈ [0, π/2]"""
    m, W = _streamnodes(int(n))
    return m.copy(), W.copy()
