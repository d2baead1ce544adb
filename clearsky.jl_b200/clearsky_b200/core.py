"""Numerical-core tag types and FluxPack (reference: src/core/shared.jl:36-121)."""
from dataclasses import dataclass

import numpy as np


class AbstractNumericalCore:
    pass


@dataclass
class Discretized(AbstractNumericalCore):
    """Discretized(; nstream=5, nlobatto=2) -- shared.jl:55-62.  On this engine the Discretized core IS the
    B200 core (the Julia wrapper calls it B200Discretized <: AbstractNumericalCore)."""
    nstream: int = 5
    nlobatto: int = 2


B200Discretized = Discretized


class FluxPack:
    """FluxPack(np, nν) -- shared.jl:73-106.  Arrays use the Julia shapes transposed to C order:
    τ[nν, np-1], M⁺/M⁻[nν, np] (pressure index fastest in memory, exactly like Julia's [np, nν])."""

    def __init__(self, np_, nν):
        self.τ = np.zeros((nν, np_ - 1))
        self.Mup = np.zeros((nν, np_))
        self.Mdn = np.zeros((nν, np_))
        self.Fup = np.zeros(np_)
        self.Fdn = np.zeros(np_)
        self.Fnet = np.zeros(np_)

    @property
    def size(self):
        return (self.Mup.shape[1], self.Mup.shape[0])
