"""Host-side temperature profiles needed to generate inputs for the hot path.

Only the closed-form DryAdiabat is mirrored (reference: src/atmospherics.jl:290-346, :482-504);
moist adiabats, hydrostatics etc. are out of scope.
"""
import numpy as np

from . import constants as K


def _temperature(P, Ts, Ps, cp, μ):
    """atmospherics.jl:344"""
    return Ts * (P / Ps) ** (K.R / (μ * cp))


def _lapserate(T, P, cp, μ):
    """dry lapse rate dT/dP = R T/(μ cp P) (atmospherics.jl:180-192 dry branch)"""
    return K.R * T / (μ * cp * P)


class DryAdiabat:
    """DryAdiabat(Tₛ, Pₛ, cₚ, μ; Tstrat=0, Ptropo=0, smooth=1e2, Pₜ=1e-9) -- atmospherics.jl:322-341"""

    def __init__(self, Ts, Ps, cp, μ, Tstrat=0.0, Ptropo=0.0, smooth=1e2, Pt=K.Pmin):
        self.Ts, self.Ps, self.Pt, self.cp, self.μ = float(Ts), float(Ps), float(Pt), float(cp), float(μ)
        if Tstrat != 0:
            # invert T(P) = Tstrat in closed form (the reference root-finds with regulafalsi, tol 1e-6)
            Ptropo = Ps * (Tstrat / Ts) ** (μ * cp / K.R)
        elif Ptropo != 0:
            Tstrat = _temperature(Ptropo, Ts, Ps, cp, μ)
        self.Tstrat, self.Ptropo, self.smooth = float(Tstrat), float(Ptropo), float(smooth)
        self.T2 = 0.0
        self.h2 = 0.0
        if Ptropo != 0:
            P2 = Ptropo + smooth
            self.T2 = _temperature(P2, Ts, Ps, cp, μ)
            self.h2 = smooth * _lapserate(self.T2, P2, cp, μ)

    def _scalar(self, P):
        """(Γ::AbstractAdiabat)(P) -- atmospherics.jl:482-504"""
        if P < self.Ptropo:
            return self.Tstrat
        if self.Ptropo != 0 and self.smooth != 0:
            if self.Ptropo < P < self.Ptropo + self.smooth:
                ψ = (P - self.Ptropo) / self.smooth
                T1, T2, h2 = self.Tstrat, self.T2, self.h2
                return ψ**3 * (2 * T1 - 2 * T2 + h2) + ψ**2 * (-3 * T1 + 3 * T2 - h2) + T1
        T = _temperature(P, self.Ts, self.Ps, self.cp, self.μ)
        if T < self.Tstrat:
            return self.Tstrat
        assert T > 0
        return T

    def __call__(self, P, *_):
        if np.ndim(P) == 0:
            return self._scalar(float(P))
        return np.array([self._scalar(float(p)) for p in np.ravel(P)]).reshape(np.shape(P))
