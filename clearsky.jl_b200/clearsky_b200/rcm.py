"""Radiative-convective model driver on top of the GPU flux path.

Reference: src/radiative_convective.jl -- RCM :6-103, heating! :109-144, step! :147-151, jacobian! :154-171.
The arithmetic outside `radiate!` is O(np) and stays on the host exactly as in the reference; every
`heating!` is one GPU flux solve with the AcceleratedAbsorber (which, like the reference, is NOT updated
between steps: heating! never calls update!, radiative_convective.jl:109-144).
"""
import numpy as np

from .absorbers import AcceleratedAbsorber, unifyabsorbers
from .core import Discretized, FluxPack
from .fluxes import fluxes_batch, radiate_
from .sharding import ShardedAbsorber
from .util import AtmosphericProfile


class RCM:
    """RCM(Pₑ, Tₑ, g, 𝒻μ, 𝒻S, 𝒻a, 𝒻cₚ, cₛ, absorbers...; core=Discretized(), radmul=2) -- radiative_convective.jl:42-103"""

    def __init__(self, Pe, Te, g, fμ, fS, fa, fcp, cs, *absorbers, core=None, radmul=2):
        Pe = np.asarray(Pe, dtype=np.float64)
        Te = np.asarray(Te, dtype=np.float64)
        idx = np.argsort(Pe, kind="stable")
        self.Pe, self.Te = Pe[idx].copy(), Te[idx].copy()
        n = len(self.Pe)
        assert len(self.Te) == n, "must have same number of initial temperature and pressure values"
        # cell centres + surface (radiative_convective.jl:62-69)
        self.P = np.empty(n)
        self.T = np.empty(n)
        self.P[:-1] = (self.Pe[:-1] + self.Pe[1:]) / 2
        self.T[:-1] = (self.Te[:-1] + self.Te[1:]) / 2
        self.P[-1], self.T[-1] = self.Pe[-1], self.Te[-1]
        # extra radiative nodes by weighted averaging (:71-85)
        assert (radmul % 2 == 0) or (radmul == 1), "radmul must be an even integer or 1"
        nrad = radmul * (n - 1) + 1
        Pr = np.empty(nrad)
        P1, P2 = self.Pe[:-1], self.Pe[1:]
        Pr[: n - 1] = P1
        i = n - 1
        for j in range(2, radmul + 1):
            w1 = j - 1
            w2 = radmul - w1
            Pr[i: i + n - 1] = (w1 * P1 + w2 * P2) / radmul
            i += n - 1
        Pr[-1] = self.Pe[-1]
        self.Pr = np.sort(Pr)
        if len(absorbers) == 1 and isinstance(absorbers[0], ShardedAbsorber):
            # ν-sharded over a DeviceGroup: one AcceleratedAbsorber per slice, resident on the GPU that owns it
            self.A = absorbers[0].accelerate(self.Te, self.Pe)
            self.sharded = True
            nν = len(self.A.ν)
        else:
            U, _, nν = unifyabsorbers(absorbers)
            self.A = U if isinstance(U, AcceleratedAbsorber) else AcceleratedAbsorber(self.Te, self.Pe, U)   # :89
            self.sharded = False
        self.ν, self.nν = self.A.ν, nν
        self.g, self.cs = float(g), float(cs)
        self.fμ, self.fS, self.fa, self.fcp = fμ, fS, fa, fcp
        self.core = core or Discretized()
        self.F = FluxPack(nrad, 1 if self.sharded else nν)     # sharded: only the integrated fluxes come back
        self.np = n
        self.R = np.zeros(n)
        self.H = np.zeros(n)
        self.J = np.zeros((n, n))

    def heating_(self):
        """heating!(ℛ) -- radiative_convective.jl:109-144"""
        fT = AtmosphericProfile(self.P, self.T)
        if self.sharded:
            self.F.Fup[:], self.F.Fdn[:], self.F.Fnet[:] = self.A.fluxes(self.Pr, self.g, fT, self.fμ, self.fS, self.fa,
                                                                         core=self.core)
        else:
            radiate_(self.F, self.core, self.Pr, self.g, fT, self.fμ, self.fS, self.fa, self.A, materialize=False)
        fF = AtmosphericProfile(self.Pr, self.F.Fnet)
        self.R[:] = -fF(self.Pe)
        for i in range(self.np - 1):
            cp = self.fcp(self.T[i], self.P[i]) if callable(self.fcp) else float(self.fcp)
            ΔP = self.Pe[i + 1] - self.Pe[i]
            ΔR = self.R[i] - self.R[i + 1]
            self.H[i] = (self.g / cp) * ΔR / ΔP
        self.H[-1] = self.R[-1] / self.cs
        return None

    def step_(self, Δt):
        """step!(ℛ, Δt) -- radiative_convective.jl:147-151"""
        self.heating_()
        self.T += Δt * self.H
        return None

    def _heating_from(self, Fnet):
        """the O(np) tail of heating! (radiative_convective.jl:123-143) for a given net-flux profile -> H"""
        fF = AtmosphericProfile(self.Pr, Fnet)
        R = -fF(self.Pe)
        H = np.empty(self.np)
        for i in range(self.np - 1):
            cp = self.fcp(self.T[i], self.P[i]) if callable(self.fcp) else float(self.fcp)
            H[i] = (self.g / cp) * (R[i] - R[i + 1]) / (self.Pe[i + 1] - self.Pe[i])
        H[-1] = R[-1] / self.cs
        return H

    def jacobian_(self, ϵ=1.0, batched=True):
        """jacobian!(ℛ, ϵ) -- radiative_convective.jl:154-171.  The np+1 flux solves differ only in the temperature profile
        and the AcceleratedAbsorber ignores T, so they run as ONE batched call (cs_fluxes_batch) that shares every layer
        depth and transmittance; batched=False is the reference's loop of heating! calls."""
        if batched and not self.sharded and (not callable(self.fμ)):
            Ts = [self.T.copy()]
            for i in range(self.np):
                T = self.T.copy()
                T[i] += ϵ
                Ts.append(T)
            Fup, Fdn = fluxes_batch(self.Pr, self.g, [AtmosphericProfile(self.P, T) for T in Ts], self.fμ, self.fS, self.fa,
                                    self.A, core=self.core)
            H0 = self._heating_from(Fup[0] - Fdn[0])
            self.F.Fup[:], self.F.Fdn[:], self.F.Fnet[:] = Fup[0], Fdn[0], Fup[0] - Fdn[0]
            self.H[:] = H0
            self.R[:] = -AtmosphericProfile(self.Pr, self.F.Fnet)(self.Pe)
            for i in range(self.np):
                self.J[:, i] = (self._heating_from(Fup[i + 1] - Fdn[i + 1]) - H0) / ϵ
            return None
        self.heating_()
        H = self.H.copy()
        for i in range(self.np):
            self.T[i] += ϵ
            self.heating_()
            self.J[:, i] = (self.H - H) / ϵ
            self.T[i] -= ϵ
        return None
