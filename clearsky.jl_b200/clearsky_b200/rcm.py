"""Radiative-convective model driver on top of the GPU flux path.

Reference: src/radiative_convective.jl -- RCM :6-103, heating! :109-144, step! :147-151, jacobian! :154-171.
The arithmetic outside `radiate!` is O(np) and stays on the host exactly as in the reference; every
`heating!` is one GPU flux solve with the AcceleratedAbsorber (which, like the reference, is NOT updated
between steps: heating! never calls update!, radiative_convective.jl:109-144).
"""
import ctypes as C

import numpy as np

from ._lib import check, f64, lib, ptr
from .absorbers import AcceleratedAbsorber, SigmaWorkspace, unifyabsorbers
from .core import Discretized, FluxPack
from .fluxes import _eval_spectral, _prepare, _unique_nodes, _vec, fluxes_batch, formprofile, lobattoevaluations, radiate_
from .quadrature import lobattonodes, streamnodes
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

    # ---- device-resident loop (cs_rcm_*): the AcceleratedAbsorber is frozen during an RCM run (heating! never calls
    # update!), so Σ, the layer depths and the stream transmittances are computed once and every step is three kernels
    # replayed from a CUDA graph; T, H, R and the fluxes come back to the host when the call returns
    def _device_handles(self):
        h = self.__dict__.get("_dev")
        if h is not None:
            return h
        core = self.core
        assert isinstance(core, Discretized), "the device-resident loop runs on the Discretized core"
        fT = AtmosphericProfile(self.P, self.T)
        fμ = formprofile(self.Pr, self.fμ)
        m, W = (f64(x) for x in streamnodes(core.nstream))
        wl = f64(lobattonodes(core.nlobatto)[1])
        cp = f64([self.fcp(self.T[i], self.P[i]) if callable(self.fcp) else float(self.fcp) for i in range(self.np - 1)])
        Pr = f64(self.Pr)

        def create(A, ν, ctx, ν_weights, fSν, faν):
            Tl, μl, Pn = lobattoevaluations(Pr, fT, fμ, core.nlobatto)
            Tn, Pq = _unique_nodes(Pr, Tl, Pn, core.nlobatto)
            ws = SigmaWorkspace(ν, len(Tn), ctx)
            A.sigma_nodes(ws, f64(Tn), f64(Pq))
            r = C.c_void_p()
            check(lib().cs_rcm_create(ws.h, self.np, ptr(f64(self.Pe)), ptr(f64(self.P)), ptr(f64(self.T)), ptr(cp), self.cs,
                                      len(Pr), ptr(Pr), core.nlobatto, ptr(wl), ptr(f64(μl)), self.g, ptr(fSν), ptr(faν),
                                      0.841, core.nstream, ptr(m), ptr(W), ptr(ν_weights) if ν_weights is not None else None,
                                      C.byref(r)))
            return r

        self.A.checkpressures(Pr[-1], Pr[0])
        if self.sharded:
            grp = self.A.group
            fSν, faν = _eval_spectral(self.fS, self.A.ν), _eval_spectral(self.fa, self.A.ν)
            hs = []
            for i, part in enumerate(self.A.parts):
                assert part is not None, "every device of the group needs a non-empty ν slice"
                a, b = part["a"], part["b"]
                hs.append(create(part["A"], part["ν"], grp.ctx[i], part["w"],
                                 None if fSν is None else f64(fSν[a:b]), None if faν is None else f64(faν[a:b])))
            h = dict(handles=hs, group=grp)
        else:
            ctx = self.A._ws.ctx
            h = dict(handles=[create(self.A, self.ν, ctx, None, _eval_spectral(self.fS, self.ν), _eval_spectral(self.fa, self.ν))],
                     group=None)
        self._dev = h
        return h

    def steps_(self, Δt, nsteps=1):
        """nsteps × step!(ℛ, Δt) without leaving the device (cs_rcm_step / cs_group_rcm_step).  The heat capacities
        𝒻cₚ(T, P) and the molar masses 𝒻μ are evaluated once, at the temperatures of the first call."""
        h = self._device_handles()
        T = f64(self.T)
        for r in h["handles"]:
            check(lib().cs_rcm_set_temperature(r, ptr(T)))
        if h["group"] is None:
            check(lib().cs_rcm_step(h["handles"][0], float(Δt), int(nsteps)))
        else:
            arr = (C.c_void_p * len(h["handles"]))(*[r.value for r in h["handles"]])
            check(lib().cs_group_rcm_step(h["group"].h, arr, float(Δt), int(nsteps)))
        Tn, H, R = np.empty(self.np), np.empty(self.np), np.empty(self.np)
        nr = len(self.Pr)
        Fup, Fdn, Fnet = np.empty(nr), np.empty(nr), np.empty(nr)
        check(lib().cs_rcm_state(h["handles"][0], ptr(Tn), ptr(H), ptr(R), ptr(Fup), ptr(Fdn), ptr(Fnet)))
        self.T[:], self.H[:], self.R[:] = Tn, H, R
        self.F.Fup[:], self.F.Fdn[:], self.F.Fnet[:] = Fup, Fdn, Fnet
        return None

    def close(self):
        h = self.__dict__.pop("_dev", None)
        if h is not None:
            for r in h["handles"]:
                lib().cs_rcm_free(r)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

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
