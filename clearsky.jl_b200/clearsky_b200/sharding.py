"""ν-sharded multi-GPU evaluation: contiguous wavenumber slices balanced by cost, lines within
slice ± cut-off only, global trapezoid weights (every interval counted exactly once, no halo), and ONE all-reduce of
the 2·np spectrally integrated fluxes.

Two host models use the same slicing:
  * one process per GPU under torchrun (bench.py): partial fluxes go straight into a torch tensor via
    cs_fluxes_device and torch.distributed (NCCL) reduces them;
  * one process driving all GPUs (`DeviceGroup` + `sharded_fluxes` below, what the Julia wrapper would do with one
    task per device): cs_group_* owns the contexts and the NCCL communicator.
"""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib
from ._lib import check, f64, lib, ptr
from .absorbers import SigmaWorkspace
from .core import Discretized
from .fluxes import _unique_nodes, _vec, checkazimuth, formprofile, lobattoevaluations
from .gases import LineGas
from .par import SpectralLines
from .quadrature import lobattonodes, streamnodes


def per_point_counts(ν, νl, cut):
    """number of lines within ± cut of every wavenumber (νl sorted)"""
    lo = np.searchsorted(νl, ν - cut, side="left")
    hi = np.searchsorted(νl, ν + cut, side="right")
    return (hi - lo).astype(np.int64)


# per-ν cost = evaluations * (1 + kappa * ν): fitted to the per-slice line-sum times of an 8-way split of C2 (tools/slice_balance.py).
# The slope is the near-centre work (Doppler widths grow with ν); in expansion mode the far wings are nearly free, so it weighs more.
KAPPA = {"direct": 1.24e-4, "expansion": 3.2e-4}   # direct refitted for the two-launch line sum with the wider near band of the default Faddeyeva borders


def slice_cost(ν, line_lists, cut, kappa=None, farfield="direct"):
    """per-ν cost model for balancing slices: evaluations, inflated by the near-centre work that grows with ν
    (Doppler widths are proportional to ν, so the share of near-centre lines per tile is too)."""
    counts = sum(per_point_counts(ν, νl, cut) for νl in line_lists)
    return counts * (1.0 + (KAPPA[farfield] if kappa is None else kappa) * ν)


def balanced_slices(cost, n):
    """n contiguous index slices with ~equal sums of `cost` -> n+1 edges"""
    c = np.concatenate(([0], np.cumsum(cost, dtype=np.float64)))
    edges = [int(np.searchsorted(c, c[-1] * k / n)) for k in range(n + 1)]
    edges[0], edges[-1] = 0, len(cost)
    for k in range(1, n + 1):
        edges[k] = max(edges[k], edges[k - 1])
    return edges


def trapz_weights(ν):
    """per-point weights of trapz(ν, ·) (util.jl:26-33): w_j = (Δν_{j-1} + Δν_j)/2"""
    d = np.diff(ν)
    w = np.zeros(len(ν))
    w[:-1] += d / 2
    w[1:] += d / 2
    return w


def slice_lines(sl, νlo, νhi, cut, grid=None):
    """every line a slice [νlo, νhi] can see (the per-point inclusive rule decides inside the kernel).
    grid = (first, last) point of the GLOBAL wavenumber grid: recorded on the slice's line list so that the library applies
    the strict `includedlines` prefilter (line_shapes.jl:18-22) to the grid the reference would see -- a line exactly one
    cut-off away from an INTERIOR slice edge then counts exactly as in the unsharded run (cs_lines_set_grid_range)."""
    keep = (sl.ν >= νlo - cut - 1e-9) & (sl.ν <= νhi + cut + 1e-9)
    if not keep.any():
        keep[np.argmin(np.abs(sl.ν - νlo))] = True      # keep one (out-of-window) line: an upload cannot be empty
    out = SpectralLines(sl.name, sl.formula, int(keep.sum()), sl.M, sl.I[keep], sl.μ[keep], sl.A[keep], sl.ν[keep],
                        sl.S[keep], sl.γa[keep], sl.γs[keep], sl.Epp[keep], sl.na[keep])
    if grid is not None:
        out.grid_range = (float(grid[0]), float(grid[1]))
    return out


class DeviceGroup:
    """cs_group: one context per device + a single-node NCCL communicator, all inside this process"""

    def __init__(self, devices=None):
        if devices is None:
            devices = list(range(_lib.device_count()))
        self.devices = list(devices)
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        self.h = C.c_void_p()
        check(lib().cs_group_create(len(self.devices), arr, C.byref(self.h)))
        self.ctx = []
        for i in range(len(self.devices)):
            c = C.c_void_p()
            check(lib().cs_group_ctx(self.h, i, C.byref(c)))
            self.ctx.append(_lib.Context(self.devices[i], borrowed=c))
        self.pool = ThreadPoolExecutor(max_workers=len(self.devices))

    def __len__(self):
        return len(self.devices)

    def buffer(self, i, count):
        p = C.c_void_p()
        check(lib().cs_group_buffer(self.h, i, int(count), C.byref(p)))
        return p

    def allreduce_sum(self, count):
        check(lib().cs_group_allreduce_sum(self.h, int(count)))

    def read(self, i, count):
        out = np.empty(int(count))
        check(lib().cs_group_read(self.h, i, int(count), ptr(out)))
        return out

    def close(self):
        if self.h:
            self.pool.shutdown(wait=True)
            for c in self.ctx:
                c.h = C.c_void_p()
            lib().cs_group_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedAbsorber:
    """Any set of absorbers sharded over a DeviceGroup by contiguous ν slices: tables, line
    gases, CIA and accelerated absorbers stay resident on the GPU that owns the slice; every `fluxes` call is Σ + K6/K7 per
    device (one host thread each) and ONE all-reduce of the 2·np integrated fluxes.

    build(ν_slice, ctx) -> absorber or tuple of absorbers living on `ctx` (e.g. `Gas(sl, fC, ν_slice, Ω, ctx=ctx)`),
    called once per device; cost: optional per-ν weights used to balance the slices (default: equal point counts)."""

    def __init__(self, group, ν, build, cost=None):
        from .absorbers import AbstractAbsorber, UnifiedAbsorber
        self.group = group
        self.ν = f64(np.asarray(ν, dtype=np.float64))
        assert np.all(np.diff(self.ν) > 0), "wavenumbers must be unique and in ascending order"
        n = len(group)
        self.edges = balanced_slices(np.ones(len(self.ν)) if cost is None else cost, n)
        self.wg = trapz_weights(self.ν)
        self.parts = []

        def make(i):
            a, b = self.edges[i], self.edges[i + 1]
            if b <= a:
                return None
            νs = np.ascontiguousarray(self.ν[a:b])
            objs = build(νs, group.ctx[i])
            A = objs if isinstance(objs, AbstractAbsorber) else UnifiedAbsorber(*(objs if isinstance(objs, (tuple, list)) else (objs,)))
            return dict(a=a, b=b, ν=νs, A=A, ws={}, w=np.ascontiguousarray(self.wg[a:b]))

        self.parts = list(group.pool.map(make, range(n)))     # uploads / bakes of all devices run concurrently

    def accelerate(self, T, P):
        """wrap every slice in an AcceleratedAbsorber built at the levels (T, P) (absorbers.jl:135-157); returns self"""
        from .absorbers import AcceleratedAbsorber

        def acc(i):
            part = self.parts[i]
            if part is not None and not isinstance(part["A"], AcceleratedAbsorber):
                part["A"] = AcceleratedAbsorber(T, P, part["A"], ctx=self.group.ctx[i])
        list(self.group.pool.map(acc, range(len(self.group))))
        return self

    def update(self, T):
        """update!(A, T) on every slice (absorbers.jl:173-200)"""
        list(self.group.pool.map(lambda i: self.parts[i] is not None and self.parts[i]["A"].update(T), range(len(self.group))))

    def checkpressures(self, Ps, Pt):
        for part in self.parts:
            if part is not None:
                part["A"].checkpressures(Ps, Pt)

    def fluxes(self, P, g, T, μ, fS=None, fa=None, core=None, θs=0.841):
        """fluxes(P, g, T, μ, 𝒻S, 𝒻a, absorbers...; core, θₛ) (fluxes.jl:311-340) -> (F⁺, F⁻, Fnet)"""
        core = core or Discretized()
        group, n = self.group, len(self.group)
        P = f64(np.asarray(P, dtype=np.float64))
        assert np.all(np.diff(P) >= 0), "pressure coordinates must be in ascending order (sorted)"
        checkazimuth(θs)
        self.checkpressures(P[-1], P[0])
        fT, fμ = formprofile(P, T), formprofile(P, μ)
        Tl, μl, Pn = lobattoevaluations(P, fT, fμ, core.nlobatto)
        Tn, Pq = _unique_nodes(P, Tl, Pn, core.nlobatto)
        Tn, Pq, μl = f64(Tn), f64(Pq), f64(μl)
        Tlev = f64(_vec(fT, P))
        m, W = (f64(x) for x in streamnodes(core.nstream))
        wl = f64(lobattonodes(core.nlobatto)[1])
        npl = len(P)
        fSν = None if fS is None else f64(_vec(fS, self.ν) if callable(fS) else np.full(len(self.ν), float(fS)))
        faν = None if fa is None else f64(_vec(fa, self.ν) if callable(fa) else np.full(len(self.ν), float(fa)))

        def run(i):
            part = self.parts[i]
            buf = group.buffer(i, 2 * npl)
            ctx = group.ctx[i]
            if part is None:     # empty slice: contribute zeros
                z = SigmaWorkspace(self.ν[:1], len(Tn), ctx)
                check(lib().cs_fluxes_device(z.h, npl, ptr(P), core.nlobatto, ptr(wl), ptr(μl), ptr(Tlev), float(g), None, None,
                                             float(θs), core.nstream, ptr(m), ptr(W), ptr(np.zeros(1)), buf))
                return
            a, b = part["a"], part["b"]
            ws = part["ws"].get(len(Tn))
            if ws is None:
                ws = part["ws"][len(Tn)] = SigmaWorkspace(part["ν"], len(Tn), ctx)
            else:
                ws.zero()
            part["A"].sigma_nodes(ws, Tn, Pq)
            check(lib().cs_fluxes_device(ws.h, npl, ptr(P), core.nlobatto, ptr(wl), ptr(μl), ptr(Tlev), float(g),
                                         ptr(np.ascontiguousarray(fSν[a:b])) if fSν is not None else None,
                                         ptr(np.ascontiguousarray(faν[a:b])) if faν is not None else None,
                                         float(θs), core.nstream, ptr(m), ptr(W), ptr(part["w"]), buf))

        list(group.pool.map(run, range(n)))      # one host thread per device; ctypes releases the GIL inside the calls
        group.allreduce_sum(2 * npl)
        F = group.read(0, 2 * npl)
        return F[:npl].copy(), F[npl:].copy(), F[:npl] - F[npl:]


class ShardedLineByLine(ShardedAbsorber):
    """Exact line-by-line gases sharded over a DeviceGroup: every device uploads only the lines within its slice ± the
    cut-off, and slices are balanced by the evaluation-count cost model.   gases: list of (SpectralLines, fC, shape, Δνcut);
    cia: iterable of CIATables (or (CIATables, i, j) tuples), paired with the gases by formula exactly like
    UnifiedAbsorber does (collision_induced_absorption.jl:431-465)."""

    def __init__(self, group, gases, ν, cia=()):
        self.gases = list(gases)
        self.cia = [x[0] if isinstance(x, (tuple, list)) else x for x in cia]
        ν = f64(np.asarray(ν, dtype=np.float64))
        mode = group.ctx[0].get_farfield()
        cost = sum(slice_cost(ν, [sl.ν], cut, farfield=mode if shape in ("voigt", "lorentz", 1, 2) else "direct")
                   for sl, _, shape, cut in self.gases)

        def build(νs, ctx):
            lg = [LineGas(slice_lines(sl, νs[0], νs[-1], cut, grid=(ν[0], ν[-1])), fC, νs, shape, cut, ctx=ctx)
                  for sl, fC, shape, cut in self.gases]
            return tuple(lg) + tuple(self.cia)

        super().__init__(group, ν, build, cost=cost)
        for part in self.parts:
            if part is not None:
                part["gases"] = list(part["A"].gas)


def sharded_fluxes(group, P, g, T, μ, fS, fa, gases, ν, core=None, θs=0.841):
    """one-shot form of ShardedLineByLine(group, gases, ν).fluxes(...) -> (F⁺, F⁻, Fnet)"""
    return ShardedLineByLine(group, gases, ν).fluxes(P, g, T, μ, fS, fa, core=core, θs=θs)
