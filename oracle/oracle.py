"""ctypes wrapper of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
PARITY STATUS: parity unpinned (see the header of oracle.c and DESIGN.md).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle.so")
DOPPLER, LORENTZ, VOIGT, PHCO2 = 0, 1, 2, 3
MAXCHEB = 16

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_lib = None


def build(force=False):
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        _lib = C.CDLL(SO)
        for name in ("orc_planck", "orc_faddeyeva985", "orc_cheby_qrefq", "orc_scaleintensity", "orc_alpha_doppler",
                     "orc_gamma_lorentz", "orc_doppler", "orc_lorentz", "orc_voigt", "orc_phco2", "orc_chi_phco2",
                     "orc_fvoigt", "orc_cia_sigma", "orc_linterp", "orc_bilinterp", "orc_opacity_table_eval"):
            getattr(_lib, name).restype = C.c_double
        _lib.orc_count_evals.restype = C.c_int64
        _lib.orc_bake.restype = C.c_int64
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def _pi(a):
    return a.ctypes.data_as(_ip)


def max_threads():
    return int(lib().orc_max_threads())


def _lines_args(sl):
    niso, ncheb, cheb, has = sl.cheb_table()
    arrs = [_f(sl.ν), _f(sl.S), _f(sl.γa), _f(sl.γs), _f(sl.Epp), _f(sl.na), _f(sl.μ)]
    iso = np.ascontiguousarray(sl.I, dtype=np.int16)
    cheb = _f(cheb)
    ncheb = np.ascontiguousarray(ncheb, dtype=np.int32)
    keep = (arrs, iso, cheb, ncheb)
    args = [C.c_int64(len(sl.ν))] + [_p(a) for a in arrs] + [
        iso.ctypes.data_as(C.POINTER(C.c_int16)), C.c_int32(niso),
        ncheb.ctypes.data_as(C.POINTER(C.c_int32)), _p(cheb)]
    return args, keep


def xsec(shape, sl, ν, T, P, Pp, cut, nthreads=1):
    """[nlev, nν] cross-sections; restates shape!(σ, ν, sl, T, P, Pₚ, Δνcut)"""
    ν, T, P, Pp = _f(np.atleast_1d(ν)), _f(np.atleast_1d(T)), _f(np.atleast_1d(P)), _f(np.atleast_1d(Pp))
    out = np.empty((len(T), len(ν)))
    args, keep = _lines_args(sl)
    rc = lib().orc_xsec(C.c_int(shape), *args, C.c_int64(len(ν)), _p(ν), C.c_int64(len(T)), _p(T), _p(P), _p(Pp),
                        C.c_double(cut), _p(out), C.c_int(nthreads))
    if rc != 0:
        raise ValueError("oracle: temperature outside [25,1000] K")
    return out


def count_evals(ν, νl, cut):
    ν, νl = _f(ν), _f(νl)
    return int(lib().orc_count_evals(C.c_int64(len(ν)), _p(ν), C.c_int64(len(νl)), _p(νl), C.c_double(cut)))


def included_lines(ν, νl, cut):
    """the strict prefilter of includedlines(ν::Vector, ...) (line_shapes.jl:18-22)"""
    νl = np.asarray(νl)
    return νl[(νl > np.min(ν) - cut) & (νl < np.max(ν) + cut)]


def bake(shape, sl, ν, Tg, Pg, Cg, cut, nthreads=1):
    """σ block as [nP, nT, nν] (memory order of Julia's σ[nν, nT, nP]); Cg[j, i] = fC(T_i, P_j)"""
    ν, Tg, Pg, Cg = _f(ν), _f(Tg), _f(Pg), _f(Cg)
    out = np.zeros((len(Pg), len(Tg), len(ν)))
    args, keep = _lines_args(sl)
    nz = lib().orc_bake(C.c_int(shape), *args, C.c_int64(len(ν)), _p(ν), C.c_int(len(Tg)), _p(Tg), C.c_int(len(Pg)),
                        _p(Pg), _p(Cg), C.c_double(cut), _p(out), C.c_int(nthreads))
    if nz < 0:
        raise ValueError(f"oracle bake failed ({nz})")
    return out, int(nz)


def table_fit(block):
    """block [nP, nT, nν] -> coefficients A [nν, nP, nT] (A[v, j, i])"""
    block = _f(block)
    nP, nT, nν = block.shape
    A = np.empty((nν, nP, nT))
    lib().orc_table_fit_all(C.c_int64(nν), C.c_int(nT), C.c_int(nP), _p(block), _p(A))
    return A


def gas_nodes(A, Tg, Pg, T, P, Cn, nthreads=1):
    """Σ_gas at nodes: out[node, ν] = C[node]·exp(Φ_ν(T, ln P)); interpolator bounds = grid end points"""
    A = _f(A)
    nν, nP, nT = A.shape
    T, P, Cn = _f(T), _f(P), _f(Cn)
    out = np.zeros((len(T), nν))
    lnP = np.log(_f(Pg))
    lib().orc_gas_nodes(C.c_int64(nν), C.c_int(nT), C.c_int(nP), _p(A), C.c_double(Tg[0]), C.c_double(Tg[-1]),
                        C.c_double(lnP[0]), C.c_double(lnP[-1]), C.c_int64(len(T)), _p(T), _p(P), _p(Cn), _p(out),
                        C.c_int(nthreads))
    return out


def _cia_args(tables):
    g_nnu, g_nT, g_nu, g_T, g_lnk, s_n, s_nu, s_lnk = tables.flat()
    off = lambda n: np.concatenate(([0], np.cumsum(n)[:-1])).astype(np.int64) if len(n) else np.zeros(1, np.int64)
    g_off_nu, g_off_T, g_off_k, s_off = off(g_nnu), off(g_nT), off(g_nnu * g_nT), off(s_n)
    z, zi = np.zeros(1), np.zeros(1, np.int64)
    pad = lambda a, zz: a if len(a) else zz
    keep = [pad(g_nnu, zi), pad(g_nT, zi), g_off_nu, g_off_T, g_off_k, pad(g_nu, z), pad(g_T, z), pad(g_lnk, z),
            pad(s_n, zi), s_off, pad(s_nu, z), pad(s_lnk, z)]
    args = [C.c_int32(len(g_nnu)), _pi(keep[0]), _pi(keep[1]), _pi(keep[2]), _pi(keep[3]), _pi(keep[4]), _p(keep[5]),
            _p(keep[6]), _p(keep[7]), C.c_int32(len(s_n)), _pi(keep[8]), _pi(keep[9]), _p(keep[10]), _p(keep[11]),
            C.c_int32(int(tables.extrapolate)), C.c_int32(int(tables.singles))]
    return args, keep


def cia_nodes(tables, ν, T, P, C1, C2):
    """CIA functor at nodes: out[node, ν] = cia(ν, x, T, P, P·C1, P·C2)"""
    ν, T, P, C1, C2 = _f(ν), _f(T), _f(P), _f(C1), _f(C2)
    out = np.zeros((len(T), len(ν)))
    args, keep = _cia_args(tables)
    lib().orc_cia_nodes(*args, C.c_int64(len(ν)), _p(ν), C.c_int64(len(T)), _p(T), _p(P), _p(C1), _p(C2), _p(out))
    return out


def cia_k(tables, ν, T):
    ν, T = _f(ν), _f(T)
    out = np.zeros(len(ν))
    args, keep = _cia_args(tables)
    lib().orc_cia_k_vec(*args, C.c_int64(len(ν)), _p(ν), _p(T), _p(out))
    return out


def accel_nodes(lnP, lnσ, P):
    lnP, lnσ, P = _f(lnP), _f(lnσ), _f(P)
    nlev, nν = lnσ.shape
    out = np.zeros((len(P), nν))
    lib().orc_accel_nodes(C.c_int64(nν), C.c_int64(nlev), _p(lnP), _p(lnσ), C.c_int64(len(P)), _p(P), _p(out))
    return out


def fluxes(ν, P, nlob, wl, μ, Tlev, σnodes, g, fS, fa, θs, nstream, m, W, nthreads=1, full=True):
    """monochromaticfluxes!(Discretized) + ∫F! + Fnet.  σnodes [nnode, nν]; μ [np-1, nlob] (Julia [nlob, np-1])"""
    ν, P, wl, μ, Tlev, σnodes = _f(ν), _f(P), _f(wl), _f(μ), _f(Tlev), _f(σnodes)
    nν, npl = len(ν), len(P)
    fS = _f(np.zeros(nν) if fS is None else fS)
    fa = _f(np.zeros(nν) if fa is None else fa)
    m, W = _f(m), _f(W)
    τ = np.empty((nν, npl - 1)) if full else None
    Mup = np.empty((nν, npl)) if full else None
    Mdn = np.empty((nν, npl)) if full else None
    Fup, Fdn, Fnet = np.empty(npl), np.empty(npl), np.empty(npl)
    lib().orc_fluxes(C.c_int64(nν), _p(ν), C.c_int64(npl), _p(P), C.c_int(nlob), _p(wl), _p(μ), _p(Tlev), _p(σnodes),
                     C.c_double(g), _p(fS), _p(fa), C.c_double(θs), C.c_int(nstream), _p(m), _p(W),
                     _p(τ) if full else None, _p(Mup) if full else None, _p(Mdn) if full else None,
                     _p(Fup), _p(Fdn), _p(Fnet), C.c_int(nthreads))
    return dict(τ=τ, Mup=Mup, Mdn=Mdn, Fup=Fup, Fdn=Fdn, Fnet=Fnet)


def opticaldepth(P, nlob, wl, μ, σnodes, g, θ):
    P, wl, μ, σnodes = _f(P), _f(wl), _f(μ), _f(σnodes)
    nν = σnodes.shape[1]
    out = np.empty(nν)
    lib().orc_opticaldepth(C.c_int64(nν), C.c_int64(len(P)), _p(P), C.c_int(nlob), _p(wl), _p(μ), _p(σnodes),
                           C.c_double(g), C.c_double(1 / np.cos(θ)), _p(out))
    return out


def planck(ν, T):
    ν, T = np.broadcast_arrays(_f(ν), _f(T))
    ν, T = _f(ν), _f(T)
    out = np.empty(ν.shape)
    lib().orc_planck_vec(C.c_int64(ν.size), _p(ν), _p(T), _p(out))
    return out


W985_MAPS = {0: dict(S1=1.6e4, S2=160.0, S3=107.0, S4=28.5, Y4=6e-14, S5=3.5, Y5=0.026),
             1: dict(S1=3.8e4, S2=256.0, S3=62.0, S4=30.0, Y4=1e-13, S5=2.5, Y5=0.072)}


def set_w985_map(m):
    """region borders of the Faddeyeva restatement: 1 (default) = the 4e-5 design the Algorithm 985 paper describes,
    0 = the 1e-4 design (SURVEY.md 8c's recollection); see the header of orc_faddeyeva985 in oracle.c"""
    lib().orc_set_w985_map(C.c_int(int(m)))


def get_w985_map():
    return int(lib().orc_get_w985_map())


def faddeyeva985(x, y):
    x, y = np.broadcast_arrays(_f(x), _f(y))
    x, y = _f(x), _f(y)
    out = np.empty(x.shape)
    lib().orc_faddeyeva985_vec(C.c_int64(x.size), _p(x), _p(y), _p(out))
    return out


def scalar(name, *args):
    """call a scalar double(double...) oracle function, e.g. scalar('orc_lorentz', ν, νl, S, γ)"""
    fn = getattr(lib(), name)
    return fn(*[a if isinstance(a, (C.c_int, C.c_int64, _dp)) else C.c_double(a) for a in args])
