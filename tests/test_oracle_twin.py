"""Independent numpy restatements (written separately from oracle/oracle.c, vectorised over ν instead of looping) of
the pieces of the path that carry the most logic: the Algorithm-985 region map, the Discretized depth/sweep/∫F core,
the accelerated absorber and the CIA amagat conversion.  Oracle-C vs numpy twin must agree to ~1e-13
(SURVEY.md section 8c 'oracle self-checks')."""
import numpy as np
import pytest

from conftest import relerr

K = dict(h=6.62607015e-34, c=299792458.0, k=1.38064852e-23, Na=6.02214076e23)


def w985_numpy(x, y, W):
    """region map W (oracle.W985_MAPS[m]), complex arithmetic"""
    x, y = np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float))
    z = x + 1j * y
    s = x * x + y * y
    y2 = y * y
    i = 1j / np.sqrt(np.pi)
    zz = z * z
    t = y - 1j * x
    a = [122.607931777104326, 214.382388694706425, 181.928533092181549, 93.155580458138441, 30.180142196210589,
         5.912626209773153, 0.564189583562615]
    b = [122.607931773875350, 352.730625110963558, 457.334478783897737, 348.703917719495792, 170.354001821091472,
         53.992906912940207, 10.479857114260399]
    num = a[6]
    for k in range(5, -1, -1):
        num = num * t + a[k]
    den = 1.0
    for k in range(6, -1, -1):
        den = den * t + b[k]
    u = t * t
    P = 36183.31 - u * (3321.9905 - u * (1540.787 - u * (219.0313 - u * (35.76683 - u * (1.320522 - u * 0.56419)))))
    Q = 32066.6 - u * (24322.84 - u * (9022.228 - u * (2186.181 - u * (364.2191 - u * (61.57037 - u * (1.841439 - u))))))
    with np.errstate(all="ignore"):
        hum = np.exp(u) - t * P / Q
        out = np.where(s >= W["S1"], i / z,
              np.where(s >= W["S2"], i * z / (zz - 0.5),
              np.where(s >= W["S3"], i * (zz - 1) / (z * (zz - 1.5)),
              np.where((s >= W["S4"]) & (y2 >= W["Y4"]), i * z * (zz - 2.5) / (zz * (zz - 3) + 0.75),
              np.where((s >= W["S5"]) & (y2 < W["Y5"]), hum, num / den)))))
    return out.real


def test_faddeyeva985_twin(orc):
    rng = np.random.default_rng(1)
    x = np.concatenate([10 ** rng.uniform(-6, 4, 20000), [0.0, 0.0]])
    y = np.concatenate([10 ** rng.uniform(-12, 3, 20000), [0.5, 3.0]])
    try:
        for m in (1, 0):
            orc.set_w985_map(m)
            a, b = orc.faddeyeva985(x, y), w985_numpy(x, y, orc.W985_MAPS[m])
            assert np.max(np.abs(a - b) / np.abs(b)) < 5e-12, m
    finally:
        orc.set_w985_map(1)


def fluxes_numpy(ν, P, nlob, wl, μn, Tlev, σ, g, fS, fa, θs, m, W):
    """𝒹depth! + 𝒹monoflux! + ∫F! (core/discretized.jl:136-177,249-326; core/shared.jl:125-137), vectorised over ν"""
    L = len(P) - 1
    Cg = 1e-4 * K["Na"] / g
    νm = 100.0 * ν
    B = np.array([100.0 * (2 * K["h"] * K["c"] ** 2 * νm ** 3) / (np.exp(K["h"] * K["c"] * νm / (K["k"] * T)) - 1.0) for T in Tlev])
    τ = np.empty((L, len(ν)))
    for i in range(L):
        dP = P[i + 1] - P[i]
        acc = np.zeros(len(ν))
        for n in range(nlob):
            node = n + (nlob - 1) * i
            # the layer's first node reuses β of the previous layer's end node (discretized.jl:150,172), i.e. its μ
            μ = μn[i - 1, nlob - 1] if (n == 0 and i > 0) else μn[i, n]
            acc += (dP * wl[n]) * (Cg * (σ[node] / μ))
        τ[i] = np.maximum(acc, 1e-6)
    Mup, Mdn = np.zeros((L + 1, len(ν))), np.zeros((L + 1, len(ν)))
    lp = lambda B1, B2, tt, t: B2 * (1 - t) - (B1 - B2) * t + (1 - t) * (B1 - B2) / tt
    for k in range(len(m)):
        I = np.zeros(len(ν))
        for i in range(L):
            tt = τ[i] * m[k]
            t = np.exp(-tt)
            I = I * t + lp(B[i], B[i + 1], tt, t)
            Mdn[i + 1] += W[k] * I
    c = np.cos(θs)
    Mdn[0] += c * fS
    Ms = Mdn[0].copy()
    for i in range(L):
        Ms = Ms * np.exp(-τ[i] / c)
        Mdn[i + 1] += Ms
    Is = Mdn[-1] * fa / np.pi + B[-1]
    Mup[-1] = Is * np.pi
    for k in range(len(m)):
        I = Is.copy()
        for i in range(L - 1, -1, -1):
            tt = τ[i] * m[k]
            t = np.exp(-tt)
            I = I * t + lp(B[i + 1], B[i], tt, t)
            Mup[i] += W[k] * I
    dν = np.diff(ν)
    F = lambda M: np.array([np.sum(dν * (M[i, :-1] + M[i, 1:]) / 2) for i in range(L + 1)])
    return dict(τ=τ.T, Mup=Mup.T, Mdn=Mdn.T, Fup=F(Mup), Fdn=F(Mdn))


@pytest.mark.parametrize("nlob,nstream", [(2, 5), (4, 3)])
def test_flux_core_twin(orc, cs, nlob, nstream):
    rng = np.random.default_rng(3)
    ν = np.sort(rng.uniform(10, 2500, 157))
    P = cs.pressuregrid(30.0, 9e4, 12)
    L = len(P) - 1
    nn = L * (nlob - 1) + 1
    σ = 10 ** rng.uniform(-27, -20, (nn, len(ν)))
    μn = rng.uniform(0.02, 0.04, (L, nlob))
    Tlev = np.linspace(180, 290, len(P))
    fS, fa = rng.uniform(0, 2, len(ν)), rng.uniform(0, 0.6, len(ν))
    m, W = cs.streamnodes(nstream)
    x, wl = cs.lobattonodes(nlob)
    a = orc.fluxes(ν, P, nlob, wl, μn, Tlev, σ, 9.8, fS, fa, 0.7, nstream, m, W)
    b = fluxes_numpy(ν, P, nlob, wl, μn, Tlev, σ, 9.8, fS, fa, 0.7, m, W)
    for key in ("τ", "Mup", "Mdn", "Fup", "Fdn"):
        assert relerr(a[key], b[key], 1e-300) < 1e-12, key
    # opticaldepth: Σ_i τ_i·m without the floor (core/discretized.jl:92-134)
    od = orc.opticaldepth(P, nlob, wl, μn, σ, 9.8, 0.4)
    Cg = 1e-4 * K["Na"] / 9.8
    ref = np.zeros(len(ν))
    for i in range(L):
        dP = P[i + 1] - P[i]
        ref += sum((dP * wl[n]) * (Cg * σ[n + (nlob - 1) * i] / (μn[i - 1, nlob - 1] if (n == 0 and i > 0) else μn[i, n]))
                   for n in range(nlob)) / np.cos(0.4)
    assert relerr(od, ref) < 1e-13


def test_accel_and_cia_twin(orc, cs):
    rng = np.random.default_rng(5)
    lnP = np.linspace(2, 11, 9) + rng.uniform(-0.3, 0.3, 9)
    lnσ = rng.uniform(-60, -45, (9, 23))
    Pq = np.exp(rng.uniform(1.5, 11.5, 17))       # includes extrapolation off both ends
    got = orc.accel_nodes(lnP, lnσ, Pq)
    i = np.clip(np.searchsorted(lnP, np.log(Pq), side="right") - 1, 0, 7)
    w = (np.log(Pq) - lnP[i]) / (lnP[i + 1] - lnP[i])
    ref = np.exp(lnσ[i] + w[:, None] * (lnσ[i + 1] - lnσ[i]))
    assert relerr(got, ref) < 1e-12
    k, T, Pa, P1, P2 = 2.5e-44, 233.0, 1.3e5, 1e5, 2e3
    ref = k * 7.21879268e38 * (P1 / 101325 * 273.15 / T) * (P2 / 101325 * 273.15 / T) / (1e-6 * Pa / (K["k"] * T))
    assert abs(orc.scalar("orc_cia_sigma", k, T, Pa, P1, P2) / ref - 1) < 1e-14
