#!/usr/bin/env python
"""Numerical prototype (numpy, CPU) of the next step for the far-field expansion of K2:
cluster moments + moment-to-local translation, so that the per-tile cost of the far field no longer scales with the
number of lines.

  far-wing line j at nul_j, half width g_j, strength K_j = S g/pi:     K_j / ((nu - nul_j)^2 + g_j^2) = Im[ s_j / (nu - z_j) ],
  z_j = nul_j + i g_j,  s_j = K_j / g_j  (real).

  P2M   cluster C with centre zc (real):  Im sum_j s_j/(nu - z_j) = sum_{m>=1} mu_m / (nu - zc)^(m+1),
        mu_m = Im sum_j s_j (z_j - zc)^m = sum_j K_j C_m(u_j, g_j),   C_1 = 1,  A_1 = u,
        C_{m+1} = u C_m + A_m,  A_{m+1} = u A_m - g^2 C_m      (u = nul_j - zc; only g^2 and K are needed: both are in the record)
  M2L   tile with centre c, half width h, t = (nu - c)/h,  R = zc - c:
        1/(nu - zc)^(m+1) = (-1)^(m+1) R^-(m+1) sum_k C(m+k, k) (h t / R)^k
        => a_k = -(h/R)^k / R * S_k,   S_k = sum_m C(m+k, k) beta_m,   beta_m = mu_m (-1/R)^m
        and S_k is the first element after k+1 suffix-sum passes over beta  (no binomials, p^2/2 additions).

The script checks the identities and the truncation error against the direct sum for a geometry like C2's
(128-point tiles of 1.27 cm^-1, 83 lines per cm^-1 and gas, clusters of 32 lines) and prints the cost model.
"""
import numpy as np

rng = np.random.default_rng(7)


def moments(K, u, g2, p):
    """mu_1..mu_p of one cluster (P2M recurrence above)"""
    C = np.ones_like(u)
    A = u.copy()
    mu = np.empty(p)
    for m in range(1, p + 1):
        mu[m - 1] = np.sum(K * C)
        C, A = u * C + A, u * A - g2 * C
    return mu


def m2l(mu, R, h, p):
    """local coefficients a_0..a_{p-1} (in t = (nu - c)/h) of one cluster's moments, by suffix sums"""
    m = np.arange(1, len(mu) + 1)
    beta = mu * (-1.0 / R) ** m
    a = np.empty(p)
    c = np.concatenate(([0.0], beta))                 # index = m (mu_0 = 0): element 0 after k+1 passes is S_k
    for k in range(p):
        n = len(c) - k if len(c) - k > 0 else 0      # triangular truncation: m + k <= p
        c[:n] = np.cumsum(c[:n][::-1])[::-1]          # suffix sums over the live part
        a[k] = -(h / R) ** k / R * c[0]
    return a


def main():
    h, dens, cut = 0.635, 83.0, 25.0
    c = 1500.0
    t = np.linspace(-1, 1, 128)
    nu = c + h * t
    n = int(2 * cut * dens)
    nul = np.sort(rng.uniform(c - cut + h, c + cut - h, n))
    g = rng.uniform(1e-5, 0.13, n)
    K = 10 ** rng.uniform(-30, -19, n) * g / np.pi
    direct = (K[:, None] / ((nu[None, :] - nul[:, None]) ** 2 + g[:, None] ** 2))
    for theta, p in ((5.0, 18), (8.0, 14), (16.0, 11), (8.0, 12), (6.0, 16)):      # the kernel uses (5, 18)
        a = np.zeros(p)
        used = np.zeros(n, bool)
        ncl = 0
        for s in range(0, n - 31, 32):
            sl = slice(s, s + 32)
            zc = 0.5 * (nul[s] + nul[s + 31])
            rho = np.hypot(0.5 * (nul[s + 31] - nul[s]), g[sl].max())
            R = zc - c
            if abs(R) < theta * (h + rho):
                continue                                   # too close: stays with the per-line expansion / direct sum
            a += m2l(moments(K[sl], nul[sl] - zc, g[sl] ** 2, p), R, h, p)
            used[sl] = True
            ncl += 1
        approx = np.polynomial.polynomial.polyval(t, a)
        ref = direct[used].sum(0)
        err = np.max(np.abs(approx - ref) / ref)
        ops_m2l = ncl * (p * p / 2 + 3 * p)
        ops_p2l = used.sum() * (3 * 20 + 10)
        print(f"theta {theta:4.1f} p {p:2d}: {ncl:3d} clusters ({used.sum()} of {n} lines), max rel err {err:.2e}, "
              f"M2L ~{ops_m2l:.0f} FP64 ops per (tile, level) vs per-line expansion ~{ops_p2l:.0f}")
        assert err < 1e-9


if __name__ == "__main__":
    main()
