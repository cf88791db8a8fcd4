"""Line-by-line CPU model of the rotated-axial operator form (data layout + index formulas of
optimet_b200/csrc/ob_rot.cu), used by tests/test_oracle_kats.py to check the factorisation and the layout against the
oracle's full coupling blocks (test infrastructure, like the oracle).  Layout per pair (i < j), R = R_i - R_j = (d, theta, phi):
  ph[m + NM]            = exp(i m phi), m = -NM..NM                                   (2 NM + 1 complex)
  dmat[offD(n) + a (2n+1) + b] = d^n_{m' mu}(theta), m' = n - a, mu = n - b            (sum (2n+1)^2 reals)
  Aax[offX(mu) + (n - n0)(NM - n0 + 1) + (l - n0)] = A(d z)[(n, mu), (l, mu)], mu = 0..NM, n0 = max(mu, 1)   (same Bax)
"""
import sys
import numpy as np
from math import factorial as f
from oracle import oracle as O
from tests import util as U

def flat(nn, m): return nn * (nn + 1) - m - 1

def wd(j, mp, m, b):
    s = 0.0
    for t in range(max(0, m - mp), min(j + m, j - mp) + 1):
        s += (-1) ** (mp - m + t) * np.cos(b / 2) ** (2 * j + m - mp - 2 * t) * np.sin(b / 2) ** (mp - m + 2 * t) / (
            f(j + m - t) * f(t) * f(mp - m + t) * f(j - mp - t))
    return s * np.sqrt(float(f(j + mp) * f(j - mp) * f(j + m) * f(j - m)))

FACT = [float(f(i)) for i in range(2 * 13 + 2)]

def small_d_column(NM, mp, m, beta):
    """d^j_{mp m}(beta) for j = 1..NM by the three-term recurrence in j (the CUDA thread for (mp, m))."""
    out = np.zeros(NM + 1)
    j0 = max(abs(mp), abs(m))
    cb, sb = np.cos(beta / 2), np.sin(beta / 2)
    # seed: the explicit sum has one term at j = j0
    t = max(0, m - mp)
    assert t == min(j0 + m, j0 - mp)
    seed = ((-1) ** (mp - m + t) * np.sqrt(FACT[j0 + mp] * FACT[j0 - mp] * FACT[j0 + m] * FACT[j0 - m]) /
            (FACT[j0 + m - t] * FACT[t] * FACT[mp - m + t] * FACT[j0 - mp - t]) *
            cb ** (2 * j0 + m - mp - 2 * t) * sb ** (mp - m + 2 * t))
    c = np.cos(beta)
    dm1, dcur = 0.0, seed
    if j0 >= 1:
        out[j0] = seed
    for j in range(j0 + 1, NM + 1):
        if mp == 0 and m == 0:
            dn = ((2 * j - 1) * c * dcur - (j - 1) * dm1) / j
        else:
            dn = ((2 * j - 1) * (j * (j - 1) * c - m * mp) * dcur
                  - j * np.sqrt(float(((j - 1) ** 2 - mp * mp) * ((j - 1) ** 2 - m * m))) * dm1) / (
                      (j - 1) * np.sqrt(float((j * j - mp * mp) * (j * j - m * m))))
        dm1, dcur = dcur, dn
        out[j] = dn
    return out

def offD(n): return sum((2 * j + 1) ** 2 for j in range(1, n))
def n0_of(mu): return max(mu, 1)
def offX(NM, mu): return sum((NM - n0_of(u) + 1) ** 2 for u in range(mu))

def build_pair(NM, d, the, phi, k):
    ph = np.exp(1j * np.arange(-NM, NM + 1) * phi)
    dmat = np.zeros(offD(NM + 1))
    for mp in range(-NM, NM + 1):
        for m in range(-NM, NM + 1):
            col = small_d_column(NM, mp, m, the)
            for n in range(max(abs(mp), abs(m), 1), NM + 1):
                dmat[offD(n) + (n - mp) * (2 * n + 1) + (n - m)] = col[n]
    Az, Bz = O.coupling([d, 0.0, 0.0], k, NM, True)
    Aax = np.zeros(offX(NM, NM + 1), dtype=complex); Bax = np.zeros_like(Aax)
    for mu in range(NM + 1):
        n0 = n0_of(mu); w = NM - n0 + 1
        for n in range(n0, NM + 1):
            for l in range(n0, NM + 1):
                Aax[offX(NM, mu) + (n - n0) * w + (l - n0)] = Az[flat(n, mu), flat(l, mu)]
                Bax[offX(NM, mu) + (n - n0) * w + (l - n0)] = Bz[flat(n, mu), flat(l, mu)]
    return ph, dmat, Aax, Bax, Az, Bz

def apply_pair(NM, ph, dmat, Aax, Bax, X, sgnB, parity):
    """X[2][n] -> W[2][n] = [A^T B^T; B^T A^T](+-R) X, phases 1..4 as the kernel runs them."""
    n = NM * (NM + 2)
    par = np.array([(-1.0) ** nn if parity else 1.0 for nn in range(1, NM + 1) for _ in range(2 * nn + 1)])
    T = np.zeros((2, n), dtype=complex); Uv = np.zeros_like(T); V = np.zeros_like(T); W = np.zeros_like(T)
    for v in range(2):
        for nn in range(1, NM + 1):
            for m in range(-nn, nn + 1):
                T[v, flat(nn, m)] = ph[m + NM] * par[flat(nn, m)] * X[v, flat(nn, m)]
    for v in range(2):                                   # phase 1: u = d^T t
        for nn in range(1, NM + 1):
            for mu in range(-nn, nn + 1):
                s = 0
                for mp in range(-nn, nn + 1):
                    s += dmat[offD(nn) + (nn - mp) * (2 * nn + 1) + (nn - mu)] * T[v, flat(nn, mp)]
                Uv[v, flat(nn, mu)] = s
    for v in range(2):                                   # phase 2: axial, A^T u_v + sgnB B^T u_(1-v)
        for nn in range(1, NM + 1):
            for mu in range(-nn, nn + 1):
                am = abs(mu); n0 = n0_of(am); w = NM - n0 + 1
                sB = sgnB * (-1.0 if mu < 0 else 1.0)      # B(d z)[(n,-mu),(l,-mu)] = -B(d z)[(n,mu),(l,mu)]
                s = 0
                for l in range(n0, NM + 1):
                    e = offX(NM, am) + (l - n0) * w + (nn - n0)    # A^T[(n,mu),(l,mu)] = A[(l,mu),(n,mu)]
                    s += Aax[e] * Uv[v, flat(l, mu)] + sB * Bax[e] * Uv[1 - v, flat(l, mu)]
                V[v, flat(nn, mu)] = s
    for v in range(2):                                   # phase 3 + 4: w = d v, conj phase, parity
        for nn in range(1, NM + 1):
            for m in range(-nn, nn + 1):
                s = 0
                for mu in range(-nn, nn + 1):
                    s += dmat[offD(nn) + (nn - m) * (2 * nn + 1) + (nn - mu)] * V[v, flat(nn, mu)]
                W[v, flat(nn, m)] = par[flat(nn, m)] * np.conj(ph[m + NM]) * s
    return W

