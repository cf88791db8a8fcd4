"""Prototype (CPU): the whole operator apply y = S x in the rotation -- axial translation -- rotation form, against
the oracle's dense matrix.  Per unordered pair only (d, theta, phi), the axial A, B (m-diagonal) and the real
small-d matrices are needed; both directions (i <- j and j <- i) come from the same data."""
import sys
import numpy as np
from math import factorial as f
sys.path.insert(0, "/root/repo")
from oracle import oracle as O
from tests import util as U

def wd(j, mp, m, b):
    s = 0.0
    for t in range(max(0, m - mp), min(j + m, j - mp) + 1):
        s += (-1) ** (mp - m + t) * np.cos(b / 2) ** (2 * j + m - mp - 2 * t) * np.sin(b / 2) ** (mp - m + 2 * t) / (
            f(j + m - t) * f(t) * f(mp - m + t) * f(j - mp - t))
    return s * np.sqrt(f(j + mp) * f(j - mp) * f(j + m) * f(j - m))

def run(spec, harmonic=1):
    orc = U.oracle_case(spec)
    nMax = spec.nMax
    n = nMax * (nMax + 2)
    idx = [(nn, m) for nn in range(1, nMax + 1) for m in range(nn, -nn - 1, -1)]
    ms = np.array([t[1] for t in idx]); ns = np.array([t[0] for t in idx])
    S = orc.matrix(harmonic)
    k = orc.info()["waveK"] * harmonic
    nobj = len(spec.xyz)
    T = [orc.particle_factors(j, 0 if harmonic == 1 else 1) for j in range(nobj)]
    rng = np.random.RandomState(0)
    x = rng.standard_normal(2 * n * nobj) + 1j * rng.standard_normal(2 * n * nobj)
    y = x.copy()  # identity diagonal blocks
    xyz = U.spherical_roundtrip(spec.xyz)
    stored = 0
    for i in range(nobj):
        for j in range(i + 1, nobj):
            v = xyz[i] - xyz[j]
            d = np.linalg.norm(v); the = np.arccos(v[2] / d); phi = np.arctan2(v[1], v[0])
            Az, Bz = O.coupling([d, 0.0, 0.0], k, nMax, True)      # axial: only m_p == m_q entries are non-zero
            dm = np.zeros((n, n))
            for p, (nn, mp) in enumerate(idx):
                for q, (ll, m) in enumerate(idx):
                    if nn == ll:
                        dm[p, q] = wd(nn, mp, m, the)
            ph = np.exp(1j * ms * phi)
            stored += (np.abs(Az) > 0).sum() // 1 + (np.abs(dm) > 0).sum() // 2  # complex-equivalents (A and B share a pattern: x2 below)
            def apply_T(xa, xb, sgnB, par):
                """[A^T B^T; B^T A^T](+-R) applied to (xa, xb); par = parity signs (-1)^n for the reversed direction."""
                ua, ub = dm.T @ (ph * (par * xa)), dm.T @ (ph * (par * xb))
                va = Az.T @ ua + sgnB * (Bz.T @ ub)
                vb = sgnB * (Bz.T @ ua) + Az.T @ ub
                return par * (np.conj(ph) * (dm @ va)), par * (np.conj(ph) * (dm @ vb))
            one = np.ones(n); par = (-1.0) ** ns
            xa, xb = x[2 * n * j:2 * n * j + n], x[2 * n * j + n:2 * n * (j + 1)]
            wa, wb = apply_T(xa, xb, 1.0, one)                       # block (i, j): Coupling(R_i - R_j)
            y[2 * n * i:2 * n * i + n] += -T[i][:n] * wa
            y[2 * n * i + n:2 * n * (i + 1)] += -T[i][n:] * wb
            xa, xb = x[2 * n * i:2 * n * i + n], x[2 * n * i + n:2 * n * (i + 1)]
            wa, wb = apply_T(xa, xb, -1.0, par)                      # block (j, i): A(-R) = P A P, B(-R) = -P B P
            y[2 * n * j:2 * n * j + n] += -T[j][:n] * wa
            y[2 * n * j + n:2 * n * (j + 1)] += -T[j][n:] * wb
    ref = S @ x
    npairs = nobj * (nobj - 1) // 2
    print(spec.name, "harmonic", harmonic, "relerr vs dense oracle:", U.relerr(y, ref),
          "| complex entries per pair: axial", 2 * (np.abs(Az) > 0).sum(), "+ small-d (real)", (np.abs(dm) > 0).sum(),
          "vs pair form", 2 * n * n)

if __name__ == "__main__":
    run(U.random_cluster(5, 4, seed=3))
    run(U.random_cluster(4, 6, seed=5), harmonic=2)
    run(U.three_au(nMax=3))
    run(U.Spec("lossy_bg", [[0, 0, 0], [260, 40, -90], [-30, 310, 120]], [60, 80, 70], U.fixed(9.0 + 0.4j, 7.0 + 0.9j), 4, 700.0,
               theta_deg=30, phi_deg=20, Eth=0.6, Eph=0.8j, background=(1.7 + 0.0j, 1.0 + 0.0j)))
    run(U.two_si(nMax=6))   # pair on the z axis: theta = 0
