"""Prototype (CPU, numpy + oracle): rotation -- axial translation -- rotation factorisation of the coupling blocks.
A(R), B(R) (Coupling, n x n) = D^-1 [A(d z), B(d z)] D with D block-diagonal Wigner rotations and the axial blocks
diagonal in m.  Determines the conventions numerically against the oracle."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
from oracle import oracle as O

nMax = 5
n = nMax * (nMax + 2)
k = 2 * np.pi / 800e-9 * (1.0 + 0.0j)

def flat(nn, m): return nn * (nn + 1) - m - 1
idx = [(nn, m) for nn in range(1, nMax + 1) for m in range(nn, -nn - 1, -1)]
assert all(flat(*t) == i for i, t in enumerate(idx))

def rot_zyz(alpha, beta):
    ca, sa, cb, sb = np.cos(alpha), np.sin(alpha), np.cos(beta), np.sin(beta)
    Rz = np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]])
    Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    return Rz @ Ry  # maps z to (sin b cos a, sin b sin a, cos b)

def sph(v):
    r = np.linalg.norm(v); return np.arccos(v[2] / r), np.arctan2(v[1], v[0])

def wigner_block(Rot, nn, K=200, seed=0):
    """D with Y_n^m(Rot^-1 r) = sum_m' D[m', m] Y_n^m'(r), columns/rows ordered m = n..-n (flat order)."""
    rng = np.random.RandomState(seed)
    dirs = rng.standard_normal((K, 3)); dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    ms = list(range(nn, -nn - 1, -1))
    Y = np.array([[O.ynm(*sph(d), nn, m) for m in ms] for d in dirs])
    Yr = np.array([[O.ynm(*sph(Rot.T @ d), nn, m) for m in ms] for d in dirs])
    D, res, *_ = np.linalg.lstsq(Y, Yr, rcond=None)
    assert np.abs(Y @ D - Yr).max() < 1e-10
    return D

def big_D(Rot):
    D = np.zeros((n, n), dtype=complex)
    for nn in range(1, nMax + 1):
        i0 = flat(nn, nn)
        D[i0:i0 + 2 * nn + 1, i0:i0 + 2 * nn + 1] = wigner_block(Rot, nn)
    return D

d = 260e-9
the, phi = 1.1, 0.7
A, B = O.coupling([d, the, phi], k, nMax, True)
Az, Bz = O.coupling([d, 0.0, 0.0], k, nMax, True)
# axial blocks couple equal m only
mask = np.array([[idx[p][1] == idx[q][1] for q in range(n)] for p in range(n)])
print("axial off-m leakage", np.abs(Az[~mask]).max(), np.abs(Bz[~mask]).max(), "max", np.abs(Az).max())
Rot = rot_zyz(phi, the)
D = big_D(Rot)
print("D unitary", np.abs(D.conj().T @ D - np.eye(n)).max())
cands = {"D^-1 Az D": np.linalg.inv(D) @ Az @ D, "D Az D^-1": D @ Az @ np.linalg.inv(D), "D^T Az D^-T": D.T @ Az @ np.linalg.inv(D.T),
         "D^-T Az D^T": np.linalg.inv(D.T) @ Az @ D.T, "D* Az D*^-1": D.conj() @ Az @ np.linalg.inv(D.conj()),
         "D^H Az D^-H": D.conj().T @ Az @ np.linalg.inv(D.conj().T)}
for name, M in cands.items():
    print(name, np.abs(M - A).max() / np.abs(A).max())

print("---- B and the real small-d ----")
Dc = D.conj()
print("B:", np.abs(Dc @ Bz @ np.linalg.inv(Dc) - B).max() / np.abs(B).max())
ms = np.array([t[1] for t in idx])
E = np.diag(np.exp(-1j * ms * phi))
dsmall = E @ Dc            # remove exp(i m' alpha) from the rows
print("imag part of E conj(D):", np.abs(dsmall.imag).max())
# standard Wigner small-d (Edmonds / Varshalovich 4.3.1 (2)): explicit sum
from math import factorial as f
def wd(j, mp, m, b):
    s = 0.0
    for t in range(max(0, m - mp), min(j + m, j - mp) + 1):
        s += (-1) ** (mp - m + t) * np.cos(b / 2) ** (2 * j + m - mp - 2 * t) * np.sin(b / 2) ** (mp - m + 2 * t) / (
            f(j + m - t) * f(t) * f(mp - m + t) * f(j - mp - t))
    return s * np.sqrt(f(j + mp) * f(j - mp) * f(j + m) * f(j - m))
for variant in ("d[mp,m](b)", "d[m,mp](b)", "d[mp,m](-b)", "(-1)^(mp-m) d[mp,m](b)"):
    W = np.zeros((n, n))
    for p, (nn, mp) in enumerate(idx):
        for q, (ll, m) in enumerate(idx):
            if nn != ll: continue
            if variant == "d[mp,m](b)": W[p, q] = wd(nn, mp, m, the)
            elif variant == "d[m,mp](b)": W[p, q] = wd(nn, m, mp, the)
            elif variant == "d[mp,m](-b)": W[p, q] = wd(nn, mp, m, -the)
            else: W[p, q] = (-1) ** (mp - m) * wd(nn, mp, m, the)
    print(variant, np.abs(W - dsmall.real).max())
