"""CPU model of the symmetry-reduced rotated-axial operator form (record layout v2 of optimet_b200/csrc/ob_rot.cu),
line by line as the kernels index it; tests/test_oracle_kats.py holds it to the oracle's full coupling blocks.  Test
infrastructure, like the oracle.

With R = R_i - R_j = (d, theta, phi):  [A^T B^T; B^T A^T](R) = P^* D Ax D^T P,  P = diag(exp(i m phi)),
D = blockdiag_n d^n(theta) (Wigner small-d), Ax the axial (theta = 0) operator, diagonal in mu with
A(-mu) = A(mu), B(-mu) = -B(mu)   (tests/rot_model.py, pinned in round 1).  Two more symmetries halve the work:

 * d^n commutes with the flip F e_m = (-1)^m e_{-m}, so in the orthonormal basis
       s_0 = e_0,  s_a = (e_a + (-1)^a e_-a)/sqrt2,  a_a = (e_a - (-1)^a e_-a)/sqrt2      (a = 1..n)
   it splits into Ds ((n+1) x (n+1)) and Da (n x n), with Ds[a',a] = (-1)^(a'-a) Ds[a,a'] (same for Da);
 * in that basis A keeps the class (s/a) and B swaps it, so (TE_s, TM_a) and (TM_s, TE_a) are closed and their
   +- combinations diagonalise [A B; B A]:  (V_TE,s +- V_TM,a) = (A^T +- B^T)(U_TE,s +- U_TM,a), same for (TM_s, TE_a).

Record per pair (i < j):
  ph[m + NM] = exp(i m phi)                                                     (2 NM + 1 complex)
  Cp[offX(a) + (l - n0) w + (n - n0)] = A[(l,a),(n,a)] + B[(l,a),(n,a)],  a = 0..NM, n0 = max(a,1), w = NM - n0 + 1
  Cm[offX(a) - NM^2 + ...]            = A - B,                            a = 1..NM   (a = 0: B = 0, Cm = Cp)
  Ds[offDs(n) + a' (n + 1) + a],  Da[offDa(n) + (a' - 1) n + (a - 1)]      (reals)
F-vector layout (shared memory): degree n at offF(n) = (n - 1)(n + 2): s[0..n] then a[0..n] with a[0] = 0.
The reversed direction (block (j,i)) is the same operator applied to (s x_TE, -s x_TM), s = (-1)^deg, with the same
signs on the result (A(-R) = (-1)^(n+l) A(R), B(-R) = (-1)^(n+l+1) B(R)).
"""
import numpy as np

from oracle import oracle as O
from tests import rot_model as R1

SQ2 = np.sqrt(2.0)


def flat(nn, m):
    return nn * (nn + 1) - m - 1


def n0_of(a):
    return max(a, 1)


def offX(NM, a):
    return sum((NM - n0_of(u) + 1) ** 2 for u in range(a))


def offDs(n):
    return sum((j + 1) ** 2 for j in range(1, n))


def offDa(n):
    return sum(j * j for j in range(1, n))


def offF(n):
    return (n - 1) * (n + 2)


def lenF(NM):
    return NM * (NM + 3)


def small_d(NM, beta):
    """d[n][m' + n][m + n] = d^n_{m' m}(beta) by the recurrence of the device (tests/rot_model.small_d_column)."""
    d = [None] + [np.zeros((2 * n + 1, 2 * n + 1)) for n in range(1, NM + 1)]
    for mp in range(-NM, NM + 1):
        for m in range(-NM, NM + 1):
            col = R1.small_d_column(NM, mp, m, beta)
            for n in range(max(abs(mp), abs(m), 1), NM + 1):
                d[n][mp + n, m + n] = col[n]
    return d


def build_pair(NM, dist, the, phi, k):
    ph = np.exp(1j * np.arange(-NM, NM + 1) * phi)
    d = small_d(NM, the)
    Ds = np.zeros(offDs(NM + 1))
    Da = np.zeros(offDa(NM + 1))
    for n in range(1, NM + 1):
        dn = d[n]
        for ap in range(n + 1):
            for a in range(n + 1):
                if ap == 0 and a == 0:
                    v = dn[n, n]
                elif a == 0:
                    v = SQ2 * dn[ap + n, n]
                elif ap == 0:
                    v = SQ2 * dn[n, a + n]
                else:
                    v = dn[ap + n, a + n] + (-1) ** a * dn[ap + n, -a + n]
                    Da[offDa(n) + (ap - 1) * n + (a - 1)] = dn[ap + n, a + n] - (-1) ** a * dn[ap + n, -a + n]
                Ds[offDs(n) + ap * (n + 1) + a] = v
    Az, Bz = O.coupling([dist, 0.0, 0.0], k, NM, True)
    X = offX(NM, NM + 1)
    Cp = np.zeros(X, dtype=complex)
    Cm = np.zeros(X - NM * NM, dtype=complex)
    for a in range(NM + 1):
        n0 = n0_of(a)
        w = NM - n0 + 1
        for n in range(n0, NM + 1):
            for l in range(n0, NM + 1):
                e = offX(NM, a) + (l - n0) * w + (n - n0)
                av, bv = Az[flat(l, a), flat(n, a)], Bz[flat(l, a), flat(n, a)]
                Cp[e] = av + bv
                if a >= 1:
                    Cm[e - NM * NM] = av - bv
    return ph, Cp, Cm, Ds, Da


def apply_pair(NM, rec, X, reverse):
    """X[2][n] (TE, TM) -> [A^T B^T; B^T A^T](+-R) X, phase by phase as k_matvec_rot runs them."""
    ph, Cp, Cm, Ds, Da = rec
    n_h = NM * (NM + 2)
    LF = lenF(NM)
    # phase 0: signs of the reversed direction, phases, flip-symmetric / antisymmetric combinations
    T = np.zeros((2, LF), dtype=complex)
    for v in range(2):
        for n in range(1, NM + 1):
            sg = (((-1.0) ** n) if reverse else 1.0) * ((-1.0) if (reverse and v == 1) else 1.0)
            tp = lambda m: sg * ph[m + NM] * X[v, flat(n, m)]
            T[v, offF(n)] = tp(0)
            for a in range(1, n + 1):
                T[v, offF(n) + a] = (tp(a) + (-1) ** a * tp(-a)) / SQ2
                T[v, offF(n) + n + 1 + a] = (tp(a) - (-1) ** a * tp(-a)) / SQ2
    # phase 1: u = D^T t per class, then the +- channel combinations p = (TE_s +- TM_a), r = (TM_s +- TE_a)
    Uv = np.zeros_like(T)
    for v in range(2):
        for n in range(1, NM + 1):
            for a in range(n + 1):
                Uv[v, offF(n) + a] = sum(Ds[offDs(n) + ap * (n + 1) + a] * T[v, offF(n) + ap] for ap in range(n + 1))
            for a in range(1, n + 1):
                Uv[v, offF(n) + n + 1 + a] = sum(Da[offDa(n) + (ap - 1) * n + (a - 1)] * T[v, offF(n) + n + 1 + ap]
                                                 for ap in range(1, n + 1))
    P = np.zeros((4, LF // 2 + NM), dtype=complex)  # channels p+, p-, r+, r- indexed [n][a] at offC(n) + a
    offC = lambda n: offF(n) // 2
    for n in range(1, NM + 1):
        for a in range(n + 1):
            tes, tma = Uv[0, offF(n) + a], Uv[1, offF(n) + n + 1 + a]
            tms, tea = Uv[1, offF(n) + a], Uv[0, offF(n) + n + 1 + a]
            P[0, offC(n) + a], P[1, offC(n) + a] = tes + tma, tes - tma
            P[2, offC(n) + a], P[3, offC(n) + a] = tms + tea, tms - tea
    # phase 2: axial, q+ = Cp p+, q- = Cm p- (a = 0: Cm = Cp), same for r; back to the class vectors with the sign
    # (-1)^a of the transposed small-d read folded in
    V = np.zeros_like(T)
    for n in range(1, NM + 1):
        for a in range(n + 1):
            n0 = n0_of(a)
            w = NM - n0 + 1
            q = np.zeros(4, dtype=complex)
            for l in range(n0, NM + 1):
                e = offX(NM, a) + (l - n0) * w + (n - n0)
                cp = Cp[e]
                cm = Cp[e] if a == 0 else Cm[e - NM * NM]
                q += np.array([cp * P[0, offC(l) + a], cm * P[1, offC(l) + a], cp * P[2, offC(l) + a],
                               cm * P[3, offC(l) + a]])
            sa = (-1.0) ** a
            V[0, offF(n) + a] = sa * 0.5 * (q[0] + q[1])              # TE_s
            V[1, offF(n) + n + 1 + a] = sa * 0.5 * (q[0] - q[1])      # TM_a (zero at a = 0)
            V[1, offF(n) + a] = sa * 0.5 * (q[2] + q[3])              # TM_s
            V[0, offF(n) + n + 1 + a] = sa * 0.5 * (q[2] - q[3])      # TE_a (zero at a = 0)
    # phase 3: w = D v through the same row-major arrays: D[a',a] = (-1)^(a'-a) D[a,a']; phase 4: back to m, conj phase
    W = np.zeros((2, n_h), dtype=complex)
    for v in range(2):
        for n in range(1, NM + 1):
            sg = (((-1.0) ** n) if reverse else 1.0) * ((-1.0) if (reverse and v == 1) else 1.0)
            for ap in range(n + 1):
                ws = (-1.0) ** ap * sum(Ds[offDs(n) + a * (n + 1) + ap] * V[v, offF(n) + a] for a in range(n + 1))
                if ap == 0:
                    W[v, flat(n, 0)] = sg * np.conj(ph[NM]) * ws
                    continue
                wa = (-1.0) ** ap * sum(Da[offDa(n) + (a - 1) * n + (ap - 1)] * V[v, offF(n) + n + 1 + a]
                                        for a in range(1, n + 1))
                W[v, flat(n, ap)] = sg * np.conj(ph[ap + NM]) * (ws + wa) / SQ2
                W[v, flat(n, -ap)] = sg * np.conj(ph[-ap + NM]) * (-1.0) ** ap * (ws - wa) / SQ2
    return W
