"""CPU model of an axial-only assembly for the rotated-axial operator form: with theta = 0 the scalar translation
coefficients beta(n, m, l, k) vanish unless k = m, and the reference's recursion (TranslationAdditionCoefficients.cpp:
102-124) closes on those entries: O(nMax^3) work per pair instead of the O(nMax^4) of the full block.  This is the
specification of the next assembly kernel; tests/test_oracle_kats.py checks it against the
oracle's Coupling at theta = 0.  Test infrastructure, like the oracle."""
import numpy as np

from oracle import oracle as O


def a_plus(n, m):
    return 0.0 if (n < 0 or abs(m) > n) else -np.sqrt((n + m + 1) * (n - m + 1) / ((2 * n + 1) * (2 * n + 3)))


def a_minus(n, m):
    return 0.0 if (n < 0 or abs(m) > n) else np.sqrt((n + m) * (n - m) / ((2 * n + 1) * (2 * n - 1)))


def b_plus(n, m):
    return 0.0 if (n < 0 or abs(m) > n) else np.sqrt((n + m + 2) * (n + m + 1) / ((2 * n + 1) * (2 * n + 3)))


def b_minus(n, m):
    return 0.0 if (n < 0 or abs(m) > n) else np.sqrt((n - m) * (n - m - 1) / ((2 * n + 1) * (2 * n - 1)))


def axial_beta(NM, z, regular=False):
    """beta[m][n][l] = beta(n, m, l, m) for translation d z (z = k d), 0 <= m <= n <= NM, l <= 2 NM - n."""
    L = 2 * NM
    zl, _ = O.bessel(0 if regular else 1, z, L + 1)
    beta = np.zeros((NM + 2, NM + 1, L + 3), dtype=complex)  # one guard column / chain for the reads at l + 1, m + 1

    def g(m, n, l):
        return beta[m, n, l] if (0 <= l <= L + 1 and 0 <= m <= n) else 0.0

    for l in range(L + 1):  # seeds (n = m = 0): sqrt(4 pi) (-1)^l Y_l0(0) z_l = (-1)^l sqrt(2l + 1) z_l
        beta[0, 0, l] = (-1) ** l * np.sqrt(2 * l + 1.0) * zl[l]
    for n in range(1, NM + 1):
        for m in range(n + 1):
            for l in range(m, L - n + 1):
                if m == n:   # sectorial step (:113-117), k = m
                    v = (g(n - 1, n - 1, l - 1) * b_plus(l - 1, n - 1) + g(n - 1, n - 1, l + 1) * b_minus(l + 1, n - 1)) / b_plus(n - 1, n - 1)
                else:        # general step (:119-124)
                    prev2 = beta[m, n - 2, l] if n - 2 >= m else 0.0
                    v = (-prev2 * a_minus(n - 1, m) + g(m, n - 1, l - 1) * a_plus(l - 1, m) + g(m, n - 1, l + 1) * a_minus(l + 1, m)) / a_plus(n - 1, m)
                beta[m, n, l] = v
    return beta


def axial_AB(NM, z, regular=False):
    """A[(n, mu), (l, mu)], B[...] for mu = 0..NM from the axial betas (Coupling.cpp:30-51 with k = m = mu):
    returns dicts keyed (mu, n, l)."""
    beta = axial_beta(NM, z, regular)

    def ta(n, m, l):  # beta(n, m, l, m); m < 0 by the phase-free symmetry beta(n,-m,l,-m) = beta(n,m,l,m)
        m = abs(m)
        if m > n or m > l or l < 0:
            return 0.0
        return beta[m, n, l]

    A, B = {}, {}
    for mu in range(NM + 1):
        for n in range(max(mu, 1), NM + 1):
            for l in range(max(mu, 1), NM + 1):
                m = k = mu
                f = 0.5 / np.sqrt(l * (l + 1) * n * (n + 1))
                c0 = 2 * k * m
                c1 = np.sqrt((n - m) * (n + m + 1) * (l - k) * (l + k + 1))
                c2 = np.sqrt((n + m) * (n - m + 1) * (l + k) * (l - k + 1))
                A[(mu, n, l)] = f * (c0 * ta(n, m, l) + c1 * ta(n, m + 1, l) + c2 * ta(n, m - 1, l))
                fb = -0.5j * np.sqrt((2 * l + 1) / ((2 * l - 1) * l * (l + 1) * n * (n + 1)))
                d0 = 2 * m * np.sqrt((l - k) * (l + k))
                d1 = np.sqrt((n - m) * (n + m + 1) * (l - k) * (l - k - 1))
                d2 = np.sqrt((n + m) * (n - m + 1) * (l + k) * (l + k - 1))
                B[(mu, n, l)] = fb * (d0 * ta(n, m, l - 1) + d1 * ta(n, m + 1, l - 1) - d2 * ta(n, m - 1, l - 1))
    return A, B
