"""CPU known-answer tests that pin the oracle (oracle/optimet_oracle.cpp) -- no GPU needed.

The reference ships no tests, golden vectors or fixtures for this path (the regression
harness is orphaned, the test-data submodule is absent) and cannot be built here (Eigen/Boost/GSL/HDF5 are
missing), so the oracle is pinned by
  * the reference's OWN Bessel/Hankel code: srcAna/amos.c + fortran.c compiled unmodified into
    oracle/_ref/libamos_ref.so (oracle/Makefile) -- values and the -conj(k) branch behaviour;
  * independent libraries for the third-party arithmetic the reference delegates (scipy for Bessel and
    Y_nm, sympy exact rationals for the Wigner 3j/6j/9j symbols of GSL);
  * closed-form physics identities that exercise the recursion, the sign/index conventions, the T-matrix
    and the cross-section prefactors at once (translation group, plane-wave phase, single-sphere Mie).
"""
import numpy as np
import pytest
import scipy.special as sp

from oracle import oracle as O
from tests import util as U

K800 = 2 * np.pi / 800e-9


# --------------------------------------------------------------------------------------------------
# special functions
# --------------------------------------------------------------------------------------------------
ZS = [0.3, 1.3 + 0.2j, 7.5, 3.0 + 2.0j, 0.05 + 0.01j, 25.0 + 0.5j, 1e-3]


@pytest.mark.parametrize("z", ZS)
def test_bessel_vs_scipy(z):
    # Bessel.h:58-143: data[i] = z_i, ddata[i] = -z_{i+1} + (i/z) z_i
    nmax = 26
    n = np.arange(nmax + 1)
    j, dj = O.bessel(0, z, nmax)
    h, dh = O.bessel(1, z, nmax)
    jr = sp.spherical_jn(n, z)
    hr = jr + 1j * sp.spherical_yn(n, z)
    assert np.max(np.abs(j - jr) / np.abs(jr)) < 5e-13
    assert np.max(np.abs(h - hr) / np.abs(hr)) < 5e-13
    djr = sp.spherical_jn(n, z, derivative=True)
    dhr = djr + 1j * sp.spherical_yn(n, z, derivative=True)
    assert np.max(np.abs(dj - djr) / np.abs(djr)) < 1e-10
    assert np.max(np.abs(dh - dhr) / np.abs(dhr)) < 1e-10


@pytest.mark.skipif(not O.have_amos(), reason="oracle/_ref/libamos_ref.so not built (reference tree absent)")
@pytest.mark.parametrize("z", ZS + [-1.3 + 0.2j, -2.0 + 0.0j, -0.7 + 1e-3j])
def test_bessel_restatement_vs_reference_amos(z):
    """Own restatement (backend 0) against the reference's AMOS (backend 1), including arguments on/near the
    negative real axis where the irregular 'negative' TA instance evaluates h1(-conj(k) r)
    (TranslationAdditionCoefficients.h:72-74)."""
    nmax = 24
    try:
        O.set_bessel_backend(1)
        ja, dja = O.bessel(0, z, nmax)
        ha, dha = O.bessel(1, z, nmax)
    finally:
        O.set_bessel_backend(0)
    jo, djo = O.bessel(0, z, nmax)
    ho, dho = O.bessel(1, z, nmax)
    assert np.max(np.abs(jo - ja) / np.abs(ja)) < 2e-12
    assert np.max(np.abs(ho - ha) / np.abs(ha)) < 2e-12
    assert np.max(np.abs(dho - dha) / np.abs(dha)) < 1e-10


@pytest.mark.skipif(not O.have_amos(), reason="oracle/_ref/libamos_ref.so not built")
def test_hankel_minus_conj_identity_on_reference_amos():
    # h_l(-conj(k) r) as AMOS + sqrt(pi/2z) produce it equals (-1)^l conj(h_l(k r))
    try:
        O.set_bessel_backend(1)
        for z in (1.7 + 0.0j, 2.4 + 0.3j, 0.9 + 0.05j):
            h, _ = O.bessel(1, z, 20)
            hm, _ = O.bessel(1, -np.conj(z), 20)
            sign = (-1.0) ** np.arange(21)
            assert np.max(np.abs(hm - sign * np.conj(h)) / np.abs(h)) < 1e-12
    finally:
        O.set_bessel_backend(0)


def test_ynm_condon_shortley():
    # boost::math::spherical_harmonic(n, m, theta, phi) (TranslationAdditionCoefficients.cpp:64-68)
    rng = np.random.RandomState(0)
    for _ in range(40):
        n = rng.randint(0, 25)
        m = rng.randint(-n, n + 1)
        the, phi = rng.uniform(0, np.pi), rng.uniform(-np.pi, np.pi)
        ref = sp.sph_harm_y(n, m, the, phi)
        assert abs(O.ynm(the, phi, n, m) - ref) < 1e-12 * max(1.0, abs(ref))


def test_wigner_symbols_exact():
    # gsl_sf_coupling_3j/6j/9j (Symbol.cpp:28-48) against sympy's exact rationals
    from sympy.physics.wigner import wigner_3j, wigner_6j, wigner_9j
    rng = np.random.RandomState(1)
    n3 = n6 = n9 = 0
    while n3 < 40:
        j1, j2 = rng.randint(0, 9, 2)
        j3 = rng.randint(abs(j1 - j2), j1 + j2 + 1)
        m1, m2 = rng.randint(-j1, j1 + 1), rng.randint(-j2, j2 + 1)
        m3 = -m1 - m2
        if abs(m3) > j3:
            continue
        ref = float(wigner_3j(j1, j2, j3, m1, m2, m3))
        assert abs(O.wigner(3, [j1, j2, j3, m1, m2, m3]) - ref) < 1e-13
        n3 += 1
    while n6 < 25:
        js = list(rng.randint(0, 7, 6))
        try:
            ref = float(wigner_6j(*js))
        except ValueError:
            continue
        assert abs(O.wigner(6, js) - ref) < 1e-13
        n6 += 1
    while n9 < 12:
        js = list(rng.randint(0, 4, 9))
        try:
            ref = float(wigner_9j(*js, prec=None))
        except ValueError:
            continue
        assert abs(O.wigner(9, js) - ref) < 1e-13
        n9 += 1
    # selection-rule zeros
    assert O.wigner(3, [1, 1, 1, 0, 0, 0]) == 0.0
    assert O.wigner(3, [2, 2, 5, 0, 0, 0]) == 0.0


# --------------------------------------------------------------------------------------------------
# translation-addition coefficients / Coupling
# --------------------------------------------------------------------------------------------------
def test_ta_seed_values():
    # TranslationAdditionCoefficients.cpp:102-111: beta(0,0,l,k) = sqrt(4pi) (-1)^(l+k) Y_{l,-k} z_l(kr)
    R = [310e-9, 0.8, -2.1]
    for regular in (True, False):
        for l, k in ((0, 0), (1, 1), (3, -2), (6, 5), (9, 0)):
            z = K800 * R[0]
            zl = sp.spherical_jn(l, z) if regular else sp.spherical_jn(l, z) + 1j * sp.spherical_yn(l, z)
            ref = np.sqrt(4 * np.pi) * (-1.0) ** (l + k) * sp.sph_harm_y(l, -k, R[1], R[2]) * zl
            got = O.ta(R, K800, regular, 0, 0, l, k)
            assert abs(got - ref) < 1e-12 * abs(ref)


def _full(A, B):
    return np.block([[A, B], [B, A]])


def test_translation_group_identity():
    """[A B;B A](R) [A B;B A](-R) = I on the low orders (regular coefficients; truncation-limited):
    pins the recursion, the m<0 branch, the signs and the flat indexing at once."""
    nMax = 8
    n = nMax * (nMax + 2)
    R = np.array([60e-9, -45e-9, 80e-9])
    r = np.linalg.norm(R)
    sph = [r, np.arccos(R[2] / r), np.arctan2(R[1], R[0])]
    msph = [r, np.arccos(-R[2] / r), np.arctan2(-R[1], -R[0])]
    A1, B1 = O.coupling(sph, K800, nMax, False)
    A2, B2 = O.coupling(msph, K800, nMax, False)
    P = _full(A1, B1) @ _full(A2, B2)
    low = [p for p in range(n) if p < 15] + [n + p for p in range(15)]  # orders <= 3
    err = np.max(np.abs(P[np.ix_(low, low)] - np.eye(len(low))))
    assert err < 1e-8


@pytest.mark.parametrize("regular_flag", [True, False])
def test_inversion_parity(regular_flag):
    # A(-R) = (-1)^(n+l) A(R), B(-R) = (-1)^(n+l+1) B(R)   (capacity note)
    nMax = 6
    R = np.array([160e-9, 90e-9, -210e-9])
    r = np.linalg.norm(R)
    sph = [r, np.arccos(R[2] / r), np.arctan2(R[1], R[0])]
    msph = [r, np.arccos(-R[2] / r), np.arctan2(-R[1], -R[0])]
    A1, B1 = O.coupling(sph, K800 * (1.2 + 0.05j), nMax, regular_flag)
    A2, B2 = O.coupling(msph, K800 * (1.2 + 0.05j), nMax, regular_flag)
    deg = np.array([int(np.sqrt(p + 1)) for p in range(nMax * (nMax + 2))])
    sg = (-1.0) ** (deg[:, None] + deg[None, :])
    assert U.relerr(A2, sg * A1) < 1e-12
    assert U.relerr(B2, -sg * B1) < 1e-12


def test_zero_translation_is_identity():
    A, B = O.coupling([0.0, 0.0, 0.0], K800, 4, False)  # Coupling.cpp:82-84
    assert np.array_equal(A, np.eye(24)) and not B.any()


def test_plane_wave_phase_identity():
    """Excitation::getIncLocal (Excitation.cpp:79-129): translating the plane-wave coefficients to R_j equals
    multiplying them by exp(i k.R_j), up to truncation."""
    nMax = 12
    spec = U.Spec("pw", [[0, 0, 0], [40.0, -25.0, 30.0]], 10.0, U.fixed(4.0, 4.0), nMax, 800.0, theta_deg=35.0,
                  phi_deg=70.0, Eth=0.8, Eph=0.6j)
    orc = U.oracle_case(spec)
    a0 = orc.inc_local(0)
    a1 = orc.inc_local(1)
    the, phi = spec.theta, spec.phi
    khat = np.array([np.sin(the) * np.cos(phi), np.sin(the) * np.sin(phi), np.cos(the)])
    phase = np.exp(1j * K800 * khat.dot(spec.xyz[1]))
    low = np.r_[0:15, nMax * (nMax + 2):nMax * (nMax + 2) + 15]
    assert U.relerr(a1[low], phase * a0[low]) < 1e-9


# --------------------------------------------------------------------------------------------------
# single sphere == Mie (pins populate(), getTLocal and the Result prefactors)
# --------------------------------------------------------------------------------------------------
def _mie_cross_sections(m, x, k, nmax):
    n = np.arange(1, nmax + 1)
    jx, jmx = sp.spherical_jn(n, x), sp.spherical_jn(n, m * x)
    djx, djmx = sp.spherical_jn(n, x, derivative=True), sp.spherical_jn(n, m * x, derivative=True)
    hx = jx + 1j * sp.spherical_yn(n, x)
    dhx = djx + 1j * sp.spherical_yn(n, x, derivative=True)
    psi, dpsi = x * jx, jx + x * djx
    psim, dpsim = m * x * jmx, jmx + m * x * djmx
    xi, dxi = x * hx, hx + x * dhx
    an = (m * psim * dpsi - psi * dpsim) / (m * psim * dxi - xi * dpsim)
    bn = (psim * dpsi - m * psi * dpsim) / (psim * dxi - m * xi * dpsim)
    cext = 2 * np.pi / k ** 2 * np.sum((2 * n + 1) * np.real(an + bn))
    csca = 2 * np.pi / k ** 2 * np.sum((2 * n + 1) * (np.abs(an) ** 2 + np.abs(bn) ** 2))
    return cext, csca


def test_single_sphere_is_mie():
    # eps_r = 12.25+0.01i, r = 150 nm, lambda = 1000 nm, theta 45, phi 90, nMax 10
    spec = U.Spec("mie", [[0, 0, 0]], 150.0, U.fixed(12.25 + 0.01j, 12.25 + 0.01j), 10, 1000.0, sh=False)
    orc = U.oracle_case(spec)
    orc.solve(O.SOLVER_DIRECT)
    cs = orc.cross_sections()
    k = 2 * np.pi / 1000e-9
    cext, csca = _mie_cross_sections(np.sqrt(12.25 + 0.01j), k * 150e-9, k, 10)
    assert abs(cs["ext"] / cext - 1) < 1e-12
    assert abs(cs["sca"] / csca - 1) < 1e-12
    assert abs(cs["ext"] / 3.2480529874220e-13 - 1) < 1e-11
    assert abs(cs["sca"] / 3.2370363598382e-13 - 1) < 1e-11


# --------------------------------------------------------------------------------------------------
# matrix structure and solvers
# --------------------------------------------------------------------------------------------------
def test_matrix_block_structure():
    """preconditioned_scattering_matrix (PreconditionedMatrix.cpp:350-400): identity diagonal blocks,
    off-diagonal block (i,j) = -T_i [[A^T,B^T],[B^T,A^T]] of Coupling(R_i - R_j)."""
    spec = U.three_au(nMax=3)
    orc = U.oracle_case(spec)
    S = orc.matrix(1)
    n = 15
    b = 2 * n
    for i in range(3):
        assert np.array_equal(S[i * b:(i + 1) * b, i * b:(i + 1) * b], np.eye(b))
    xyz = U.spherical_roundtrip(spec.xyz)
    for i, j in ((0, 1), (2, 0), (1, 2)):
        d = xyz[i] - xyz[j]
        r = np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        A, B = O.coupling([r, np.arccos(d[2] / r), np.arctan2(d[1], d[0])], orc.info()["waveK"], 3, True)
        T = orc.particle_factors(i, 0)
        ref = -T[:, None] * np.block([[A.T, B.T], [B.T, A.T]])
        assert U.relerr(S[i * b:(i + 1) * b, j * b:(j + 1) * b], ref) < 1e-13


def test_row_slab_matches_full_matrix():
    spec = U.random_cluster(5, 3, seed=2)
    orc = U.oracle_case(spec)
    S = orc.matrix(2)
    b = 30
    assert np.array_equal(orc.matrix(2, 1, 4), S[b:4 * b])


@pytest.mark.parametrize("maker", [lambda: U.two_si(nMax=6), lambda: U.three_au(nMax=3),
                                   lambda: U.random_cluster(6, 4, seed=4)])
def test_gmres_flavours_vs_direct(maker):
    orc = U.oracle_case(maker())
    S, Q = orc.matrix(1), orc.source()
    xd, _, _ = O.solve_dense(S, Q, O.SOLVER_DIRECT)
    assert U.relerr(S @ xd, Q) < 1e-12
    assert U.relerr(O.matvec(S, xd), S @ xd) < 1e-13
    xz, itz, _ = O.solve_dense(S, Q, O.SOLVER_ZCOMP, tol=1e-13, maxit=300, max_restarts=3)
    xb, itb, _ = O.solve_dense(S, Q, O.SOLVER_BELOS, tol=1e-13, maxit=600, restart=60, max_restarts=20)
    assert U.relerr(xz, xd) < 1e-9 and U.relerr(xb, xd) < 1e-9
    # at the shipped tolerances both flavours stop within the tolerance of the direct solution
    xz, itz, rz = O.solve_dense(S, Q, O.SOLVER_ZCOMP, tol=1e-6, maxit=240, max_restarts=2)
    xb, itb, rb = O.solve_dense(S, Q, O.SOLVER_BELOS, tol=1e-5, maxit=50, restart=30, max_restarts=20)
    assert 0 < itz <= 240 and 0 < itb <= 50
    assert U.relerr(xz, xd) < 1e-3 and U.relerr(xb, xd) < 1e-3


def test_sh_pipeline_consistency():
    """solve(): X_int = Iaux .* X_sca (Solver.cpp:57-77), K from conj(X_int), X_int_SH = IauxSH1 .* X_sca_SH - K1ana
    (Solver.cpp:95-116), V X_sca_SH = K."""
    spec = U.three_au(nMax=3)
    orc = U.oracle_case(spec)
    orc.solve(O.SOLVER_DIRECT)
    xs, xi, xsS, xiS = (orc.vector(w) for w in range(4))
    Iaux = np.concatenate([orc.particle_factors(j, 4) for j in range(3)])
    assert U.relerr(xi, Iaux * xs) < 1e-14
    K, K1 = orc.sh_source(np.conj(xi))
    V = orc.matrix(2)
    assert U.relerr(V @ xsS, K) < 1e-10
    I1 = np.concatenate([orc.particle_factors(j, 5) for j in range(3)])
    assert U.relerr(xiS, I1 * xsS - K1) < 1e-13
    cs = orc.cross_sections()
    assert cs["ext"] > cs["sca"] > 0 and cs["sca_SH"] > 0 and cs["abs_SH"] > 0


def test_as_shipped_switches_do_not_change_results():
    # the "as-shipped" cost model (dense T multiply, Bessel calls inside the SH loops) is arithmetic-neutral
    spec = U.three_au(nMax=3)
    orc = U.oracle_case(spec)
    S0 = orc.matrix(1)
    xi = np.linspace(1, 2, S0.shape[0]) * (1e-3 + 2e-3j)
    K0, _ = orc.sh_source(xi)
    try:
        O.set_as_shipped(1, 1)
        orc2 = U.oracle_case(spec)
        S1 = orc2.matrix(1)
        K1, _ = orc2.sh_source(xi)
    finally:
        O.set_as_shipped(0, 0)
    assert U.relerr(S1, S0) < 1e-13
    assert U.relerr(K1, K0) < 1e-12


def test_rotation_axial_factorisation():
    """A(R) = U A(d z) U^-1 with U = diag(exp(i m phi)) d(theta) (Wigner small-d, Varshalovich 4.3.1), the axial blocks
    diagonal in m with A(-mu) = A(mu), B(-mu) = -B(mu), and the reversed direction by parity: an independent pin of the
    angular dependence of the oracle's translation coefficients, and the exact data layout / index formulas of the
    device's rotated-axial operator (csrc/ob_rot.cu) through tests/rot_model.py."""
    from tests import rot_model as R
    k = 2 * np.pi / 800e-9 * (1.2 + 0.05j)
    NM = 5
    n = NM * (NM + 2)
    # small-d recurrence against the explicit sum, incl. both poles
    for beta in (0.0, 0.3, np.pi / 2, 2.9, np.pi):
        for mp in range(-NM, NM + 1):
            for m in range(-NM, NM + 1):
                col = R.small_d_column(NM, mp, m, beta)
                for j in range(max(abs(mp), abs(m), 1), NM + 1):
                    assert abs(col[j] - R.wd(j, mp, m, beta)) < 1e-13
    rng = np.random.RandomState(1)
    for (d, the, phi) in ((260e-9, 1.1, 0.7), (190e-9, 2.7, -2.0), (400e-9, 0.0, 0.0), (210e-9, np.pi, 0.0),
                          (330e-9, np.pi / 2, np.pi)):
        A, B = O.coupling([d, the, phi], k, NM, True)
        vec = -d * np.array([np.sin(the) * np.cos(phi), np.sin(the) * np.sin(phi), np.cos(the)])
        Ar, Br = O.coupling([d, np.arccos(vec[2] / d), np.arctan2(vec[1], vec[0])], k, NM, True)
        ph, dmat, Aax, Bax, Az, Bz = R.build_pair(NM, d, the, phi, k)
        mask = np.array([[R.flat(1, 1) >= 0 and (p_m == q_m) for q_m in _ms(NM)] for p_m in _ms(NM)])
        assert not np.abs(Az[~mask]).any() and not np.abs(Bz[~mask]).any()      # axial: equal m only
        X = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
        W = R.apply_pair(NM, ph, dmat, Aax, Bax, X, 1.0, False)
        assert U.relerr(W, np.stack([A.T @ X[0] + B.T @ X[1], B.T @ X[0] + A.T @ X[1]])) < 1e-13
        Wr = R.apply_pair(NM, ph, dmat, Aax, Bax, X, -1.0, True)
        assert U.relerr(Wr, np.stack([Ar.T @ X[0] + Br.T @ X[1], Br.T @ X[0] + Ar.T @ X[1]])) < 1e-13


@pytest.mark.parametrize("NM", [1, 2, 5, 10])
def test_rotation_axial_factorisation_flip_basis(NM):
    """Record layout v2 of csrc/ob_rot.cu (tests/rot2_model.py): the flip-symmetric / antisymmetric split of the small-d
    matrices and the (A + B), (A - B) channels reproduce [A^T B^T; B^T A^T](+-R) of the oracle's full blocks."""
    from tests import rot2_model as R2
    k = 2 * np.pi / 800e-9 * (1.2 + 0.05j)
    n = NM * (NM + 2)
    rng = np.random.RandomState(1)
    for (d, the, phi) in ((260e-9, 1.1, 0.7), (190e-9, 2.7, -2.0), (400e-9, 0.0, 0.0), (210e-9, np.pi, 0.0),
                          (330e-9, np.pi / 2, np.pi)):
        A, B = O.coupling([d, the, phi], k, NM, True)
        vec = -d * np.array([np.sin(the) * np.cos(phi), np.sin(the) * np.sin(phi), np.cos(the)])
        Ar, Br = O.coupling([d, np.arccos(vec[2] / d), np.arctan2(vec[1], vec[0])], k, NM, True)
        rec = R2.build_pair(NM, d, the, phi, k)
        X = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
        W = R2.apply_pair(NM, rec, X, False)
        assert U.relerr(W, np.stack([A.T @ X[0] + B.T @ X[1], B.T @ X[0] + A.T @ X[1]])) < 1e-13
        Wr = R2.apply_pair(NM, rec, X, True)
        assert U.relerr(Wr, np.stack([Ar.T @ X[0] + Br.T @ X[1], Br.T @ X[0] + Ar.T @ X[1]])) < 1e-13


def _ms(NM):
    return [m for nn in range(1, NM + 1) for m in range(nn, -nn - 1, -1)]


@pytest.mark.parametrize("NM,k,d", [(6, 2 * np.pi / 800e-9 * (1.2 + 0.05j), 300e-9), (10, 2 * np.pi / 800e-9, 190e-9),
                                    (13, 2 * np.pi / 400e-9, 160e-9)])
def test_axial_only_recursion(NM, k, d):
    """theta = 0: the TA recursion closes on the k = m entries (O(nMax^3) per pair).  tests/rot_axial_model.py is the
    specification of the axial-only assembly kernel of the rotated-axial operator form; here against Coupling."""
    from tests import rot_axial_model as M
    Az, Bz = O.coupling([d, 0.0, 0.0], k, NM, True)
    A, B = M.axial_AB(NM, k * d)
    fl = lambda n, m: n * (n + 1) - m - 1
    sa, sb = np.abs(Az).max(), np.abs(Bz).max()
    for (mu, n, l), v in A.items():
        assert abs(v - Az[fl(n, mu), fl(l, mu)]) < 1e-13 * sa
        assert abs(B[(mu, n, l)] - Bz[fl(n, mu), fl(l, mu)]) < 1e-13 * sb


def test_second_harmonic_at_nmax_10_is_solver_dependent_in_the_reference():
    """The headline order on the headline particles (50 nm Si spheres, 800 nm, nMax 10): the reference's own algorithm,
    restated, does not determine the second-harmonic results to better than a few digits.  The scattered coefficients of
    degree >= 8 are at round-off level and the internal coefficients (Solver.cpp:57-77) multiply them by up to 1e9 before
    the SH source reads them (PreconditionedMatrix.cpp:1347-1436): the oracle's direct solve and its Belos-flavoured GMRES at
    tol 1e-13 agree to 1e-9 on the fundamental harmonic and differ by more than 1e-6 in C_sca,SH.  This is why
    tests/test_gpu_headline.py checks the SH chain at nMax 10 stage by stage on identical inputs."""
    spec = U.random_cluster(4, 10, seed=5)
    orc = U.oracle_case(spec)
    O.set_threads(8)
    try:
        orc.solve(O.SOLVER_DIRECT)
        cd, xd = orc.cross_sections(), orc.vector(0)
        orc.solve(O.SOLVER_BELOS, tol=1e-13, maxit=600, restart=150, max_restarts=5)
        cb, xb = orc.cross_sections(), orc.vector(0)
    finally:
        O.set_threads(1)   # later test modules run the reference's AMOS, which is not thread-safe
    assert abs(cb["ext"] / cd["ext"] - 1) < 1e-9 and abs(cb["sca"] / cd["sca"] - 1) < 1e-9
    assert U.relerr(xb, xd) < 1e-9
    assert abs(cb["sca_SH"] / cd["sca_SH"] - 1) > 1e-6
