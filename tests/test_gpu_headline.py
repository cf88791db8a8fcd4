"""Parity at the headline order (nMax = 10) and on the headline configurations at full size.

* a multi-sphere nMax = 10 FH+SH full step in every operator form (dense slab, TMA-streamed pair form, rotated-axial
  form) against the oracle's direct solve: 1e-9 on the scattered coefficients and all cross sections (BASELINE.json);
* C5 (1000 spheres, nMax 10, N = 240 000) at full size on ONE GPU in the rotated-axial form through size-independent
  properties: block-rows of the product restricted to sampled columns against the oracle's Coupling (FF and SH
  operator), the FF source and the SH sources K, K1ana of sampled particles against the oracle, the residuals
  ||S x - Q||, ||V x - K|| of the solver's answers, and agreement with the pair form on a sub-cluster;
* C4 at full size including the second harmonic; C3 pinned by the device direct solve at 1e-9.
The reference's own sources behind these: srcAna/PreconditionedMatrix.cpp:350-400, 555-610, 1327-1436, Solver.cpp:57-116.
"""
import numpy as np
import pytest

import optimet_b200 as ob
from optimet_b200 import host as H, xmlgen
from oracle import oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("operator", [0, 1, 3])
def test_full_step_at_nmax_10(gpu_ctx, operator):
    """nMax = 10 on 50 nm spheres at 800 nm is beyond what FP64 determines in the SECOND harmonic: the scattered
    coefficients of degree >= 8 sit at round-off level (1e-15 of the dipole terms) and the internal coefficients multiply
    them by up to 1e9 before they enter the SH source (srcAna/Solver.cpp:57-77, PreconditionedMatrix.cpp:1347-1436), so two
    correct solvers differ by 1e-3 in C_sca,SH -- the oracle's own direct and GMRES solves do
    (tests/test_oracle_kats.py::test_second_harmonic_at_nmax_10_is_solver_dependent_in_the_reference).  The step is therefore
    checked link by link: the fundamental harmonic end to end at 1e-9; every second-harmonic stage on the ORACLE's inputs at
    1e-9 (sources, SH solve, cross-section reductions); and the device's own end-to-end SH numbers within the spread the
    reference itself shows."""
    spec = U.random_cluster(7, 10, seed=5)
    orc = U.oracle_case(spec)
    ctx = gpu_ctx
    ctx.set_option("operator", operator)
    try:
        U.configure_ctx(ctx, spec, orc)
        tight = ob.GmresOpts(ob.OB_GMRES_BELOS, 1e-13, 600, 150, 5)
        res = ctx.run(tight)
        orc.solve(O.SOLVER_DIRECT)
        cs = orc.cross_sections()
        xo = [orc.vector(i) for i in range(4)]   # X_sca, X_int, X_sca_SH, X_int_SH
        # fundamental harmonic, end to end
        for key in ("ext", "sca"):
            assert abs(res[key] / cs[key] - 1) < 1e-9, (key, res[key], cs[key])
        assert U.relerr(res["X_sca"], xo[0]) < 1e-9
        # second harmonic, stage by stage on the oracle's inputs
        Ko, K1o = orc.sh_source(np.conj(xo[1]))
        K, K1 = ctx.source_sh(np.conj(xo[1]))
        assert U.relerr(K, Ko) < 1e-9 and U.relerr(K1, K1o) < 1e-9
        xsh, _, _ = ctx.solve(2, Ko, tight)
        assert U.relerr(xsh, xo[2]) < 1e-9
        xish = ctx.unprecondition_sh(xo[2], K1o)
        big = np.abs(xo[3]) > 1e-12 * np.abs(xo[3]).max()
        assert np.max(np.abs(xish[big] / xo[3][big] - 1)) < 1e-9          # element-wise map of Solver.cpp:95-116
        red = ctx.cross_sections(xo[0], xo[1], xo[2], xo[3])
        for key in ("ext", "sca", "sca_SH", "abs_SH"):
            assert abs(red[key] / cs[key] - 1) < 1e-9, (key, red[key], cs[key])
        # the device's own chain end to end: inside the spread of the reference's own solvers
        assert abs(res["sca_SH"] / cs["sca_SH"] - 1) < 2e-2
    finally:
        ctx.set_option("operator", 1)


def _block_row_restricted(orc_info, xyz_m, T, k, nMax, i, cols, x):
    """(S x)_i for x supported on the particles `cols` (i not among them contributes x_i itself): x_i - T_i sum_j
    [[A^T, B^T], [B^T, A^T]](R_i - R_j) x_j from the oracle's Coupling (PreconditionedMatrix.cpp:384-390)."""
    n = nMax * (nMax + 2)
    acc = np.zeros(2 * n, dtype=complex)
    for j in cols:
        if j == i:
            continue
        d = xyz_m[i] - xyz_m[j]
        r = float(np.linalg.norm(d))
        A, B = O.coupling([r, float(np.arccos(d[2] / r)), float(np.arctan2(d[1], d[0]))], k, nMax, True)
        xj = x[j * 2 * n:(j + 1) * 2 * n]
        acc[:n] += A.T @ xj[:n] + B.T @ xj[n:]
        acc[n:] += B.T @ xj[:n] + A.T @ xj[n:]
    return x[i * 2 * n:(i + 1) * 2 * n] - T * acc


def test_c5_full_size_properties(gpu_ctx):
    xyz = xmlgen.random_sites(1000, 2200.0, 150.0, 20261017)
    nobj, nMax = 1000, 10
    spec = U.Spec("c5", xyz, 50.0, U.SI, nMax, 800.0)
    orc = U.oracle_case(spec)
    ctx = gpu_ctx
    ctx.set_option("operator", 3)
    try:
        U.configure_ctx(ctx, spec, orc)
        info = orc.info()
        n = nMax * (nMax + 2)
        b = 2 * n
        xyz_m = np.asarray(xyz) * 1e-9
        rng = np.random.RandomState(11)
        cols = sorted(set(rng.randint(0, nobj, 24).tolist()) | {0, 1, 999})
        rows = [0, 421, 999, cols[3]]
        xs = np.zeros(b * nobj, dtype=complex)
        for j in cols:
            xs[j * b:(j + 1) * b] = rng.standard_normal(b) + 1j * rng.standard_normal(b)
        opts = ob.GmresOpts(ob.OB_GMRES_BELOS, 1e-5, 600, 30, 20)
        for harmonic, kk in ((1, complex(info["waveK"])), (2, 2.0 * complex(info["waveK"]))):
            ctx.assemble(harmonic)
            T = ctx.particle_factors(0 if harmonic == 1 else 1)
            y = ctx.matvec(harmonic, xs)
            for i in rows:   # sampled block-rows of the product against the oracle's Coupling
                ref = _block_row_restricted(info, xyz_m, T[i], kk, nMax, i, cols, xs)
                assert U.relerr(y[i * b:(i + 1) * b], ref) < 1e-11, (harmonic, i)
            untouched = [p for p in range(nobj) if p not in cols][:3]
            assert all(np.abs(xs[p * b:(p + 1) * b]).max() == 0 for p in untouched)
            if harmonic == 1:
                Q = ctx.source_ff()
                Qo = orc.source()
                assert U.relerr(Q, Qo) < 1e-10
                xff, it, _ = ctx.solve(1, Q, opts)
                assert it <= 30 and U.relerr(ctx.matvec(1, xff), Q) < 2e-5   # ||S x - Q|| / ||Q||
                xint = ctx.unprecondition_ff(xff)
        # SH sources of sampled particles: the oracle on a case of those particles alone, fed with the device's X_int
        K, K1 = ctx.source_sh(np.conj(xint))
        pick = [0, 421, 999]
        spec2 = U.Spec("c5pick", [xyz[p] for p in pick], 50.0, U.SI, nMax, 800.0)
        orc2 = U.oracle_case(spec2)
        xi2 = np.concatenate([xint[p * b:(p + 1) * b] for p in pick])
        K2, K12 = orc2.sh_source(np.conj(xi2))
        for t, p in enumerate(pick):
            assert U.relerr(K[p * b:(p + 1) * b], K2[t * b:(t + 1) * b]) < 1e-9
            assert U.relerr(K1[p * b:(p + 1) * b], K12[t * b:(t + 1) * b]) < 1e-9
        xsh, it2, _ = ctx.solve(2, K, opts)
        assert it2 <= 60 and U.relerr(ctx.matvec(2, xsh), K) < 2e-5       # ||V x - K|| / ||K||
        ctx.release_matrix(1)
        ctx.release_matrix(2)
        # the same operator in the TMA-streamed pair form on the first 120 spheres
        sub = 120
        spec3 = U.Spec("c5sub", xyz[:sub], 50.0, U.SI, nMax, 800.0)
        orc3 = U.oracle_case(spec3)
        U.configure_ctx(ctx, spec3, orc3)
        xv = rng.standard_normal(b * sub) + 1j * rng.standard_normal(b * sub)
        ys = {}
        for op in (3, 1):
            ctx.set_option("operator", op)
            for harmonic in (1, 2):
                ctx.assemble(harmonic)
                ys[(op, harmonic)] = ctx.matvec(harmonic, xv)
                ctx.release_matrix(harmonic)
        for harmonic in (1, 2):
            assert U.relerr(ys[(3, harmonic)], ys[(1, harmonic)]) < 1e-12
    finally:
        ctx.set_option("operator", 1)


def test_c4_full_size_second_harmonic(gpu_ctx):
    """C4 at full size, SH operator and SH source: sampled blocks of V against the oracle, one block-row of the product,
    K of sampled particles, the residual of the SH solve; pair form against the rotated-axial form."""
    xyz = xmlgen.cube_sites(7, 200, 190.0)
    nobj, nMax = 200, 8
    spec = U.Spec("c4", xyz, 50.0, U.SI, nMax, 800.0)
    orc = U.oracle_case(spec)
    ctx = gpu_ctx
    U.configure_ctx(ctx, spec, orc)
    ctx.set_option("operator", 1)
    try:
        opts = ob.GmresOpts(ob.OB_GMRES_BELOS, 1e-5, 600, 30, 20)
        rng = np.random.RandomState(9)
        b = 2 * ctx.n(1)
        ctx.assemble(1)
        Q = ctx.source_ff()
        xff, _, _ = ctx.solve(1, Q, opts)
        xint = ctx.unprecondition_ff(xff)
        ctx.release_matrix(1)
        ctx.assemble(2)
        for _ in range(4):
            i, j = rng.randint(0, nobj, 2)
            blk = ctx.fetch_block(2, int(i), int(j))
            ref = orc.matrix(2, int(i), int(i) + 1)[:, j * b:(j + 1) * b]
            assert U.relerr(blk, ref) < 1e-11, (i, j)
        x = rng.standard_normal(ctx.N(2)) + 1j * rng.standard_normal(ctx.N(2))
        y = ctx.matvec(2, x)
        i = 61
        assert U.relerr(y[i * b:(i + 1) * b], O.matvec(orc.matrix(2, i, i + 1), x)) < 1e-12
        K, K1 = ctx.source_sh(np.conj(xint))
        pick = [0, 61, 199]
        spec2 = U.Spec("c4pick", [xyz[p] for p in pick], 50.0, U.SI, nMax, 800.0)
        orc2 = U.oracle_case(spec2)
        K2, K12 = orc2.sh_source(np.conj(np.concatenate([xint[p * b:(p + 1) * b] for p in pick])))
        for t, p in enumerate(pick):
            assert U.relerr(K[p * b:(p + 1) * b], K2[t * b:(t + 1) * b]) < 1e-9
            assert U.relerr(K1[p * b:(p + 1) * b], K12[t * b:(t + 1) * b]) < 1e-9
        xsh, it, _ = ctx.solve(2, K, opts)
        assert it <= 60 and U.relerr(ctx.matvec(2, xsh), K) < 2e-5
        ctx.set_option("operator", 3)
        ctx.assemble(2)
        assert U.relerr(ctx.matvec(2, x), y) < 1e-12
    finally:
        ctx.set_option("operator", 1)


def test_c3_pinned_by_the_direct_solve():
    """C3 (ElevenParticlesSi geometry, nMax 12, N = 3696) without any iterative tolerance: OB_SOLVE_DIRECT on the device
    against the oracle's direct solve.  The preconditioned matrix of this input has a 2-norm condition number of 3e19
    (numpy on the oracle's matrix): LAPACK's and the oracle's own direct solves of the same system differ by 3e-8 in the
    coefficients, the oracle's direct and tight-GMRES solves by 3.7e-9 in C_ext, 1.6e-9 in C_sca and by 100 % in the
    second-harmonic cross sections.  The pin is therefore 2e-8 on the fundamental cross sections and the residual of the
    device's answer; the second harmonic of this input is not determined by FP64 in the reference either."""
    from tests.test_gpu_configs import ELEVEN, ELEVEN_R
    case = H.Case(xml=xmlgen.cluster_xml(ELEVEN, ELEVEN_R, 12, 1000.0))
    solver = H.Solver(case, device=0)
    solver.set_gmres(ob.GmresOpts(ob.OB_SOLVE_DIRECT, 0.0, 1, 1, 0))
    res = solver.step()
    ctx = solver.ctx()
    ctx.set_option("operator", 0)
    ctx.assemble(1)
    Q = ctx.source_ff()
    assert U.relerr(ctx.matvec(1, res["X_sca"]), Q) < 1e-12      # the device's direct solve solves ITS system
    solver.close()
    orc = O.Case()
    for p, r in zip(ELEVEN, ELEVEN_R):
        orc.add_sphere([v * 1e-9 for v in p], r * 1e-9, 12, U.SI[0], U.SI[1])
    orc.set_source(1000e-9, np.deg2rad(45.0), np.deg2rad(90.0), 1.0, 0.0, True)
    assert U.relerr(Q, orc.source()) < 1e-11
    # cross sections of the oracle's tight GMRES (seconds; its direct solve of N = 3696 takes minutes on the CPU): the two
    # differ by 3.7e-9 / 1.6e-9 from each other, see above
    O.set_threads(8)
    try:
        orc.solve(O.SOLVER_BELOS, tol=1e-13, maxit=2000, restart=300, max_restarts=5)
    finally:
        O.set_threads(1)
    cs = orc.cross_sections()
    for key in ("ext", "sca"):
        assert abs(res[key] / cs[key] - 1) < 2e-8, (key, res[key], cs[key])
