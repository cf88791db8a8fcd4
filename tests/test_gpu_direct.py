"""GPU parity of the device direct solve (csrc/ob_lu.cu, OB_SOLVE_DIRECT): the counterpart of the serial reference's
S.colPivHouseholderQr().solve(Q) (srcAna/PreconditionedMatrixSolver.h:58,75 -- the route taken whenever
<ACA compression="no"> and no Belos GMRES is selected) and of pzgesv_ (srcAna/ScalapackSolver.cpp:54-128).

Checked (i) at unit level against numpy's LAPACK zgesv on caller-supplied matrices -- sizes around the 64-column
panel width, matrices that force row interchanges in every column, badly row-scaled ones -- and (ii) on the
scattering systems against the CPU oracle's direct solve: 1e-9 on scattered coefficients and cross sections
(BASELINE.json), with no GMRES tolerance in the way.
"""
import numpy as np
import pytest

import optimet_b200 as ob
from optimet_b200 import host as H, xmlgen
from oracle import oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


def _rand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@pytest.mark.parametrize("n", [1, 2, 31, 63, 64, 65, 127, 128, 129, 200, 777, 1500])
def test_dense_solve_random(gpu_ctx, n):
    rng = np.random.RandomState(100 + n)
    A = _rand(rng, n, n)  # no diagonal dominance: partial pivoting interchanges rows in most columns
    b = _rand(rng, n)
    x = gpu_ctx.dense_solve(A, b)
    xr = np.linalg.solve(A, b)
    assert U.relerr(x, xr) < 1e-10 * max(1.0, np.linalg.cond(A) * 1e-3)
    assert np.linalg.norm(A @ x - b) <= 1e-12 * (np.linalg.norm(A, 2) * np.linalg.norm(x) + np.linalg.norm(b))


def test_dense_solve_needs_pivoting_everywhere(gpu_ctx):
    # anti-diagonal dominant: the pivot of column j is the LAST remaining row; exercises pivots inside and outside
    # the panel's top block and the in-place -> sequential interchange conversion (k_lu_swaplist)
    rng = np.random.RandomState(7)
    n = 300
    A = 1e-3 * _rand(rng, n, n) + np.fliplr(np.diag(5.0 + rng.rand(n)))
    b = _rand(rng, n)
    assert U.relerr(gpu_ctx.dense_solve(A, b), np.linalg.solve(A, b)) < 1e-12
    # a permutation matrix times a well-conditioned one: exact interchanges
    P = np.eye(n)[rng.permutation(n)]
    A2 = P @ (np.eye(n) + 0.01 * _rand(rng, n, n))
    assert U.relerr(gpu_ctx.dense_solve(A2, b), np.linalg.solve(A2, b)) < 1e-12


def test_dense_solve_row_scaled_like_the_scattering_matrix(gpu_ctx):
    # rows scaled over 60 decades (the T_n scaling of S, SURVEY.md section 7 "badly scaled operator")
    rng = np.random.RandomState(9)
    n = 260
    A = np.eye(n) + 0.3 * _rand(rng, n, n) / np.sqrt(n)
    d = 10.0 ** rng.uniform(-30, 30, n)
    A = d[:, None] * A
    b = d * _rand(rng, n)
    assert U.relerr(gpu_ctx.dense_solve(A, b), np.linalg.solve(A, b)) < 1e-10


def test_dense_solve_singular_is_reported(gpu_ctx):
    A = np.ones((70, 70), dtype=complex)
    A[:, 3] = 0.0
    with pytest.raises(RuntimeError, match="singular"):
        gpu_ctx.dense_solve(A, np.ones(70, dtype=complex))


SPECS = {
    "two_si": lambda: U.two_si(nMax=6),             # N = 192: three full panels
    "three_au": lambda: U.three_au(nMax=3),         # N = 90: one full + one ragged panel
    "random7": lambda: U.random_cluster(7, 5, seed=3),  # N = 490
}


@pytest.mark.parametrize("name", sorted(SPECS))
def test_direct_solve_matches_oracle(gpu_ctx, name):
    spec = SPECS[name]()
    orc = U.oracle_case(spec)
    U.configure_ctx(gpu_ctx, spec, orc)
    direct = ob.GmresOpts(ob.OB_SOLVE_DIRECT, 0.0, 0, 0, 0)
    So, Q = orc.matrix(1), orc.source()
    x, it, _ = gpu_ctx.solve(1, Q, direct)
    xd, _, _ = O.solve_dense(So, Q, O.SOLVER_DIRECT)
    assert it == 0
    assert U.relerr(x, xd) < 1e-9
    # whole step on the direct route: update() + solve() FF and SH + cross sections
    res = gpu_ctx.run(direct, do_sh=True)
    orc.solve(O.SOLVER_DIRECT)
    cs = orc.cross_sections()
    assert U.relerr(res["X_sca"], orc.vector(0)) < 1e-9
    assert U.relerr(res["X_sca_SH"], orc.vector(2)) < 1e-9
    for k in ("ext", "sca", "sca_SH", "abs_SH"):
        assert abs(res[k] / cs[k] - 1) < 1e-9, (k, res[k], cs[k])


def test_host_default_is_the_direct_solve_when_aca_is_off():
    # PreconditionedMatrixSolver.h:55-58: ACA off -> colPivHouseholderQr; the host layer must select OB_SOLVE_DIRECT
    xyz = [[-100, -100, 100], [-150, 150, 100], [200, 200, 100]]
    mats = ("gold", 1.0, -1.0, 1.0)
    case = H.Case(xml=xmlgen.cluster_xml(xyz, [50, 100, 150], 3, 400.0, material=mats, aca=False))
    assert case.gmres_defaults().flavour == ob.OB_SOLVE_DIRECT
    solver = H.Solver(case, device=0)
    res = solver.step()
    solver.close()
    orc = U.oracle_case(U.three_au(nMax=3))
    orc.solve(O.SOLVER_DIRECT)
    cs = orc.cross_sections()
    for k in ("ext", "sca", "sca_SH", "abs_SH"):
        assert abs(res[k] / cs[k] - 1) < 1e-9, (k, res[k], cs[k])


def test_direct_solve_refuses_what_does_not_fit(gpu_ctx):
    # 16 N^2 bytes: 1000 spheres at nMax = 10 is 921.6 GB -- must fail loudly, not fall back
    nobj = 1000
    g = np.arange(10) * 400e-9
    xyz = np.array([[x, y, z] for x in g for y in g for z in g])
    gpu_ctx.set_cluster(xyz, np.full(nobj, 50e-9), 10)
    one = np.ones(nobj)
    k0 = 2 * np.pi / 800e-9
    gpu_ctx.set_frequency(299792458.0 * k0, k0, U.EPS0, U.MU0, 13.0 * U.EPS0 * one, U.MU0 * one, 30.0 * U.EPS0 * one,
                          U.MU0 * one, 1e-19 * one, 1e-19 * one, 1e-19 * one)
    with pytest.raises(RuntimeError, match="direct solve needs"):
        gpu_ctx.solve(1, np.zeros(gpu_ctx.N(1), dtype=complex), ob.GmresOpts(ob.OB_SOLVE_DIRECT, 0.0, 0, 0, 0))
