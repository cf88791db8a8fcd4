"""GPU parity of the rotated-axial operator form (operator = 3, optimet_b200/csrc/ob_rot.cu) against the oracle's dense
operator: same reference operator as the dense / pair forms (srcAna/PreconditionedMatrix.cpp:350-400, 555-610 applied
by srcAna/scalapack/Belos.hpp:74-90), stored per pair as phases, axial A/B and Wigner small-d matrices.  Complex FP64;
matvec 1e-12 against the oracle's dense product, full step 1e-9 (BASELINE.json), iteration counts +-1."""
import numpy as np
import pytest

import optimet_b200 as ob
from oracle import oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture
def rot_ctx(gpu_ctx):
    gpu_ctx.set_option("operator", 3)
    yield gpu_ctx
    gpu_ctx.set_option("operator", 1)


CLUSTERS = {
    "random40_n3": lambda: U.random_cluster(40, 3, seed=43),
    "random12_n6": lambda: U.random_cluster(12, 6, seed=18),
    "random3_n13": lambda: U.random_cluster(3, 13, seed=16),
    "random23_n8": lambda: U.random_cluster(23, 8, seed=31),
    "random9_n10": lambda: U.random_cluster(9, 10, seed=77),          # the headline order (C5), several row blocks
    "random4_n2": lambda: U.random_cluster(4, 2, seed=102),          # every order 1 .. 13 has its own kernel instantiation
    "random4_n7": lambda: U.random_cluster(4, 7, seed=107),
    "random4_n9": lambda: U.random_cluster(4, 9, seed=109),
    "random4_n11": lambda: U.random_cluster(4, 11, seed=111),
    "random4_n12": lambda: U.random_cluster(4, 12, seed=112),        # 12 rows = a full tile + a stacked tile of four
    "pair_n1": lambda: U.random_cluster(2, 1, seed=3),
    "single_n4": lambda: U.random_cluster(1, 4, seed=5),
    "two_si_z_axis": lambda: U.two_si(nMax=6),                       # theta = 0 / pi
    "cube27_n4": lambda: U.cube_lattice(3, 4),                       # axis-aligned pairs: theta, phi at 0, pi/2, pi
    "three_au": lambda: U.three_au(nMax=3),
    "lossy_bg": lambda: U.Spec("lossy_bg", [[0, 0, 0], [260, 40, -90], [-30, 310, 120]], [60, 80, 70],
                               U.fixed(9.0 + 0.4j, 7.0 + 0.9j), 4, 700.0, theta_deg=30, phi_deg=20, Eth=0.6, Eph=0.8j,
                               background=(1.7 + 0.0j, 1.0 + 0.0j)),
}


@pytest.mark.parametrize("name", sorted(CLUSTERS))
@pytest.mark.parametrize("harmonic", [1, 2])
def test_rot_operator_matvec(rot_ctx, name, harmonic):
    spec = CLUSTERS[name]()
    orc = U.oracle_case(spec)
    U.configure_ctx(rot_ctx, spec, orc)
    rot_ctx.assemble(harmonic)
    So = orc.matrix(harmonic)
    rng = np.random.RandomState(7)
    for _ in range(2):
        x = rng.standard_normal(So.shape[1]) + 1j * rng.standard_normal(So.shape[1])
        assert U.relerr(rot_ctx.matvec(harmonic, x), O.matvec(So, x)) < 1e-12
    with pytest.raises(RuntimeError, match="rotated-axial"):
        rot_ctx.fetch_block(harmonic, 0, 0)


@pytest.mark.parametrize("name", ["random12_n6", "two_si_z_axis", "three_au", "lossy_bg", "cube27_n4"])
def test_rot_operator_full_step(rot_ctx, name):
    spec = CLUSTERS[name]()
    orc = U.oracle_case(spec)
    U.configure_ctx(rot_ctx, spec, orc)
    res = rot_ctx.run(ob.GmresOpts(ob.OB_GMRES_BELOS, 1e-13, 400, 100, 5))
    orc.solve(O.SOLVER_DIRECT)
    cs = orc.cross_sections()
    for key in ("ext", "sca", "sca_SH", "abs_SH"):
        assert abs(res[key] / cs[key] - 1) < 1e-9, key
    assert U.relerr(res["X_sca"], orc.vector(0)) < 1e-9 and U.relerr(res["X_sca_SH"], orc.vector(2)) < 1e-9


@pytest.mark.parametrize("flavour", ["zcomp", "belos"])
def test_rot_operator_iteration_counts(rot_ctx, flavour):
    spec = U.random_cluster(9, 5, seed=21)
    orc = U.oracle_case(spec)
    U.configure_ctx(rot_ctx, spec, orc)
    rot_ctx.assemble(1)
    So, Q = orc.matrix(1), orc.source()
    if flavour == "zcomp":
        opts = ob.GmresOpts(ob.OB_GMRES_ZCOMP, 1e-6, 240, 0, 2)
        xo, ito, _ = O.solve_dense(So, Q, O.SOLVER_ZCOMP, tol=1e-6, maxit=240, max_restarts=2)
    else:
        opts = ob.GmresOpts(ob.OB_GMRES_BELOS, 1e-5, 50, 30, 20)
        xo, ito, _ = O.solve_dense(So, Q, O.SOLVER_BELOS, tol=1e-5, maxit=50, restart=30, max_restarts=20)
    x, it, _ = rot_ctx.solve(1, Q, opts)
    assert abs(it - ito) <= 1 and U.relerr(x, xo) < 1e-7


@pytest.mark.parametrize("name", ["random12_n6", "random3_n13", "random23_n8", "two_si_z_axis", "lossy_bg"])
def test_rot_assembly_paths_agree(rot_ctx, name):
    """The default O(nMax^3) axial-only assembly ("rot_assembly" = 1, k_assemble_axial_only) and the cross-check path
    through the shared vtac_block code ("rot_assembly" = 0, k_assemble_axial) must both give the oracle's operator."""
    spec = CLUSTERS[name]()
    orc = U.oracle_case(spec)
    U.configure_ctx(rot_ctx, spec, orc)
    rng = np.random.RandomState(3)
    for harmonic in (1, 2):
        So = orc.matrix(harmonic)
        x = rng.standard_normal(So.shape[1]) + 1j * rng.standard_normal(So.shape[1])
        ys = []
        for path in (0, 1):
            rot_ctx.set_option("rot_assembly", path)
            try:
                rot_ctx.assemble(harmonic)
                ys.append(rot_ctx.matvec(harmonic, x))
            finally:
                rot_ctx.set_option("rot_assembly", 1)
        yo = O.matvec(So, x)
        assert U.relerr(ys[0], yo) < 1e-12 and U.relerr(ys[1], yo) < 1e-12


@pytest.mark.parametrize("rows,ctas", [(1, 0), (2, 1), (3, 0), (7, 2)])
def test_rot_plan_geometry(rot_ctx, rows, ctas):
    """Rows per block and CTAs per SM only change the order of the fixed-order sums: same operator for every plan
    (blocks of 1, 2, 3 and 7 rows: ragged last blocks, strips shorter than a block, one CTA per SM)."""
    spec = U.random_cluster(23, 5, seed=12)
    orc = U.oracle_case(spec)
    U.configure_ctx(rot_ctx, spec, orc)
    rng = np.random.RandomState(5)
    rot_ctx.set_option("rot_rows", rows)
    rot_ctx.set_option("rot_ctas_per_sm", ctas)
    try:
        for harmonic in (1, 2):
            So = orc.matrix(harmonic)
            x = rng.standard_normal(So.shape[1]) + 1j * rng.standard_normal(So.shape[1])
            rot_ctx.assemble(harmonic)
            assert U.relerr(rot_ctx.matvec(harmonic, x), O.matvec(So, x)) < 1e-12
    finally:
        rot_ctx.set_option("rot_rows", 0)
        rot_ctx.set_option("rot_ctas_per_sm", 0)


@pytest.mark.parametrize("name", ["random9_n10", "random23_n8", "lossy_bg"])
def test_rot_geometry_sections_shared_between_harmonics(rot_ctx, name):
    """Phases and small-d matrices do not depend on k: the harmonic assembled second fetches them from the records of
    the first ("rot_share" = 1, default) instead of computing and storing its own copy.  Same bits either way, in either
    assembly order; releasing the records the sections live in sends the other harmonic back to assembly."""
    spec = CLUSTERS[name]()
    orc = U.oracle_case(spec)
    U.configure_ctx(rot_ctx, spec, orc)
    rng = np.random.RandomState(11)
    N = orc.matrix(1).shape[1]
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    ys = {}
    try:
        for share, order in ((0, (1, 2)), (1, (1, 2)), (1, (2, 1))):
            rot_ctx.set_option("rot_share", share)   # marks both harmonics unassembled
            for h in order:
                rot_ctx.assemble(h)
            ys[(share, order)] = [rot_ctx.matvec(h, x) for h in (1, 2)]
        for key in ((1, (1, 2)), (1, (2, 1))):
            for h in (0, 1):
                assert np.array_equal(ys[key][h], ys[(0, (1, 2))][h]), (key, h)
        for h in (1, 2):
            assert U.relerr(ys[(1, (1, 2))][h - 1], O.matvec(orc.matrix(h), x)) < 1e-12
        # harmonic 1 reads its geometry sections from harmonic 2's records (last order above): drop them
        rot_ctx.release_matrix(2)
        with pytest.raises(RuntimeError, match="not assembled"):
            rot_ctx.matvec(1, x)
        rot_ctx.assemble(1)
        assert np.array_equal(rot_ctx.matvec(1, x), ys[(0, (1, 2))][0])
        # an owner that is assembled again (same cluster) keeps serving the harmonic that shares from it
        rot_ctx.assemble(2)
        rot_ctx.assemble(1)
        assert np.array_equal(rot_ctx.matvec(2, x), ys[(0, (1, 2))][1])
    finally:
        rot_ctx.set_option("rot_share", 1)
