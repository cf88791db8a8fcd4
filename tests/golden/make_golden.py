#!/usr/bin/env python
"""Generates tests/golden/*.npz -- small committed fixtures for the hot path.

Provenance: the reference itself cannot be built in this container (Eigen, Boost, GSL and HDF5 are absent) and
ships no golden data, so these vectors come from the CPU oracle (oracle/optimet_oracle.cpp) running with the
REFERENCE'S OWN Bessel/Hankel code (srcAna/amos.c compiled into oracle/_ref/libamos_ref.so, backend 1) on the
reference's own example inputs (geometry/material/source of examples/TwoParticlesSi.xml and
examples/ThreeParticlesAu.xml) plus one lossy-background case.  They freeze today's oracle so that (a) the oracle
cannot drift silently and (b) the GPU box, which has no /root/reference, still checks the CUDA path against
AMOS-backed numbers.

    python tests/golden/make_golden.py        (run in the build container; needs oracle/_ref)
    python tests/golden/make_golden.py 8f     (only the fixtures of the SURVEY 8f rows: ACA operator, field maps)
    python tests/golden/make_golden.py ref    (vtac_reference.npz: A/B blocks, Mie factors, incident coefficients, CG
                                               tables straight from the reference's own compiled translation units,
                                               oracle/_ref/libpath_ref.so -- no oracle arithmetic involved)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import util as U  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "two_si_nmax6": lambda: U.two_si(nMax=6, wavelength_nm=1240.0),
    "three_au_nmax3": lambda: U.three_au(nMax=3, wavelength_nm=400.0),
    "lossy_bg_nmax4": lambda: U.Spec("lossy_bg", [[0, 0, 0], [260, 40, -90], [-30, 310, 120]], [60, 80, 70],
                                     U.fixed(9.0 + 0.4j, 7.0 + 0.9j), 4, 700.0, theta_deg=30, phi_deg=20, Eth=0.6,
                                     Eph=0.8j, background=(1.7 + 0.0j, 1.0 + 0.0j)),
}

VTAC_CASES = [  # (R spherical [m, rad, rad], k, nMax, regular_flag)
    ([190e-9, 0.9, 2.2], 2 * np.pi / 800e-9, 6, True),
    ([120e-9, 2.4, -1.0], 2 * np.pi / 800e-9, 8, True),
    ([450e-9, 1.1, 0.3], 2 * np.pi / 800e-9 * (1.33 + 0.02j), 5, False),
    ([300e-9, 0.0, 0.0], 2 * np.pi / 800e-9, 4, True),
]


def main():
    if not O.have_amos():
        raise SystemExit("oracle/_ref/libamos_ref.so missing: run `make -C oracle` with /root/reference present")
    O.set_bessel_backend(1)
    try:
        for name, make in CASES.items():
            spec = make()
            orc = U.oracle_case(spec)
            S1 = orc.matrix(1)
            S2 = orc.matrix(2)
            Q = orc.source()
            orc.solve(O.SOLVER_DIRECT)
            cs = orc.cross_sections()
            xs, xi, xsS, xiS = (orc.vector(w) for w in range(4))
            K, K1 = orc.sh_source(np.conj(xi))
            nobj = orc.info()["nobj"]
            fac = np.array([np.concatenate([orc.particle_factors(j, w) for j in range(nobj)]) for w in range(7)])
            # iteration counts of both GMRES flavours at the shipped tolerances
            _, it_z, _ = O.solve_dense(S1, Q, O.SOLVER_ZCOMP, tol=1e-6, maxit=240, max_restarts=2)
            _, it_b, _ = O.solve_dense(S1, Q, O.SOLVER_BELOS, tol=1e-5, maxit=50, restart=30, max_restarts=20)
            b = S1.shape[0] // nobj
            np.savez_compressed(
                os.path.join(HERE, name + ".npz"), block_ff_10=S1[b:2 * b, 0:b], block_sh_01=S2[0:b, b:2 * b],
                S1_fro=np.linalg.norm(S1), S2_fro=np.linalg.norm(S2), Q=Q, X_sca=xs, X_int=xi, X_sca_SH=xsS,
                X_int_SH=xiS, K=K, K1ana=K1, factors=fac,
                cs=np.array([cs["ext"], cs["sca"], cs["sca_SH"], cs["abs_SH"]]), iters=np.array([it_z, it_b]))
            print(name, cs, it_z, it_b)
        out = {}
        for i, (R, k, nMax, flag) in enumerate(VTAC_CASES):
            A, B = O.coupling(R, k, nMax, flag)
            out["A%d" % i], out["B%d" % i] = A, B
        np.savez_compressed(os.path.join(HERE, "vtac.npz"), **out)
        T = O.cg_tables(2, 2)
        np.savez_compressed(os.path.join(HERE, "cg_tables_nmax2.npz"), T=T)
    finally:
        O.set_bessel_backend(0)


def field_points(spec, seed=11):
    """Fixed sample of points around and inside the spheres of a case (spherical coordinates, as OutputGrid returns)."""
    rng = np.random.RandomState(seed)
    lo, hi = spec.xyz.min(0) - 200e-9, spec.xyz.max(0) + 200e-9
    pts = [rng.uniform(lo, hi) for _ in range(12)]
    for c, r in zip(spec.xyz, spec.radius):
        for frac in (0.3, 0.8):
            d = rng.standard_normal(3)
            pts.append(c + frac * r * d / np.linalg.norm(d))
    pts = np.array(pts)
    r = np.linalg.norm(pts, axis=1)
    return np.stack([r, np.arccos(pts[:, 2] / r), np.arctan2(pts[:, 1], pts[:, 0])], 1)


def main_8f():
    """ACA operator (block compression of the golden FF block, cross sections of the compressed solve) and field maps
    (E/H at fixed points from the golden solution vectors) -- same provenance as above (oracle on the reference's AMOS)."""
    if not O.have_amos():
        raise SystemExit("oracle/_ref/libamos_ref.so missing: run `make -C oracle` with /root/reference present")
    O.set_bessel_backend(1)
    try:
        for name, make in CASES.items():
            spec = make()
            orc = U.oracle_case(spec)
            g = np.load(os.path.join(HERE, name + ".npz"))
            Um, Vm, I, J = O.aca_compress(g["block_ff_10"])
            nobj = orc.info()["nobj"]
            ranks = np.array([[orc.aca_block(1, i, j)[0] for j in range(nobj)] for i in range(nobj)])
            orc.solve(O.SOLVER_ACA_ZCOMP, tol=1e-13, maxit=200, max_restarts=3)
            cs = orc.cross_sections()
            x_aca = orc.vector(0)
            pts = field_points(spec)
            for w, key in enumerate(("X_sca", "X_int", "X_sca_SH", "X_int_SH")):
                orc.set_vector(w, g[key])
            fields, inner = orc.fields(pts)
            np.savez_compressed(os.path.join(HERE, "rows8f_" + name + ".npz"), aca_U=Um, aca_V=Vm, aca_I=I, aca_J=J,
                                aca_ranks=ranks, aca_X_sca=x_aca,
                                aca_cs=np.array([cs["ext"], cs["sca"], cs["sca_SH"], cs["abs_SH"]]), field_points=pts,
                                fields=fields, inner=inner)
            print(name, "rank", len(I), "ranks", ranks.tolist(), cs, "inner", inner.tolist())
    finally:
        O.set_bessel_backend(0)


def main_ref():
    """Fixtures produced by the reference's own compiled code (oracle/reference_build.py), so that the GPU box and any
    later container without /root/reference still hold the reference's numbers."""
    from oracle import reference_build as RB
    if not RB.have():
        raise SystemExit("oracle/_ref/libpath_ref.so missing: run `make -C oracle` with /root/reference present")
    out = {}
    for i, (R, k, nMax, flag) in enumerate(VTAC_CASES):
        out["A%d" % i], out["B%d" % i] = RB.coupling(R, k, nMax, flag)
    for name, make in CASES.items():
        spec = make()
        bg = spec.background if spec.background is not None else (1.0, 1.0)
        fac = np.array([np.concatenate([RB.particle_factors(m, p, r, spec.nMax, spec.wavelength, w, bg)
                                        for r, (m, p) in zip(spec.radius, spec.material)]) for w in range(7)])
        a, b, wk = RB.excitation(spec.wavelength, spec.theta, spec.phi, spec.Eth, spec.Eph, spec.nMax, bg)
        out["factors_" + name], out["a_" + name], out["b_" + name], out["waveK_" + name] = fac, a, b, np.array(wk)
    ref = RB.case_from_spec(CASES["three_au_nmax3"]())
    out["cg_tables_nmax3"] = np.array([ref.cg_table(t) for t in range(9)])
    np.savez_compressed(os.path.join(HERE, "reference_build.npz"), **out)
    print("reference_build.npz:", sorted(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ref":
        main_ref()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "8f":
        main_8f()
        sys.exit(0)
    main()
