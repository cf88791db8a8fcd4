#!/usr/bin/env python
"""Field-map throughput: grid points per second of Result::setFields on the B200 path
against the oracle port on the host cores, for a C4-like cluster (N spheres on the reference's cube lattice, nMax 8).

Algorithmic work per point (stated, not measured): outside the spheres one vector spherical wave set per particle and
harmonic, n (FF) + n_S (SH) harmonics x N_obj particles x 2 wave types (M, N) x 3 components complex MACs x (E, H).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nobj", type=int, default=200)
    ap.add_argument("--nmax", type=int, default=8)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--cpu-points", type=int, default=64)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    import optimet_b200 as ob
    from optimet_b200 import capi, host as H, xmlgen
    from oracle import oracle as O
    side = int(np.ceil(args.nobj ** (1 / 3.0)))
    xyz = xmlgen.cube_sites(side, args.nobj)
    ext = 190.0 * (side - 1)
    # two z-planes (through the second lattice layer and between layers), reaching 300 nm beyond the cluster: crosses spheres and gaps
    grid = ((-300.0, ext + 300.0, args.grid), (-300.0, ext + 300.0, args.grid), (190.0, 285.0, 2))  # steps = 1 divides by zero in OutputGrid (:53)
    case = H.Case(xml=xmlgen.cluster_xml(xyz, 50.0, args.nmax, 800.0, field=grid))
    solver = H.Solver(case, device=0)
    solver.set_gmres(ob.GmresOpts(ob.OB_GMRES_BELOS, 1e-5, 1000, 30, 20))
    res = solver.step(800e-9)
    ctx = solver.ctx()
    pts = case.grid_points()
    lib = capi.load()
    ms = C.c_double()
    ctx.fields(pts[:1024])  # warm-up
    times = []
    for _ in range(args.repeat):
        t0 = time.perf_counter()
        got, inner = ctx.fields(pts)
        times.append(time.perf_counter() - t0)
    t_gpu = min(times)
    # CPU oracle on a bounded sample of the same points, same solution vectors
    O.set_threads(os.cpu_count() or 1)
    orc = O.Case()
    for p in xyz:
        orc.add_sphere([v * 1e-9 for v in p], 50e-9, args.nmax, O.MODEL_SILICON, [1.0, 0.0])
    orc.set_source(800e-9, np.deg2rad(45.0), np.deg2rad(90.0), 1.0, 0.0, True)
    for w, key in enumerate(("X_sca", "X_int", "X_sca_SH", "X_int_SH")):
        orc.set_vector(w, res[key])
    rng = np.random.RandomState(0)
    pick = np.sort(rng.choice(len(pts), size=min(args.cpu_points, len(pts)), replace=False))
    orc.fields(pts[pick[:2]])  # builds the CG tables outside the timed region
    t0 = time.perf_counter()
    want, inner_o = orc.fields(pts[pick])
    t_cpu = time.perf_counter() - t0
    scale = np.abs(want).max(axis=(0, 2))
    err = [float(np.abs(got[pick][:, t] - want[:, t]).max() / scale[t]) for t in range(4)]
    n, ns = args.nmax * (args.nmax + 2), args.nmax * (args.nmax + 2)
    print(json.dumps({
        "metric": "field_map_points_per_s", "workload": "%d Si spheres nMax %d, %dx%dx2 grid, FH+SH" % (args.nobj, args.nmax, args.grid, args.grid),
        "points": int(len(pts)), "interior_points": int((inner >= 0).sum()),
        "gpu_s_e2e": t_gpu, "gpu_points_per_s": len(pts) / t_gpu,
        "h2d_bytes": int(pts.nbytes), "d2h_bytes": int(got.nbytes + inner.nbytes),
        "cpu_baseline": {"kind": "port", "cores": os.cpu_count(), "sample": "%d of the same points" % len(pick),
                         "s": t_cpu, "points_per_s": len(pick) / t_cpu},
        "speedup": (len(pts) / t_gpu) / (len(pick) / t_cpu),
        "max_rel_err_vs_oracle_on_sample": err,
        "wave_evaluations_per_exterior_point": args.nobj * (n + ns) + n}))
    solver.close()


if __name__ == "__main__":
    main()
