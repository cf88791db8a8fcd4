"""Timing of the device direct solve (csrc/ob_lu.cu) on caller-supplied matrices and on scattering systems.
Usage: python scripts/lu_bench.py [N ...]   (prints wall-clock per solve incl. H2D of the matrix for dense_solve,
and the device-phase timings for whole steps on the direct route)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import optimet_b200 as ob  # noqa: E402
from optimet_b200 import host as H, xmlgen  # noqa: E402


def main():
    steps = "--no-steps" not in sys.argv
    sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [1024, 4096, 8192]
    ctx = ob.Context(0)
    rng = np.random.RandomState(0)
    for n in sizes:
        A = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        ctx.dense_solve(A, b)
        t0 = time.perf_counter()
        x = ctx.dense_solve(A, b)
        dt = time.perf_counter() - t0
        h2d = 16.0 * n * n / 25e9
        res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
        print("dense_solve N=%d: %.1f ms wall (H2D of A ~%.1f ms at 25 GB/s), %.2f TFLOP/s on 8/3 N^3, |Ax-b|/|b| = %.1e"
              % (n, dt * 1e3, h2d * 1e3, 8.0 / 3.0 * n ** 3 / dt / 1e12, res), flush=True)
    ctx.close()
    # whole steps on the direct route (assembly into the LU work matrix + LU + sweeps), FF + SH
    for nobj, nMax in ((11, 12), (64, 8), (200, 8)) if steps else ():
        side = int(np.ceil(nobj ** (1 / 3)))
        xyz = [[190.0 * i, 190.0 * j, 190.0 * k] for k in range(side) for j in range(side) for i in range(side)][:nobj]
        case = H.Case(xml=xmlgen.cluster_xml(xyz, 50.0, nMax, 800.0, aca=False))
        solver = H.Solver(case, device=0)
        solver.step()
        t0 = time.perf_counter()
        res = solver.step()
        dt = time.perf_counter() - t0
        N = 2 * nMax * (nMax + 2) * nobj
        t = solver.ctx_timings()
        print("direct step nobj=%d nMax=%d N=%d: %.1f ms wall, solve_ff %.1f ms solve_sh %.1f ms -> %.2f TFLOP/s per LU; ext=%.6e"
              % (nobj, nMax, N, dt * 1e3, t["solve_ff"], t["solve_sh"], 8.0 / 3.0 * N ** 3 / (t["solve_ff"] * 1e-3) / 1e12,
                 res["ext"]), flush=True)
        solver.close()


if __name__ == "__main__":
    main()
