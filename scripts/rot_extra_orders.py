#!/usr/bin/env python
"""Rotated-axial operator at the orders the GPU suite does not parametrise (nMax 2, 7, 9, 11, 12): matvec of both
harmonics against the oracle's dense operator on a 4-sphere cluster, 1e-12.  python scripts/rot_extra_orders.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimet_b200 as ob  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import util as U  # noqa: E402

ctx = ob.Context(0)
ctx.set_option("operator", 3)
rng = np.random.RandomState(1)
worst = 0.0
for nMax in (2, 7, 9, 11, 12):
    spec = U.random_cluster(4, nMax, seed=100 + nMax)
    orc = U.oracle_case(spec)
    U.configure_ctx(ctx, spec, orc)
    for h in (1, 2):
        So = orc.matrix(h)
        x = rng.standard_normal(So.shape[1]) + 1j * rng.standard_normal(So.shape[1])
        ctx.assemble(h)
        err = U.relerr(ctx.matvec(h, x), O.matvec(So, x))
        worst = max(worst, err)
        print("nMax %2d harmonic %d: %.2e" % (nMax, h, err), flush=True)
assert worst < 1e-12, worst
print("ok", worst)
