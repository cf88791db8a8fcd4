"""Sharding checked on ONE device: the plans of rank 0..w-1 of a world of w are executed one after the other
(ob_set_shard: plan and assemble as a rank without a communicator; ob_matvec_partial: that shard's partial sums), the
partials are added on the host and y = x - T .* sum must equal the single-rank product and the oracle's.  The driver's
one-GPU box cannot run tests/test_gpu_multirank.py; this covers the partition of the pair list (pair form) and of the
strip list (rotated-axial form), ragged shards and shards without work included."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("operator", [1, 3])
@pytest.mark.parametrize("world,nobj,nMax", [(2, 9, 5), (3, 10, 4), (8, 13, 3), (8, 3, 6), (4, 23, 8)])
def test_shards_of_one_device_add_up(gpu_ctx, operator, world, nobj, nMax):
    spec = U.random_cluster(nobj, nMax, seed=100 + nobj)
    orc = U.oracle_case(spec)
    ctx = gpu_ctx
    ctx.set_option("operator", operator)
    try:
        U.configure_ctx(ctx, spec, orc)
        rng = np.random.RandomState(2)
        for harmonic in (1, 2):
            So = orc.matrix(harmonic)
            x = rng.standard_normal(So.shape[1]) + 1j * rng.standard_normal(So.shape[1])
            ctx.set_shard(0, 1)
            ctx.assemble(harmonic)
            y1 = ctx.matvec(harmonic, x)
            T = ctx.particle_factors(0 if harmonic == 1 else 1).reshape(-1)
            acc = np.zeros_like(x)
            for r in range(world):
                ctx.set_shard(r, world)
                ctx.assemble(harmonic)
                acc += ctx.matvec_partial(harmonic, x)
            y = x - T * acc
            assert U.relerr(y, y1) < 1e-13
            assert U.relerr(y, O.matvec(So, x)) < 1e-12
    finally:
        ctx.set_shard(0, 1)
        ctx.set_option("operator", 1)
