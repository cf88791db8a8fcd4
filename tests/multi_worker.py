"""Worker for tests/test_gpu_multi.py: one process driving `ngpu` GPUs through ob_create_multi; run in a subprocess
so that a deadlock between the group's ranks cannot hang the test session."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path, ngpu, operator = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    import optimet_b200 as ob
    from optimet_b200 import capi
    from tests import util as U
    spec = U.random_cluster(11, 5, seed=8)
    orc = U.oracle_case(spec)
    opts = ob.GmresOpts(ob.OB_GMRES_BELOS, 1e-12, 400, 100, 5)
    m = capi.MultiContext(list(range(ngpu)))
    assert m.size == ngpu
    m.set_option("operator", operator)
    U.configure_ctx(m, spec, orc)
    res = m.run(opts)
    res2 = m.run(opts)   # a second step on the same group
    m.close()
    assert np.array_equal(res["X_sca"], res2["X_sca"])
    np.savez(out_path, cs=np.array([res[k] for k in ("ext", "sca", "abs", "sca_SH", "abs_SH")]), X_sca=res["X_sca"],
             X_sca_SH=res["X_sca_SH"], iters=np.array([res["iters_ff"], res["iters_sh"]]))


if __name__ == "__main__":
    main()
