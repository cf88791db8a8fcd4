"""One process, several GPUs (ob_create_multi): the group must reproduce the single-context step and the oracle.  With
one visible GPU the group of one is exercised; with more, every available power of two (needs a multi-GPU box).  The
group runs in a subprocess with a timeout: a deadlock between its ranks fails the test instead of hanging the session."""
import os
import subprocess
import sys

import numpy as np
import pytest

import optimet_b200 as ob
from oracle import oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("operator", [1, 3])
@pytest.mark.parametrize("ngpu", [1, 2, 4, 8])
def test_single_process_group_matches_one_context(gpu_ctx, tmp_path, ngpu, operator):
    if _ngpu() < ngpu:
        pytest.skip("needs %d GPUs" % ngpu)
    spec = U.random_cluster(11, 5, seed=8)
    orc = U.oracle_case(spec)
    opts = ob.GmresOpts(ob.OB_GMRES_BELOS, 1e-12, 400, 100, 5)
    gpu_ctx.set_option("operator", operator)
    try:
        U.configure_ctx(gpu_ctx, spec, orc)
        one = gpu_ctx.run(opts)
    finally:
        gpu_ctx.set_option("operator", 1)
    out = str(tmp_path / "multi.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multi_worker.py"), out, str(ngpu), str(operator)],
                       capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    res = np.load(out)
    assert np.all(np.abs(res["iters"] - np.array([one["iters_ff"], one["iters_sh"]])) <= 1)
    ref = np.array([one[k] for k in ("ext", "sca", "abs", "sca_SH", "abs_SH")])
    assert np.max(np.abs(res["cs"] / ref - 1)) < 1e-9
    assert U.relerr(res["X_sca"], one["X_sca"]) < 1e-9 and U.relerr(res["X_sca_SH"], one["X_sca_SH"]) < 1e-9
    orc.solve(O.SOLVER_DIRECT)
    cs = orc.cross_sections()
    for k, key in enumerate(("ext", "sca", None, "sca_SH", "abs_SH")):
        if key:
            assert abs(res["cs"][k] / cs[key] - 1) < 1e-9, key
