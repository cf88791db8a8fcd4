"""Row-sharded multi-GPU path on real GPUs (skipped on a 1-GPU box): N ranks over NCCL must reproduce the
single-GPU step -- same iteration counts, coefficients and cross sections to 1e-9."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from tests import util as U

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,nobj,mode", [(2, 7, "pairs"), (2, 7, "dense"), (2, 7, "aca"), (2, 7, "rot"), (4, 10, "pairs"),
                                             (4, 9, "aca"), (8, 11, "pairs"), (2, 23, "rot"), (4, 10, "rot"), (8, 11, "rot")])
def test_sharded_step_matches_single_gpu(tmp_path, world, nobj, mode):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    nMax = 4
    outs = {}
    for w in (1, world):
        out = str(tmp_path / ("w%d.npz" % w))
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(w), "--master-addr",
               "127.0.0.1", "--master-port", str(_port()), os.path.join(ROOT, "tests", "multirank_worker.py"), out,
               str(nobj), str(nMax), mode]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-3000:]
        outs[w] = np.load(out)
    a, b = outs[1], outs[world]
    assert b["same"] == 1.0
    assert np.all(np.abs(a["iters"] - b["iters"]) <= 1)
    for k in ("X_sca", "X_sca_SH"):
        assert U.relerr(b[k], a[k]) < 1e-9, k
    assert np.max(np.abs(b["cs"] / a["cs"] - 1)) < 1e-9
