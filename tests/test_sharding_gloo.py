"""world_size-2 (and 3) `gloo` tests of the row-sharded path on CPU: the partition, the slice layout of the
replicated Krylov vectors, the per-iteration slice all-gather and the partial cross-section sums -- with the
oracle's row slabs standing in for the device slabs (the CUDA kernels themselves are covered by -m gpu)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nobj, nMax, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from optimet_b200 import capi, sharding
    from oracle import oracle as O
    from tests import util as U
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = U.random_cluster(nobj, nMax, seed=9)
        orc = U.oracle_case(spec)
        blk = 2 * nMax * (nMax + 2)
        N = blk * nobj
        first, count = capi.Library.partition(nobj, world, rank)
        lo, hi = sharding.row_range(nobj, world, rank, blk)
        assert (lo, hi) == (first * blk, (first + count) * blk)
        slab = orc.matrix(1, first, first + count)       # this rank's block-rows x all columns
        Q = orc.source()

        def apply(x):                                     # y = S x: local slab product + slice all-gather
            y = np.zeros(N, dtype=np.complex128)
            if count:
                y[lo:hi] = O.matvec(slab, x)
            sharding.allgather_slices(dist, torch, y.view(np.float64), nobj, world, 2 * blk)
            return y

        # restarted GMRES (Gmres_Zcomp semantics: x0 = 0, MGS) with replicated vectors: identical on every rank
        x = np.zeros(N, dtype=np.complex128)
        V = [Q / np.linalg.norm(Q)]
        Hm = np.zeros((41, 40), dtype=np.complex128)
        for j in range(40):
            w = apply(V[j])
            for t in range(j + 1):
                Hm[t, j] = np.vdot(V[t], w)
                w = w - Hm[t, j] * V[t]
            Hm[j + 1, j] = np.linalg.norm(w)
            V.append(w / Hm[j + 1, j])
            e1 = np.zeros(j + 2, dtype=np.complex128)
            e1[0] = np.linalg.norm(Q)
            yk, res, _, _ = np.linalg.lstsq(Hm[:j + 2, :j + 1], e1, rcond=None)
            r = np.linalg.norm(Hm[:j + 2, :j + 1] @ yk - e1) / e1[0].real
            if r < 1e-12:
                break
        x = sum(c * v for c, v in zip(yk, V))
        # partial extinction sums over the rank's own particles (Result.cpp:557-577), summed over ranks
        part = 0.0
        for jl in range(first, first + count):
            q = orc.inc_local(jl)
            part += float(np.real(np.vdot(q, x[jl * blk:(jl + 1) * blk])))
        total = sharding.sum_partials(dist, torch, [part])[0]
        mx = sharding.max_over_ranks(dist, torch, rank + 1.0)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), x=x, total=total, mx=mx, iters=j + 1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nobj", [(2, 5), (3, 4)])
def test_row_sharded_solve_matches_single_rank(tmp_path, world, nobj):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from tests import util as U
    nMax = 3
    mp.spawn(_worker, args=(world, _free_port(), nobj, nMax, str(tmp_path)), nprocs=world, join=True)
    spec = U.random_cluster(nobj, nMax, seed=9)
    orc = U.oracle_case(spec)
    S, Q = orc.matrix(1), orc.source()
    xd, _, _ = O.solve_dense(S, Q, O.SOLVER_DIRECT)
    blk = 2 * nMax * (nMax + 2)
    ext = sum(float(np.real(np.vdot(orc.inc_local(j), xd[j * blk:(j + 1) * blk]))) for j in range(nobj))
    outs = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for o in outs:
        assert np.array_equal(o["x"], outs[0]["x"])       # replicated Krylov vectors stay bit-identical
        assert U.relerr(o["x"], xd) < 1e-9
        assert abs(o["total"] / ext - 1) < 1e-9
        assert o["mx"] == world
        assert o["iters"] == outs[0]["iters"]


def test_reference_arm_under_torchrun_prints_one_line(tmp_path):
    """bench.py --impl reference launched like the driver does for N > 1: rank 0 alone prints the JSON line,
    the other rank exits 0 without work."""
    import json
    import subprocess
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference",
           "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "small"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference", "reference-build+port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["higher_is_better"] is False
    # a sampled arm says so, names the CPU, and carries the section 8(d) variants
    assert d["extrapolated"] is True and d["cpu_baseline"]["cpu_model"]
    assert "serial_step_s_scaled" in d["cpu_baseline"]["variants"]
    assert set(d["config"]) == {"workload", "N", "gmres", "l2", "operator"}


def test_reference_arm_does_not_load_the_product_libraries():
    """The CPU arm builds its case from the generator's arrays: neither liboptimet_b200.so nor the host layer may be
    mapped into its process (only oracle/ libraries)."""
    import subprocess
    code = ("import sys, json; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'small', '--steps', '1', "
            "'--warmup', '0']; import bench; bench.main(); "
            "maps = open('/proc/self/maps').read(); "
            "print('MAPPED', [l.split()[-1] for l in maps.splitlines() if 'liboptimet_b200' in l])")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "MAPPED []" in out.stdout
