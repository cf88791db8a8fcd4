"""The axial-only assembly of the rotated-axial operator form, checked on the CPU from the very source the device warp
runs: optimet_b200/csrc/ob_rot_axial.cuh is `__host__ __device__`; tests/rot_axial_host.cpp compiles it with g++ and
runs one pair with lane 0 of 1 (a warp of one thread executes the items of every level in order, which is a valid
schedule of the warp-parallel loops).  Compared with the oracle's Coupling at theta = 0 and, when present, with the
reference's own compiled Coupling.  No GPU needed; the kernel wrapper itself (k_assemble_axial_only) is checked on the
GPU by tests/test_gpu_rot.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from oracle import reference_build as RB

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="CUDA headers absent")


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("rot_axial") / "librot_axial_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I" + CUDA_INC,
                           os.path.join(ROOT, "tests", "rot_axial_host.cpp"), "-o", out])
    return C.CDLL(out)


def _flat(n, m):
    return n * (n + 1) - m - 1


def _offX(NM, mu):
    return sum((NM - max(u, 1) + 1) ** 2 for u in range(mu))


@pytest.mark.parametrize("NM,k,d", [(1, 2 * np.pi / 800e-9, 200e-9), (6, 2 * np.pi / 800e-9 * (1.2 + 0.05j), 300e-9),
                                    (8, 2 * np.pi / 800e-9, 190e-9), (10, 2 * np.pi / 400e-9, 160e-9),
                                    (13, 2 * np.pi / 400e-9 * (1.0 + 0.02j), 700e-9)])
def test_device_source_of_the_axial_recursion_on_the_host(host_lib, NM, k, d):
    X = _offX(NM, NM + 1)
    A = np.zeros(X, dtype=np.complex128)
    B = np.zeros(X, dtype=np.complex128)
    kk = (C.c_double * 2)(complex(k).real, complex(k).imag)
    got = host_lib.rot_axial_host(int(NM), kk, C.c_double(d), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p))
    assert got == X
    refs = [O.coupling([d, 0.0, 0.0], k, NM, True)]
    if RB.have():
        refs.append(RB.coupling([d, 0.0, 0.0], k, NM, True))  # the reference's own compiled Coupling
    for Az, Bz in refs:
        sa, sb = np.abs(Az).max(), np.abs(Bz).max()
        for mu in range(NM + 1):
            n0 = max(mu, 1)
            w = NM - n0 + 1
            for n in range(n0, NM + 1):
                for l in range(n0, NM + 1):
                    e = _offX(NM, mu) + (n - n0) * w + (l - n0)
                    assert abs(A[e] - Az[_flat(n, mu), _flat(l, mu)]) < 1e-12 * sa, (mu, n, l)
                    assert abs(B[e] - Bz[_flat(n, mu), _flat(l, mu)]) < 1e-12 * sb, (mu, n, l)


@pytest.mark.parametrize("NM", [1, 4, 10, 13])
def test_combined_output_of_the_axial_recursion(host_lib, NM):
    """combine = 1 / 2 (2 is what k_assemble_axial_only passes): A + B for every mu, A - B for mu >= 1 behind the mu = 0 block."""
    k, d = 2 * np.pi / 500e-9 * (1.1 + 0.03j), 230e-9
    X = _offX(NM, NM + 1)
    A = np.zeros(X, dtype=np.complex128)
    B = np.zeros(X, dtype=np.complex128)
    Cp = np.zeros(X, dtype=np.complex128)
    Cm = np.zeros(X - NM * NM, dtype=np.complex128)
    kk = (C.c_double * 2)(complex(k).real, complex(k).imag)
    host_lib.rot_axial_host(int(NM), kk, C.c_double(d), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p))
    host_lib.rot_axial_host_combined(int(NM), kk, C.c_double(d), Cp.ctypes.data_as(C.c_void_p),
                                     Cm.ctypes.data_as(C.c_void_p), 1)
    assert np.array_equal(Cp, A + B)
    assert np.array_equal(Cm, (A - B)[NM * NM:])
    # combine = 2 (the record layout of ob_rot.cu): the same values as [Re | Im] planes of doubles
    Pp = np.zeros(2 * X)
    Pm = np.zeros(2 * (X - NM * NM))
    host_lib.rot_axial_host_combined(int(NM), kk, C.c_double(d), Pp.ctypes.data_as(C.c_void_p),
                                     Pm.ctypes.data_as(C.c_void_p), 2)
    # ... in FRAGMENT ORDER and written through the transpose symmetry (n, l) = (-1)^(n + l) (l, n): the value (n, l)
    # of the plain layout sits, with that sign, where the fragment order keeps entry (l, n)
    perm = np.zeros(X, dtype=np.int64)
    sign = np.zeros(X)
    tr = np.zeros(X, dtype=np.int64)
    for mu in range(NM + 1):
        n0 = max(mu, 1)
        w = NM - n0 + 1
        for n in range(n0, NM + 1):
            for l in range(n0, NM + 1):
                e = _offX(NM, mu) + (n - n0) * w + (l - n0)
                perm[e] = host_lib.rot_cidx_host(NM, mu, l, n)
                sign[e] = (-1.0) ** (n + l)
                tr[e] = _offX(NM, mu) + (l - n0) * w + (n - n0)
    assert sorted(perm) == list(range(X))
    for mu in range(NM + 1):                  # every order keeps its own block of the plane
        blk = perm[_offX(NM, mu):_offX(NM, mu + 1)]
        assert blk.min() == _offX(NM, mu) and blk.max() == _offX(NM, mu + 1) - 1
    permM = perm[NM * NM:] - NM * NM
    assert np.array_equal((Pp[:X] + 1j * Pp[X:])[perm], sign * Cp)
    assert np.array_equal((Pm[:X - NM * NM] + 1j * Pm[X - NM * NM:])[permM], sign[NM * NM:] * Cm)
    # the symmetry itself, on the values of the plain layout
    assert np.abs(sign * Cp - Cp[tr]).max() < 1e-13 * np.abs(Cp).max()
    CmF = np.concatenate([np.zeros(NM * NM), Cm])
    assert np.abs(sign * CmF - CmF[tr]).max() < 1e-13 * np.abs(Cm).max() if NM > 1 else True
    # the tabulated path the kernel runs (index-only coefficients from rot_axial_tables_build): bit-identical
    Tp = np.zeros(2 * X)
    Tm = np.zeros(2 * (X - NM * NM))
    host_lib.rot_axial_host_tabulated(int(NM), kk, C.c_double(d), Tp.ctypes.data_as(C.c_void_p), Tm.ctypes.data_as(C.c_void_p))
    assert np.array_equal(Tp, Pp) and np.array_equal(Tm, Pm)
    assert not np.abs(B[:NM * NM]).any()      # mu = 0: B = 0, so Cm = Cp there and is not stored


def _lane_reads(rows, K, index, first_row=0):
    """Addresses the apply kernel's lanes read for every (row tile, K step) of a matrix in fragment order, as
    k_matvec_rot computes them (rot_dchains / rot_p2_unit): full steps at tile base + 4 Rt s + 4 fr + fc, a last step
    of Kv < 4 entries at tile base + 4 Rt (ks - 1) + Kv fr + fc; `first_row` = 1 is the a class (phantom row 0)."""
    ks = (K + 3) // 4
    for m0 in range(0, rows + first_row, 8):
        first = max(m0, first_row)
        last = min(m0 + 7, rows - 1 + first_row)
        Rt = last - first + 1
        radj = first - m0          # 1 for tile 0 of the a class
        base = (first - first_row) * K
        for s in range(ks):
            kv = min(4, K - 4 * s)
            for lane in range(32):
                fr, fc = lane >> 2, lane & 3
                row, kcol = m0 + fr, 4 * s + fc
                if row < first or row > last or kcol >= K:
                    continue       # garbage rows are never stored, K-tail entries are zeroed in registers
                if kv == 4:
                    addr = base + 4 * Rt * s + 4 * (fr - radj) + fc
                else:
                    addr = base + 4 * Rt * s + kv * (fr - radj) + fc
                assert addr == index(row, kcol), (rows, K, m0, s, lane)
                yield m0, s, lane, addr


@pytest.mark.parametrize("K", list(range(1, 15)))
def test_fragment_order_of_the_record_matrices(host_lib, K):
    """Each DMMA A fragment (8 rows x 4 K entries, compacted to the valid ones) is one contiguous run in lane order:
    a half-warp reads <= 16 consecutive doubles (no shared-memory bank conflict), the whole matrix is a permutation of
    [0, rows K) (no padding), and the kernel's address arithmetic lands on the element the lane must hold."""
    for first_row, index in ((0, lambda r, k: host_lib.rot_frag_index_host(K, K, r, k)),
                             (1, lambda r, k: host_lib.rot_frag_index_a_host(K, r, k))):
        seen = {}
        for m0, s, lane, addr in _lane_reads(K, K, index, first_row):
            seen.setdefault((m0, s), []).append((lane, addr))
        allad = sorted(a for v in seen.values() for _, a in v)
        assert allad == list(range(K * K))
        for (m0, s), v in seen.items():
            ads = [a for _, a in sorted(v)]
            assert ads == list(range(ads[0], ads[0] + len(ads)))          # contiguous, in lane order
            for half in (0, 1):                                           # 8-byte banks of a half-warp: all distinct
                hb = [a % 16 for l, a in v if (l >> 4) == half]
                assert len(hb) == len(set(hb))
