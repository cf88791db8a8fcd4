"""ctypes view of include/optimet_b200.h (the drop-in C ABI)."""
import ctypes as C
import os

import numpy as np

OB_GMRES_ZCOMP = 1
OB_GMRES_BELOS = 2
OB_SOLVE_DIRECT = 3

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "liboptimet_b200.so")


class GmresOpts(C.Structure):
    _fields_ = [("flavour", C.c_int), ("tol", C.c_double), ("max_iters", C.c_int), ("restart", C.c_int),
                ("max_restarts", C.c_int)]


# every symbol include/optimet_b200.h declares
SYMBOLS = [
    "ob_create", "ob_destroy", "ob_last_error", "ob_device_info", "ob_partition", "ob_comm_unique_id",
    "ob_comm_init", "ob_set_cluster", "ob_set_frequency", "ob_set_incident", "ob_vtac", "ob_particle_factors",
    "ob_inc_local", "ob_assemble", "ob_release_matrix", "ob_fetch_block", "ob_fetch_matrix", "ob_matvec",
    "ob_source_ff", "ob_set_cg_tables", "ob_build_cg_tables", "ob_fetch_cg_table", "ob_source_sh", "ob_solve",
    "ob_unprecondition_ff", "ob_unprecondition_sh", "ob_run", "ob_cross_sections", "ob_timings", "ob_timer", "ob_set_option",
    "ob_measure_fp64_peak", "ob_dense_solve", "ob_aca_block", "ob_aca_compress", "ob_aca_stats", "ob_fields",
    "ob_set_shard", "ob_matvec_partial", "ob_host_register", "ob_host_unregister", "ob_create_multi", "ob_destroy_multi", "ob_multi_size", "ob_multi_ctx",
    "ob_multi_last_error", "ob_multi_set_cluster", "ob_multi_set_frequency", "ob_multi_set_incident", "ob_multi_set_option",
    "ob_multi_run",
]

_lib = None


def load():
    """Load liboptimet_b200.so; raises if the extension has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError("liboptimet_b200.so is missing: run `make` (or __graft_entry__.build()); "
                               "the B200 path has no CPU fallback")
        _lib = C.CDLL(path)
        _lib.ob_last_error.restype = C.c_char_p
        _lib.ob_last_error.argtypes = [C.c_void_p]
        _lib.ob_destroy.argtypes = [C.c_void_p]
        _lib.ob_destroy.restype = None
    return _lib


class Library:
    """Host-only entry points (no GPU needed)."""

    @staticmethod
    def partition(nobj, world, rank):
        f, n = C.c_int(), C.c_int()
        if load().ob_partition(int(nobj), int(world), int(rank), C.byref(f), C.byref(n)):
            raise ValueError("bad partition arguments")
        return f.value, n.value

    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        if load().ob_comm_unique_id(buf):
            raise RuntimeError("ncclGetUniqueId failed (libnccl.so.2 not loadable)")
        return buf.raw


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cz(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    if n is not None and a.size != n:
        raise ValueError("expected %d complex values, got %d" % (n, a.size))
    return a


def _c2(z):
    z = complex(z)
    return (C.c_double * 2)(z.real, z.imag)


class Context:
    """One GPU context (one per process/GPU)."""

    def __init__(self, device=0):
        self._lib = load()
        h = C.c_void_p()
        if self._lib.ob_create(int(device), C.byref(h)):
            raise RuntimeError(self._lib.ob_last_error(None).decode())
        self.h = h
        self.nobj = self.nMax = self.nMaxS = 0
        self.rank, self.world = 0, 1

    @classmethod
    def view(cls, handle, nobj=0, nMax=0, nMaxS=None):
        """Wrap an ob_ctx owned by someone else (the host solver); close() does not destroy it."""
        self = cls.__new__(cls)
        self._lib = load()
        self.h = C.c_void_p(handle)
        self._borrowed = True
        self.nobj, self.nMax, self.nMaxS = nobj, nMax, nMax if nMaxS is None else nMaxS
        self.rank, self.world = 0, 1
        return self

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "_borrowed", False):
                self._lib.ob_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self._lib.ob_last_error(self.h).decode())

    # sizes
    def n(self, harmonic=1):
        nm = self.nMax if harmonic == 1 else self.nMaxS
        return nm * (nm + 2)

    def N(self, harmonic=1):
        return 2 * self.n(harmonic) * self.nobj

    def local(self):
        return Library.partition(self.nobj, self.world, self.rank)

    def device_info(self):
        sm, fr, tot = C.c_int(), C.c_size_t(), C.c_size_t()
        self._chk(self._lib.ob_device_info(self.h, C.byref(sm), C.byref(fr), C.byref(tot)))
        return sm.value, fr.value, tot.value

    def comm_init(self, uid, rank, world):
        self._chk(self._lib.ob_comm_init(self.h, C.c_char_p(uid), int(rank), int(world)))
        self.rank, self.world = rank, world

    def set_cluster(self, xyz_m, radius_m, nMax, nMaxS=None):
        xyz = np.ascontiguousarray(xyz_m, dtype=np.float64).reshape(-1, 3)
        rad = np.ascontiguousarray(radius_m, dtype=np.float64).reshape(-1)
        nMaxS = nMax if nMaxS is None else nMaxS
        self._chk(self._lib.ob_set_cluster(self.h, int(xyz.shape[0]), _p(xyz), _p(rad), int(nMax), int(nMaxS)))
        self.nobj, self.nMax, self.nMaxS = xyz.shape[0], nMax, nMaxS

    def set_frequency(self, omega, waveK, eps_b, mu_b, eps, mu, eps_SH, mu_SH, ksippp, ksiparppar, gamma):
        arrs = [_cz(a, self.nobj) for a in (eps, mu, eps_SH, mu_SH, ksippp, ksiparppar, gamma)]
        self._chk(self._lib.ob_set_frequency(self.h, C.c_double(omega), _c2(waveK), _c2(eps_b), _c2(mu_b),
                                            *[_p(a) for a in arrs]))

    def set_incident(self, a, b):
        a, b = _cz(a, self.n(1)), _cz(b, self.n(1))
        self._chk(self._lib.ob_set_incident(self.h, _p(a), _p(b)))

    def vtac(self, relR_sph, k, regular_flag, nMax):
        n = nMax * (nMax + 2)
        A = np.zeros((n, n), dtype=np.complex128, order="F")
        B = np.zeros((n, n), dtype=np.complex128, order="F")
        self._chk(self._lib.ob_vtac(self.h, (C.c_double * 3)(*relR_sph), _c2(k), int(regular_flag), int(nMax), _p(A),
                                   _p(B)))
        return A, B

    def particle_factors(self, which):
        nm = self.nMax if which in (0, 4) else self.nMaxS
        out = np.zeros((self.nobj, 2 * nm * (nm + 2)), dtype=np.complex128)
        self._chk(self._lib.ob_particle_factors(self.h, int(which), _p(out)))
        return out

    def inc_local(self):
        out = np.zeros(self.N(1), dtype=np.complex128)
        self._chk(self._lib.ob_inc_local(self.h, _p(out)))
        return out

    def assemble(self, harmonic):
        self._chk(self._lib.ob_assemble(self.h, int(harmonic)))

    def release_matrix(self, harmonic):
        self._chk(self._lib.ob_release_matrix(self.h, int(harmonic)))

    def fetch_block(self, harmonic, i, j):
        b = 2 * self.n(harmonic)
        out = np.zeros((b, b), dtype=np.complex128, order="F")
        self._chk(self._lib.ob_fetch_block(self.h, int(harmonic), int(i), int(j), _p(out)))
        return out

    def fetch_matrix(self, harmonic):
        _, cnt = self.local()
        out = np.zeros((2 * self.n(harmonic) * cnt, self.N(harmonic)), dtype=np.complex128, order="F")
        self._chk(self._lib.ob_fetch_matrix(self.h, int(harmonic), _p(out)))
        return out

    def aca_block(self, harmonic, i, j):
        """Block (i, j) of the ACA-compressed operator: (rank, U, V, I, J); rank -1 -> dense block in U, 0 -> identity."""
        b = 2 * self.n(harmonic)
        U = np.zeros(b * b, dtype=np.complex128)
        V = np.zeros((b, b), dtype=np.complex128)
        I = np.zeros(b, dtype=np.int32)
        J = np.zeros(b, dtype=np.int32)
        r = C.c_int()
        self._chk(self._lib.ob_aca_block(self.h, int(harmonic), int(i), int(j), C.byref(r), _p(U), _p(V), _p(I), _p(J)))
        r = r.value
        if r < 0:
            return -1, U.reshape((b, b), order="F"), None, None, None
        if r == 0:
            return 0, None, None, None, None
        return r, U[:b * r].reshape((b, r), order="F"), V[:r].copy(), I[:r].copy(), J[:r].copy()

    def aca_compress(self, block):
        """ACA_compression of a caller-supplied square block on the device: U (dim x r), V (r x dim), I, J."""
        Cm = np.asfortranarray(block, dtype=np.complex128)
        b = Cm.shape[0]
        U = np.zeros(b * b, dtype=np.complex128)
        V = np.zeros((b, b), dtype=np.complex128)
        I = np.zeros(b, dtype=np.int32)
        J = np.zeros(b, dtype=np.int32)
        r = C.c_int()
        self._chk(self._lib.ob_aca_compress(self.h, int(b), _p(Cm), C.byref(r), _p(U), _p(V), _p(I), _p(J)))
        r = r.value
        return U[:b * r].reshape((b, r), order="F"), V[:r].copy(), I[:r].copy(), J[:r].copy()

    def aca_stats(self, harmonic):
        out = (C.c_double * 6)()
        self._chk(self._lib.ob_aca_stats(self.h, int(harmonic), out))
        return dict(stored_bytes=out[0], dense_bytes=out[1], lowrank_blocks=int(out[2]), dense_blocks=int(out[3]),
                    mean_rank=out[4], max_rank=int(out[5]))

    def fields(self, pts_sph, X_sca=None, X_int=None, X_sca_SH=None, X_int_SH=None, do_sh=True):
        """Result::setFields at spherical points (npts, 3): (npts, 4, 3) complex E_FF, H_FF, E_SH, H_SH (Cartesian
        components) and checkInner per point.  Vectors left None -> device-resident solution of the last run()."""
        pts = np.ascontiguousarray(pts_sph, dtype=np.float64).reshape(-1, 3)
        out = np.zeros((len(pts), 4, 3), dtype=np.complex128)
        inner = np.zeros(len(pts), dtype=np.int32)
        vecs = [None if v is None else _cz(v) for v in (X_sca, X_int, X_sca_SH, X_int_SH)]
        self._chk(self._lib.ob_fields(self.h, C.c_long(len(pts)), _p(pts), *[None if v is None else _p(v) for v in vecs],
                                      int(bool(do_sh)), _p(out), _p(inner)))
        return out, inner

    def set_shard(self, rank, world):
        """Plan / assemble as `rank` of `world` without a communicator (see ob_set_shard)."""
        self._chk(self._lib.ob_set_shard(self.h, int(rank), int(world)))
        self.rank, self.world = rank, world

    def matvec_partial(self, harmonic, x):
        """This shard's partial sums of the pair / rotated-axial operator (before the cross-rank sum)."""
        x = _cz(x, self.N(harmonic))
        acc = np.zeros_like(x)
        self._chk(self._lib.ob_matvec_partial(self.h, int(harmonic), _p(x), _p(acc)))
        return acc

    def matvec(self, harmonic, x):
        x = _cz(x, self.N(harmonic))
        y = np.zeros_like(x)
        self._chk(self._lib.ob_matvec(self.h, int(harmonic), _p(x), _p(y)))
        return y

    def source_ff(self):
        Q = np.zeros(self.N(1), dtype=np.complex128)
        self._chk(self._lib.ob_source_ff(self.h, _p(Q)))
        return Q

    def set_cg_tables(self, tables):
        t = [np.ascontiguousarray(x, dtype=np.float64) for x in tables]
        ptrs = (C.c_void_p * 9)(*[x.ctypes.data for x in t])
        self._chk(self._lib.ob_set_cg_tables(self.h, ptrs))

    def build_cg_tables(self):
        self._chk(self._lib.ob_build_cg_tables(self.h))

    def fetch_cg_table(self, t):
        out = np.zeros(self.n(2) * self.n(1) * self.n(1), dtype=np.float64)
        self._chk(self._lib.ob_fetch_cg_table(self.h, int(t), _p(out)))
        return out

    def source_sh(self, Xint_conj):
        x = _cz(Xint_conj, self.N(1))
        K = np.zeros(self.N(2), dtype=np.complex128)
        K1 = np.zeros(self.N(2), dtype=np.complex128)
        self._chk(self._lib.ob_source_sh(self.h, _p(x), _p(K), _p(K1)))
        return K, K1

    def solve(self, harmonic, rhs, opts):
        x = np.zeros(self.N(harmonic), dtype=np.complex128)
        it, rr = C.c_int(), C.c_double()
        r = None if rhs is None else _cz(rhs, self.N(harmonic))
        self._chk(self._lib.ob_solve(self.h, int(harmonic), None if r is None else _p(r), _p(x), C.byref(opts),
                                    C.byref(it), C.byref(rr)))
        return x, it.value, rr.value

    def dense_solve(self, A, b):
        """x = A^-1 b on the device (OB_SOLVE_DIRECT's LU on a caller-supplied matrix)."""
        A = np.asfortranarray(A, dtype=np.complex128)
        n = A.shape[0]
        if A.shape != (n, n):
            raise ValueError("square matrix expected")
        bb = _cz(b, n)
        x = np.zeros(n, dtype=np.complex128)
        self._chk(self._lib.ob_dense_solve(self.h, int(n), _p(A), _p(bb), _p(x)))
        return x

    def unprecondition_ff(self, X_sca):
        x = _cz(X_sca, self.N(1))
        out = np.zeros_like(x)
        self._chk(self._lib.ob_unprecondition_ff(self.h, _p(x), _p(out)))
        return out

    def unprecondition_sh(self, X_sca_SH, K1ana):
        x, k = _cz(X_sca_SH, self.N(2)), _cz(K1ana, self.N(2))
        out = np.zeros_like(x)
        self._chk(self._lib.ob_unprecondition_sh(self.h, _p(x), _p(k), _p(out)))
        return out

    def cross_sections(self, X_sca, X_int=None, X_sca_SH=None, X_int_SH=None):
        do_sh = X_sca_SH is not None
        xs = _cz(X_sca, self.N(1))
        args = [_p(xs)]
        keep = [xs]
        for v, h in ((X_int, 1), (X_sca_SH, 2), (X_int_SH, 2)):
            if do_sh:
                a = _cz(v, self.N(h))
                keep.append(a)
                args.append(_p(a))
            else:
                args.append(None)
        cs = (C.c_double * 5)()
        self._chk(self._lib.ob_cross_sections(self.h, *args, int(do_sh), cs))
        return dict(ext=cs[0], sca=cs[1], abs=cs[2], sca_SH=cs[3], abs_SH=cs[4])

    def run(self, opts, do_sh=True, fetch=True):
        """update() + solve() + cross sections for the current frequency, device resident."""
        outs = [np.zeros(self.N(1), dtype=np.complex128), np.zeros(self.N(1), dtype=np.complex128),
                np.zeros(self.N(2), dtype=np.complex128), np.zeros(self.N(2), dtype=np.complex128)]
        cs = (C.c_double * 5)()
        st = (C.c_int * 2)()
        ptrs = [_p(o) if fetch else None for o in outs]
        self._chk(self._lib.ob_run(self.h, C.byref(opts), int(do_sh), ptrs[0], ptrs[1], ptrs[2], ptrs[3], cs, st))
        res = dict(ext=cs[0], sca=cs[1], abs=cs[2], sca_SH=cs[3], abs_SH=cs[4], iters_ff=st[0], iters_sh=st[1])
        if fetch:
            res.update(X_sca=outs[0], X_int=outs[1], X_sca_SH=outs[2], X_int_SH=outs[3])
        return res

    def timings(self):
        t = (C.c_double * 16)()
        self._lib.ob_timings(self.h, t)
        names = ["factors_source", "assemble_ff", "solve_ff", "source_sh", "assemble_sh", "solve_sh", "cross_sections",
                 "matvec_ms", "matvec_count", "launches", "operator_bytes", "trace_reduce_ms", "trace_arnoldi_ms", "trace_gap_ms"]
        return {k: t[i] for i, k in enumerate(names)}

    def set_option(self, name, value):
        self._chk(self._lib.ob_set_option(self.h, name.encode(), C.c_double(value)))

    def measure_fp64_peak(self):
        v = C.c_double()
        self._chk(self._lib.ob_measure_fp64_peak(self.h, C.byref(v)))
        return v.value


class MultiContext:
    """ob_multi: one process, several GPUs (one context per device, one host worker thread per GPU)."""

    def __init__(self, devices):
        self._lib = load()
        self._lib.ob_multi_ctx.restype = C.c_void_p
        self._lib.ob_multi_last_error.restype = C.c_char_p
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        if self._lib.ob_create_multi(len(devices), devs, C.byref(h)):
            raise RuntimeError(self._lib.ob_last_error(None).decode())
        self.h = h
        self.size = self._lib.ob_multi_size(self.h)
        self.nobj = self.nMax = self.nMaxS = 0

    def close(self):
        if getattr(self, "h", None):
            self._lib.ob_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self._lib.ob_multi_last_error(self.h).decode())

    def ctx(self, rank):
        return Context.view(self._lib.ob_multi_ctx(self.h, int(rank)), self.nobj, self.nMax, self.nMaxS)

    def set_cluster(self, xyz_m, radius_m, nMax, nMaxS=None):
        xyz = np.ascontiguousarray(xyz_m, dtype=np.float64).reshape(-1, 3)
        rad = np.ascontiguousarray(radius_m, dtype=np.float64).reshape(-1)
        nMaxS = nMax if nMaxS is None else nMaxS
        self._chk(self._lib.ob_multi_set_cluster(self.h, int(xyz.shape[0]), _p(xyz), _p(rad), int(nMax), int(nMaxS)))
        self.nobj, self.nMax, self.nMaxS = xyz.shape[0], nMax, nMaxS

    def set_frequency(self, omega, waveK, eps_b, mu_b, eps, mu, eps_SH, mu_SH, ksippp, ksiparppar, gamma):
        arrs = [_cz(a, self.nobj) for a in (eps, mu, eps_SH, mu_SH, ksippp, ksiparppar, gamma)]
        self._chk(self._lib.ob_multi_set_frequency(self.h, C.c_double(omega), _c2(waveK), _c2(eps_b), _c2(mu_b),
                                                  *[_p(a) for a in arrs]))

    def set_incident(self, a, b):
        n = self.nMax * (self.nMax + 2)
        self._chk(self._lib.ob_multi_set_incident(self.h, _p(_cz(a, n)), _p(_cz(b, n))))

    def set_option(self, name, value):
        self._chk(self._lib.ob_multi_set_option(self.h, name.encode(), C.c_double(value)))

    def run(self, opts, do_sh=True):
        N1 = 2 * self.nMax * (self.nMax + 2) * self.nobj
        N2 = 2 * self.nMaxS * (self.nMaxS + 2) * self.nobj
        outs = [np.zeros(N1, dtype=np.complex128), np.zeros(N1, dtype=np.complex128),
                np.zeros(N2, dtype=np.complex128), np.zeros(N2, dtype=np.complex128)]
        cs = (C.c_double * 5)()
        st = (C.c_int * 2)()
        self._chk(self._lib.ob_multi_run(self.h, C.byref(opts), int(do_sh), _p(outs[0]), _p(outs[1]), _p(outs[2]),
                                        _p(outs[3]), cs, st))
        return dict(ext=cs[0], sca=cs[1], abs=cs[2], sca_SH=cs[3], abs_SH=cs[4], iters_ff=st[0], iters_sh=st[1],
                    X_sca=outs[0], X_int=outs[1], X_sca_SH=outs[2], X_int_SH=outs[3])
