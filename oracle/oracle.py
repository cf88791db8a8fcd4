"""ctypes binding of the CPU oracle (oracle/optimet_oracle.cpp).

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(optimet_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboptimet_oracle.so")
AMOS_PATH = os.path.join(HERE, "_ref", "libamos_ref.so")

_lib = None


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE] + (["-B"] if force else []))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_case_create.restype = C.c_void_p
        _lib.orc_case_error.restype = C.c_char_p
        _lib.orc_case_error.argtypes = [C.c_void_p]
        _lib.orc_bessel_calls.restype = C.c_long
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c2(z):
    z = complex(z)
    return (C.c_double * 2)(z.real, z.imag)


def have_amos():
    return os.path.exists(AMOS_PATH)


def set_bessel_backend(backend):
    """0 = own restatement, 1 = the reference's AMOS compiled into oracle/_ref."""
    rc = lib().orc_set_bessel_backend(int(backend), AMOS_PATH.encode())
    if rc != 0:
        raise RuntimeError("AMOS backend unavailable (oracle/_ref/libamos_ref.so missing)")


def set_threads(n):
    lib().orc_set_threads(int(n))


def set_as_shipped(dense_T, bessel_in_loops):
    lib().orc_set_as_shipped(int(dense_T), int(bessel_in_loops))


def bessel(kind, z, nmax):
    d = np.zeros(nmax + 1, dtype=np.complex128)
    dd = np.zeros(nmax + 1, dtype=np.complex128)
    rc = lib().orc_bessel(int(kind), _c2(z), int(nmax), _p(d), _p(dd))
    if rc:
        raise RuntimeError("oracle bessel failed")
    return d, dd


def ynm(the, phi, n, m):
    out = (C.c_double * 2)()
    lib().orc_ynm(C.c_double(the), C.c_double(phi), n, m, out)
    return complex(out[0], out[1])


def wigner(kind, js):
    arr = (C.c_int * len(js))(*[int(j) for j in js])
    out = C.c_double()
    lib().orc_wigner(kind, arr, C.byref(out))
    return out.value


def ta(R, k, regular, n, m, l, kk):
    out = (C.c_double * 2)()
    rc = lib().orc_ta((C.c_double * 3)(*R), _c2(k), int(regular), n, m, l, kk, out)
    if rc:
        raise RuntimeError("oracle ta failed")
    return complex(out[0], out[1])


def coupling(R, k, nMax, regular_flag=True):
    """Coupling(relR, k, nMax, regular_flag) -> (A, B), n x n with A[p, q] (Coupling.h:29-41)."""
    n = nMax * (nMax + 2)
    A = np.zeros((n, n), dtype=np.complex128, order="F")
    B = np.zeros((n, n), dtype=np.complex128, order="F")
    rc = lib().orc_coupling((C.c_double * 3)(*R), _c2(k), int(nMax), int(regular_flag), _p(A), _p(B))
    if rc:
        raise RuntimeError("oracle coupling failed")
    return A, B


def cg_tables(nMax, nMaxS):
    n, ns = nMax * (nMax + 2), nMaxS * (nMaxS + 2)
    T = np.zeros((9, ns * n * n), dtype=np.float64)
    ptrs = (C.c_void_p * 9)(*[T[i].ctypes.data for i in range(9)])
    rc = lib().orc_cg_tables(int(nMax), int(nMaxS), ptrs)
    if rc:
        raise RuntimeError("oracle cg_tables failed")
    return T


def aux_coefficients(R_sph, k, regular, nMax):
    """AuxCoefficients(R, waveK, regular, nMax) -> dict of M, N, Xm, Xp, each (n, 3) complex (Cartesian components)."""
    n = nMax * (nMax + 2)
    out = np.zeros((4, n, 3), dtype=np.complex128)
    lib().orc_aux_coefficients((C.c_double * 3)(*R_sph), _c2(k), int(bool(regular)), int(nMax), _p(out))
    return dict(M=out[0], N=out[1], Xm=out[2], Xp=out[3])


def grid_points(params):
    """OutputGrid::getPoint enumeration of Run::params = (x0, x1, nx, y0, y1, ny, z0, z1, nz) -> (npts, 3) spherical."""
    gp = (C.c_double * 9)(*[float(v) for v in params])
    n = int(params[2]) * int(params[5]) * int(params[8])
    out = np.zeros((n, 3), dtype=np.float64)
    lib().orc_grid_points(gp, _p(out))
    return out


def matvec(S, x):
    S = np.asfortranarray(S, dtype=np.complex128)
    x = np.ascontiguousarray(x, dtype=np.complex128)
    y = np.zeros(S.shape[0], dtype=np.complex128)
    lib().orc_matvec(_p(S), C.c_long(S.shape[0]), C.c_long(S.shape[1]), _p(x), _p(y))
    return y


SOLVER_DIRECT, SOLVER_ZCOMP, SOLVER_BELOS, SOLVER_ACA_ZCOMP = 0, 1, 2, 3


def set_eps_aca(eps):
    lib().orc_set_eps_aca(C.c_double(eps))


def aca_compress(block):
    """ACA_compression (PreconditionedMatrix.cpp:760-859) of a square block -> U (dim x r), V (r x dim), I, J."""
    Cm = np.asfortranarray(block, dtype=np.complex128)
    dim = Cm.shape[0]
    U = np.zeros((dim, dim), dtype=np.complex128, order="F")
    V = np.zeros((dim, dim), dtype=np.complex128)
    I = np.zeros(dim, dtype=np.int32)
    J = np.zeros(dim, dtype=np.int32)
    r = C.c_int()
    rc = lib().orc_aca_compress(_p(Cm), int(dim), C.byref(r), _p(U), _p(V), _p(I), _p(J))
    if rc:
        raise RuntimeError("ACA_compression failed (no admissible pivot)")
    r = r.value
    return np.asfortranarray(U.ravel(order="F")[:dim * r].reshape((dim, r), order="F")), V[:r].copy(), I[:r].copy(), J[:r].copy()


def solve_dense(S, rhs, solver, tol=1e-6, maxit=240, restart=30, max_restarts=2):
    S = np.asfortranarray(S, dtype=np.complex128)
    rhs = np.ascontiguousarray(rhs, dtype=np.complex128)
    x = np.zeros_like(rhs)
    it = C.c_int()
    rr = C.c_double()
    opts = (C.c_double * 4)(tol, maxit, restart, max_restarts)
    rc = lib().orc_solve_dense(_p(S), C.c_long(S.shape[0]), _p(rhs), int(solver), opts, _p(x), C.byref(it), C.byref(rr))
    if rc:
        raise RuntimeError("Error encountered while solving the linear system")
    return x, it.value, rr.value


MODEL_FIXED, MODEL_GOLD, MODEL_SILICON = 0, 3, 4


class Case:
    """One simulation case: geometry + excitation, as Simulation::scan_wavelengths drives it."""

    def __init__(self):
        self.h = C.c_void_p(lib().orc_case_create())
        self.nMax = None

    def __del__(self):
        try:
            lib().orc_case_destroy(self.h)
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise RuntimeError(lib().orc_case_error(self.h).decode())

    def add_sphere(self, xyz_m, radius_m, nMax, model, params, nMaxS=None):
        p = np.asarray(params, dtype=np.float64)
        self._chk(lib().orc_case_add_sphere(self.h, (C.c_double * 3)(*xyz_m), C.c_double(radius_m), int(nMax),
                                           int(nMax if nMaxS is None else nMaxS), int(model), _p(p)))
        self.nMax = nMax

    def set_background(self, eps, mu):
        lib().orc_case_set_background(self.h, _c2(eps), _c2(mu))

    def set_source(self, wavelength_m, theta, phi, Eth, Eph, SH_cond, nMax=None):
        self._chk(lib().orc_case_set_source(self.h, C.c_double(wavelength_m), C.c_double(theta), C.c_double(phi),
                                           _c2(Eth), _c2(Eph), int(SH_cond), int(self.nMax if nMax is None else nMax)))

    def update_wavelength(self, lam_m):
        self._chk(lib().orc_case_update_wavelength(self.h, C.c_double(lam_m)))

    def info(self):
        nobj, nMax, nMaxS = C.c_int(), C.c_int(), C.c_int()
        om = C.c_double()
        k = (C.c_double * 2)()
        lib().orc_case_info(self.h, C.byref(nobj), C.byref(nMax), C.byref(nMaxS), C.byref(om), k)
        return dict(nobj=nobj.value, nMax=nMax.value, nMaxS=nMaxS.value, omega=om.value, waveK=complex(k[0], k[1]))

    def material(self, j):
        out = np.zeros(6, dtype=np.complex128)
        lib().orc_case_material(self.h, int(j), _p(out))
        return dict(eps_r=out[0], eps_r_SH=out[1], ksippp=out[2], ksiparppar=out[3], gamma=out[4], mu_r=out[5])

    def incident(self):
        n = self.info()["nMax"]
        n = n * (n + 2)
        a = np.zeros(n, dtype=np.complex128)
        b = np.zeros(n, dtype=np.complex128)
        lib().orc_case_incident(self.h, _p(a), _p(b))
        return a, b

    def particle_factors(self, j, which):
        i = self.info()
        nm = i["nMax"] if which in (0, 4) else i["nMaxS"]
        out = np.zeros(2 * nm * (nm + 2), dtype=np.complex128)
        self._chk(lib().orc_case_particle_factors(self.h, int(j), int(which), _p(out)))
        return out

    def matrix(self, harmonic, i0=0, i1=None):
        i = self.info()
        nm = i["nMax"] if harmonic == 1 else i["nMaxS"]
        n2 = 2 * nm * (nm + 2)
        if i1 is None:
            i1 = i["nobj"]
        S = np.zeros((n2 * (i1 - i0), n2 * i["nobj"]), dtype=np.complex128, order="F")
        self._chk(lib().orc_case_matrix(self.h, int(harmonic), int(i0), int(i1), _p(S)))
        return S

    def source(self):
        i = self.info()
        Q = np.zeros(2 * i["nMax"] * (i["nMax"] + 2) * i["nobj"], dtype=np.complex128)
        self._chk(lib().orc_case_source(self.h, _p(Q)))
        return Q

    def inc_local(self, j):
        i = self.info()
        out = np.zeros(2 * i["nMax"] * (i["nMax"] + 2), dtype=np.complex128)
        self._chk(lib().orc_case_inc_local(self.h, int(j), _p(out)))
        return out

    def sh_source(self, Xint_conj):
        i = self.info()
        x = np.ascontiguousarray(Xint_conj, dtype=np.complex128)
        n = 2 * i["nMaxS"] * (i["nMaxS"] + 2) * i["nobj"]
        K = np.zeros(n, dtype=np.complex128)
        K1 = np.zeros(n, dtype=np.complex128)
        self._chk(lib().orc_case_sh_source(self.h, _p(x), _p(K), _p(K1)))
        return K, K1

    def solve(self, solver=SOLVER_DIRECT, tol=1e-6, maxit=240, restart=30, max_restarts=2):
        opts = (C.c_double * 4)(tol, maxit, restart, max_restarts)
        self._chk(lib().orc_case_solve(self.h, int(solver), opts))

    def aca_block(self, harmonic, i, j):
        """Block (i, j) of Scattering_matrix_ACA_FF/_SH: (rank, U, V, I, J); rank -1 = dense block in U."""
        inf = self.info()
        nm = inf["nMax"] if harmonic == 1 else inf["nMaxS"]
        dim = 2 * nm * (nm + 2)
        U = np.zeros(dim * dim, dtype=np.complex128)
        V = np.zeros((dim, dim), dtype=np.complex128)
        I = np.zeros(dim, dtype=np.int32)
        J = np.zeros(dim, dtype=np.int32)
        r = C.c_int()
        self._chk(lib().orc_case_aca_block(self.h, int(harmonic), int(i), int(j), C.byref(r), _p(U), _p(V), _p(I), _p(J)))
        r = r.value
        if r < 0:
            return -1, U.reshape((dim, dim), order="F"), None, None, None
        return r, U[:dim * r].reshape((dim, r), order="F"), V[:r].copy(), I[:r].copy(), J[:r].copy()

    def force_aca_pivots(self, harmonic, i, j, I, J):
        """Tests: impose the pivot rows / columns of block (i, j) on later ACA runs (None clears)."""
        if I is None:
            lib().orc_case_force_aca_pivots(self.h, int(harmonic), int(i), int(j), 0, None, None)
            return
        I = np.ascontiguousarray(I, dtype=np.int32)
        J = np.ascontiguousarray(J, dtype=np.int32)
        lib().orc_case_force_aca_pivots(self.h, int(harmonic), int(i), int(j), int(I.size), _p(I), _p(J))

    def aca_ranks(self, harmonic):
        nobj = self.info()["nobj"]
        out = np.zeros(nobj * nobj, dtype=np.int32)
        lib().orc_case_aca_ranks(self.h, int(harmonic), _p(out))
        return out.reshape(nobj, nobj)

    def matvec_aca(self, harmonic, x):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        y = np.zeros_like(x)
        self._chk(lib().orc_case_matvec_aca(self.h, int(harmonic), _p(x), _p(y)))
        return y

    def fields(self, pts_sph):
        """Result::setFields at spherical points (npts, 3) on the current solution vectors:
        (npts, 4, 3) complex = E_FF, H_FF, E_SH, H_SH Cartesian components, and checkInner per point."""
        pts = np.ascontiguousarray(pts_sph, dtype=np.float64).reshape(-1, 3)
        out = np.zeros((len(pts), 4, 3), dtype=np.complex128)
        inner = np.zeros(len(pts), dtype=np.int32)
        self._chk(lib().orc_case_fields(self.h, C.c_long(len(pts)), _p(pts), _p(out), _p(inner)))
        return out, inner

    def coeff_part_sh(self, obj, r):
        nm = self.info()["nMaxS"]
        a = np.zeros(nm * (nm + 2), dtype=np.complex128)
        b = np.zeros_like(a)
        self._chk(lib().orc_case_coeff_part_sh(self.h, int(obj), C.c_double(r), _p(a), _p(b)))
        return a, b

    def vector(self, which):
        i = self.info()
        nm = i["nMax"] if which in (0, 1, 4) else i["nMaxS"]
        out = np.zeros(2 * nm * (nm + 2) * i["nobj"], dtype=np.complex128)
        lib().orc_case_get_vector(self.h, int(which), _p(out))
        return out

    def set_vector(self, which, v):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        lib().orc_case_set_vector(self.h, int(which), _p(v), C.c_long(v.size))

    def iters(self):
        a, b = C.c_int(), C.c_int()
        ra, rb = C.c_double(), C.c_double()
        lib().orc_case_iters(self.h, C.byref(a), C.byref(b), C.byref(ra), C.byref(rb))
        return a.value, b.value, ra.value, rb.value

    def cross_sections(self):
        out = (C.c_double * 5)()
        self._chk(lib().orc_case_cross_sections(self.h, out))
        return dict(ext=out[0], sca=out[1], abs_direct=out[2], sca_SH=out[3], abs_SH=out[4])
