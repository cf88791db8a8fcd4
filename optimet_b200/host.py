"""ctypes view of the C++11 host layer (optimet_b200/host): XML reader, material models, plane-wave
coefficients and the solver::B200Matrix adaptor.  Host-only calls (load, info, arrays) work without a
GPU; creating a Solver needs one (no CPU fallback)."""
import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

HOST_SYMBOLS = ["obh_load_xml", "obh_load_xml_string", "obh_free", "obh_error", "obh_info", "obh_set_wavelength",
                "obh_get_arrays", "obh_gmres_defaults", "obh_solver_create", "obh_solver_free", "obh_solver_error",
                "obh_solver_ctx", "obh_solver_comm", "obh_solver_set_gmres", "obh_solver_set_aca_mode", "obh_solver_step", "obh_scan",
                "obh_field_simulation", "obh_grid_points", "obh_solver_vectors"]


def load():
    global _lib
    if _lib is None:
        capi.load()  # dependency, resolved through $ORIGIN rpath as well
        path = os.path.join(_HERE, "liboptimet_b200_host.so")
        if not os.path.exists(path):
            raise RuntimeError("liboptimet_b200_host.so is missing: run `make`")
        _lib = C.CDLL(path)
        for f in ("obh_load_xml", "obh_load_xml_string", "obh_solver_create", "obh_solver_ctx"):
            getattr(_lib, f).restype = C.c_void_p
        for f in ("obh_error", "obh_solver_error"):
            getattr(_lib, f).restype = C.c_char_p
            getattr(_lib, f).argtypes = [C.c_void_p]
        _lib.obh_free.argtypes = [C.c_void_p]
        _lib.obh_solver_free.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Case:
    """A parsed OPTIMET input (Run: Geometry + Excitation + output request)."""

    def __init__(self, path=None, xml=None):
        lib = load()
        err = C.create_string_buffer(1024)
        if path is not None:
            h = lib.obh_load_xml(str(path).encode(), err, 1024)
        else:
            h = lib.obh_load_xml_string(xml.encode(), err, 1024)
        if not h:
            raise RuntimeError(err.value.decode())
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            if self.h:
                load().obh_free(self.h)
        except Exception:
            pass

    def info(self):
        info = (C.c_int * 6)()
        params = (C.c_double * 9)()
        lam = C.c_double()
        load().obh_info(self.h, info, params, C.byref(lam))
        return dict(nobj=info[0], nMax=info[1], nMaxS=info[2], SH_cond=bool(info[3]), outputType=info[4],
                    ACA_cond=bool(info[5]), params=list(params), wavelength=lam.value)

    def set_wavelength(self, lam_m):
        if load().obh_set_wavelength(self.h, C.c_double(lam_m)):
            raise RuntimeError(load().obh_error(self.h).decode())

    def arrays(self):
        i = self.info()
        nobj, n = i["nobj"], i["nMax"] * (i["nMax"] + 2)
        xyz = np.zeros((nobj, 3))
        radius = np.zeros(nobj)
        mats = np.zeros((7, nobj), dtype=np.complex128)
        a = np.zeros(n, dtype=np.complex128)
        b = np.zeros(n, dtype=np.complex128)
        scal = (C.c_double * 7)()
        load().obh_get_arrays(self.h, _p(xyz), _p(radius), _p(mats), _p(a), _p(b), scal)
        return dict(xyz=xyz, radius=radius, eps=mats[0], mu=mats[1], eps_SH=mats[2], mu_SH=mats[3], ksippp=mats[4],
                    ksiparppar=mats[5], gamma=mats[6], a=a, b=b, omega=scal[0], waveK=complex(scal[1], scal[2]),
                    eps_b=complex(scal[3], scal[4]), mu_b=complex(scal[5], scal[6]))

    def gmres_defaults(self):
        o = capi.GmresOpts()
        load().obh_gmres_defaults(self.h, C.byref(o))
        return o

    def grid_points(self):
        """OutputGrid::getPoint enumeration of the case's field grid: (npts, 3) spherical (r, theta, phi)."""
        p = self.info()["params"]
        npts = int(p[2]) * int(p[5]) * int(p[8])
        pts = np.zeros((npts, 3), dtype=np.float64)
        n = C.c_long()
        if load().obh_grid_points(self.h, _p(pts), C.c_long(npts), C.byref(n)):
            raise RuntimeError("obh_grid_points failed")
        return pts

    def scan_wavelengths_list(self):
        """Wavelengths of <scan><wavelength .../> (Simulation.cpp:634-644)."""
        p = self.info()["params"]
        steps = int(p[2])
        if steps <= 1:
            return [p[0]]
        lams = (p[1] - p[0]) / (steps - 1)
        return [p[0] + i * lams for i in range(steps)]


class Solver:
    """solver::B200Matrix -- the drop-in for the reference's AbstractSolver (needs a GPU)."""

    def __init__(self, case, device=0):
        lib = load()
        err = C.create_string_buffer(1024)
        s = lib.obh_solver_create(case.h, int(device), err, 1024)
        if not s:
            raise RuntimeError(err.value.decode())
        self.s = C.c_void_p(s)
        self.case = case

    def close(self):
        if getattr(self, "s", None):
            load().obh_solver_free(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise RuntimeError(load().obh_solver_error(self.s).decode())

    def comm_init(self, uid, rank, world):
        self._chk(load().obh_solver_comm(self.s, C.c_char_p(uid), int(rank), int(world)))

    def set_gmres(self, opts):
        load().obh_solver_set_gmres(self.s, C.byref(opts))

    def field_simulation(self, case_file=None):
        """Simulation::field_simulation on the case's <output type="field"> grid: dict with E_FF, H_FF, E_SH, H_SH
        (npts, 3) Cartesian components in OutputGrid order (x fastest), inner (npts,), dims (nx, ny, nz)."""
        p = self.case.info()["params"]
        npts = int(p[2]) * int(p[5]) * int(p[8])
        out = np.zeros((npts, 4, 3), dtype=np.complex128)
        inner = np.zeros(npts, dtype=np.int32)
        dims = (C.c_int * 3)()
        self._chk(load().obh_field_simulation(self.s, self.case.h, None if case_file is None else case_file.encode(),
                                             _p(out), _p(inner), C.c_long(npts), dims))
        return dict(E_FF=out[:, 0], H_FF=out[:, 1], E_SH=out[:, 2], H_SH=out[:, 3], inner=inner, dims=tuple(dims))

    def set_aca_mode(self, mode):
        """-1 follow <ACA compression> (default), 0 never compress, 1 always compress."""
        load().obh_solver_set_aca_mode(self.s, int(mode))

    def ctx(self):
        """The ob_ctx of this solver wrapped as a capi.Context view (not owned)."""
        i = self.case.info()
        return capi.Context.view(load().obh_solver_ctx(self.s), i["nobj"], i["nMax"], i["nMaxS"])

    def ctx_timings(self):
        t = (C.c_double * 16)()
        capi.load().ob_timings(C.c_void_p(load().obh_solver_ctx(self.s)), t)
        names = ["factors_source", "assemble_ff", "solve_ff", "source_sh", "assemble_sh", "solve_sh", "cross_sections",
                 "matvec_ms", "matvec_count", "launches", "operator_bytes", "trace_reduce_ms", "trace_arnoldi_ms", "trace_gap_ms"]
        return {k: t[i] for i, k in enumerate(names)}

    def set_option(self, name, value):
        rc = capi.load().ob_set_option(C.c_void_p(load().obh_solver_ctx(self.s)), name.encode(), C.c_double(value))
        if rc:
            raise RuntimeError("ob_set_option failed")

    def step(self, lam_m=-1.0, fetch=True):
        i = self.case.info()
        N1 = 2 * i["nMax"] * (i["nMax"] + 2) * i["nobj"]
        N2 = 2 * i["nMaxS"] * (i["nMaxS"] + 2) * i["nobj"]
        view = fetch == "view"   # results stay in the solver's own page-locked vectors; numpy views, valid until the next step
        outs = [np.zeros(N1, dtype=np.complex128), np.zeros(N1, dtype=np.complex128),
                np.zeros(N2, dtype=np.complex128), np.zeros(N2, dtype=np.complex128)] if (fetch and not view) else [None] * 4
        cs = (C.c_double * 5)()
        it = (C.c_int * 2)()
        self._chk(load().obh_solver_step(self.s, self.case.h, C.c_double(lam_m), _p(outs[0]), _p(outs[1]), _p(outs[2]),
                                        _p(outs[3]), cs, it))
        res = dict(ext=cs[0], sca=cs[1], abs=cs[2], sca_SH=cs[3], abs_SH=cs[4], iters_ff=it[0], iters_sh=it[1])
        if view:
            ptrs = (C.POINTER(C.c_double) * 4)()
            sizes = (C.c_long * 4)()
            load().obh_solver_vectors(self.s, ptrs, sizes)
            outs = [np.ctypeslib.as_array(ptrs[k], shape=(2 * sizes[k],)).view(np.complex128) if sizes[k] else
                    np.zeros(0, dtype=np.complex128) for k in range(4)]
        if fetch:
            res.update(X_sca=outs[0], X_int=outs[1], X_sca_SH=outs[2], X_int_SH=outs[3])
        return res

    def scan(self, case_file=None, maxlines=4096):
        lines = np.zeros((maxlines, 8))
        n = C.c_int()
        self._chk(load().obh_scan(self.s, self.case.h, None if case_file is None else str(case_file).encode(),
                                 _p(lines), maxlines, C.byref(n)))
        keys = ["lambda", "abs_FF", "sca_FF", "sca_SH", "abs_SH", "ext_FF", "iters_FF", "iters_SH"]
        return [dict(zip(keys, lines[i])) for i in range(n.value)]
