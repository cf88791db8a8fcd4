"""ctypes binding of oracle/_ref/libpath_ref.so: the reference's OWN translation units (TranslationAdditionCoefficients,
Coupling, Scatterer, ElectroMagnetic, AuxCoefficients, Excitation, ...) compiled from /root/reference where they lie by
`make -C oracle ref_path`, behind stand-ins for the libraries the image lacks (oracle/stub/: a container-only
<Eigen/Core>, the cmake-generated Types.h, Boost.Math's spherical_harmonic / factorial, two CBLAS calls).

TEST INFRASTRUCTURE ONLY, like oracle.py.  The library travels to the GPU box as a built file (oracle/_ref/ is
git-ignored, not gpurun-ignored); nothing here reads /root/reference at run time."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libpath_ref.so")
_lib = None


def have():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c2(z):
    z = complex(z)
    return (C.c_double * 2)(z.real, z.imag)


def coupling(R_sph, k, nMax, regular_flag=True):
    """optimet::Coupling(relR, k, nMax, regular_flag) -> (diagonal A, offdiagonal B), n x n."""
    n = nMax * (nMax + 2)
    A = np.zeros((n, n), dtype=np.complex128, order="F")
    B = np.zeros_like(A)
    lib().ref_coupling((C.c_double * 3)(*R_sph), _c2(k), int(nMax), int(bool(regular_flag)), _p(A), _p(B))
    return A, B


def material(model, params, wavelength_m):
    """ElectroMagnetic after init + update(lambda): eps_r, eps_r_SH, ksippp, ksiparppar, gamma, mu_r."""
    prm = np.asarray(params, dtype=np.float64)
    out = np.zeros(6, dtype=np.complex128)
    lib().ref_material(int(model), _p(prm), C.c_double(wavelength_m), _p(out))
    return dict(zip(("eps_r", "eps_r_SH", "ksippp", "ksiparppar", "gamma", "mu_r"), out))


def particle_factors(model, params, radius_m, nMax, wavelength_m, which, background=(1.0, 1.0), nMaxS=None):
    """Scatterer::getTLocal (0), getTLocalSH (1), getTLocalSH1_outer (2), getTLocalSH2_outer (3), getIaux (4),
    getIauxSH1 (5), getIauxSH2 (6): 2n complex."""
    prm = np.asarray(params, dtype=np.float64)
    nS = nMax if nMaxS is None else nMaxS
    nm = nMax if which in (0, 4) else nS
    out = np.zeros(2 * nm * (nm + 2), dtype=np.complex128)
    lib().ref_particle_factors(int(model), _p(prm), C.c_double(radius_m), int(nMax), int(nS), C.c_double(wavelength_m),
                               _c2(background[0]), _c2(background[1]), int(which), _p(out))
    return out


def aux_coefficients(R_sph, k, regular, nMax):
    n = nMax * (nMax + 2)
    out = np.zeros((4, n, 3), dtype=np.complex128)
    lib().ref_aux_coefficients((C.c_double * 3)(*R_sph), _c2(k), int(bool(regular)), int(nMax), _p(out))
    return dict(M=out[0], N=out[1], Xm=out[2], Xp=out[3])


def excitation(wavelength_m, theta, phi, Eth, Eph, nMax, background=(1.0, 1.0)):
    """Excitation as Reader.cpp:800-832 builds it -> (dataIncAp, dataIncBp, waveK)."""
    n = nMax * (nMax + 2)
    a = np.zeros(n, dtype=np.complex128)
    b = np.zeros(n, dtype=np.complex128)
    wk = (C.c_double * 2)()
    lib().ref_excitation(C.c_double(wavelength_m), C.c_double(theta), C.c_double(phi), _c2(Eth), _c2(Eph),
                         _c2(background[0]), _c2(background[1]), int(nMax), _p(a), _p(b), wk)
    return a, b, complex(wk[0], wk[1])


def inc_local(wavelength_m, theta, phi, Eth, Eph, nMax, R_sph, background=(1.0, 1.0)):
    """Excitation::getIncLocal at the spherical point R: 2n complex."""
    out = np.zeros(2 * nMax * (nMax + 2), dtype=np.complex128)
    lib().ref_inc_local(C.c_double(wavelength_m), C.c_double(theta), C.c_double(phi), _c2(Eth), _c2(Eph),
                        _c2(background[0]), _c2(background[1]), int(nMax), (C.c_double * 3)(*R_sph), _p(out))
    return out


def relative_position(xyz_i, xyz_j):
    """vR_i - vR_j as the reference forms it (Tools::toSpherical of both, Spherical::operator-): (r, theta, phi)."""
    out = (C.c_double * 3)()
    lib().ref_relative_position((C.c_double * 3)(*xyz_i), (C.c_double * 3)(*xyz_j), out)
    return [out[0], out[1], out[2]]


def ynm(n, m, theta, phi):
    out = (C.c_double * 2)()
    lib().ref_ynm(int(n), int(m), C.c_double(theta), C.c_double(phi), out)
    return complex(out[0], out[1])


class Case:
    """The reference's own Geometry + Excitation (as Reader.cpp builds them) for the SH-path functions."""

    def __init__(self):
        lib().refc_create.restype = C.c_void_p
        lib().refc_destroy.argtypes = [C.c_void_p]
        self.h = C.c_void_p(lib().refc_create())
        self.nobj = 0
        self.nMax = self.nMaxS = 0

    def __del__(self):
        try:
            lib().refc_destroy(self.h)
        except Exception:
            pass

    def set_background(self, eps_r, mu_r):
        lib().refc_set_background(self.h, _c2(eps_r), _c2(mu_r))

    def add_sphere(self, xyz_m, radius_m, nMax, model, params, nMaxS=None):
        prm = np.asarray(params, dtype=np.float64)
        nS = nMax if nMaxS is None else nMaxS
        if lib().refc_add_sphere(self.h, (C.c_double * 3)(*xyz_m), C.c_double(radius_m), int(nMax), int(nS), int(model), _p(prm)):
            raise RuntimeError("Geometry::pushObject failed (overlap)")
        self.nobj += 1
        self.nMax, self.nMaxS = nMax, nS

    def set_source(self, wavelength_m, theta, phi, Eth, Eph):
        lib().refc_set_source(self.h, C.c_double(wavelength_m), C.c_double(theta), C.c_double(phi), _c2(Eth), _c2(Eph),
                              int(self.nMax))

    def _n(self):
        return self.nMax * (self.nMax + 2), self.nMaxS * (self.nMaxS + 2)

    def cg_table(self, t):
        n, ns = self._n()
        out = np.zeros(ns * n * n, dtype=np.float64)
        lib().refc_cg_table(self.h, int(t), _p(out))
        return out

    def inc_local_sh(self, obj, internal_FF):
        """Geometry::getIncLocalSH: (v', u', 0, u'') blocks of nS complex each."""
        n, ns = self._n()
        x = np.ascontiguousarray(internal_FF, dtype=np.complex128)
        out = np.zeros(4 * ns, dtype=np.complex128)
        lib().refc_inc_local_sh(self.h, int(obj), _p(x), _p(out))
        return out.reshape(4, ns)

    def cabs_aux(self, obj):
        n, _ = self._n()
        out = np.zeros(2 * n, dtype=np.float64)
        lib().refc_cabs_aux(self.h, int(obj), _p(out))
        return out

    def abs_sh_coeff(self, obj, internal_FF, internal_SH):
        _, ns = self._n()
        a = np.ascontiguousarray(internal_FF, dtype=np.complex128)
        b = np.ascontiguousarray(internal_SH, dtype=np.complex128)
        out = np.zeros(ns, dtype=np.complex128)
        lib().refc_abs_sh_coeff(self.h, int(obj), _p(a), _p(b), _p(out))
        return out

    def coeff_part_sh(self, obj, internal_FF, r):
        _, ns = self._n()
        a = np.ascontiguousarray(internal_FF, dtype=np.complex128)
        xm = np.zeros(ns, dtype=np.complex128)
        xp = np.zeros(ns, dtype=np.complex128)
        lib().refc_coeff_part_sh(self.h, int(obj), _p(a), C.c_double(r), _p(xm), _p(xp))
        return xm, xp

    def check_inner(self, R_sph):
        return int(lib().refc_check_inner(self.h, (C.c_double * 3)(*R_sph)))


def case_from_spec(spec):
    """tests.util.Spec -> reference Case (same construction as tests.util.oracle_case)."""
    c = Case()
    if spec.background is not None:
        c.set_background(*spec.background)
    for p, r, (model, params) in zip(spec.xyz, spec.radius, spec.material):
        c.add_sphere(list(p), float(r), spec.nMax, model, params)
    c.set_source(spec.wavelength, spec.theta, spec.phi, spec.Eth, spec.Eph)
    return c
