"""CPU checks of the oracle's field-map restatement (srcAna/Result.cpp:74-300, AuxCoefficients.cpp:108-343).

The reference holds no fixtures for this path; the restatement is pinned by two closed-form facts:
the regular VSWF expansion with Excitation::populate's coefficients rebuilds the incident plane wave, and the
tangential E and H of a single Mie sphere are continuous across its surface (which ties the vector spherical waves,
the scattered and the internal coefficients together).
"""
import numpy as np

from oracle import oracle as O
from tests import util as U


def _sph(p):
    p = np.asarray(p, dtype=float)
    r = np.linalg.norm(p)
    return [r, np.arccos(p[2] / r), np.arctan2(p[1], p[0])]


def _single_sphere():
    spec = U.Spec("one", [[0, 0, 0]], 50.0, U.SI, 12, 800.0, theta_deg=35, phi_deg=20, Eth=0.7, Eph=0.4j, sh=True)
    c = U.oracle_case(spec)
    c.solve(O.SOLVER_DIRECT)
    return c


def test_incident_expansion_is_the_plane_wave():
    c = _single_sphere()
    k = c.info()["waveK"].real
    th, ph = np.deg2rad(35), np.deg2rad(20)
    khat = np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
    E0 = 0.7 * np.array([np.cos(th) * np.cos(ph), np.cos(th) * np.sin(ph), -np.sin(th)]) + 0.4j * np.array([-np.sin(ph), np.cos(ph), 0])
    c.set_vector(0, c.vector(0) * 0)  # no scattered field
    pts = np.array([[120e-9, 40e-9, -90e-9], [-200e-9, 150e-9, 60e-9]])
    f, inner = c.fields([_sph(p) for p in pts])
    Z = np.sqrt(U.MU0 / U.EPS0)
    assert (inner == -1).all()
    for p, fi in zip(pts, f):
        phase = np.exp(1j * k * khat @ p)
        assert np.abs(fi[0] - E0 * phase).max() < 2e-9          # truncation at nMax = 12, k r <= 2
        assert np.abs(fi[1] - np.cross(khat, E0) / Z * phase).max() < 2e-9 / Z


def test_tangential_fields_are_continuous_across_a_mie_sphere():
    c = _single_sphere()
    for d in [(1, 0, 0), (0.3, 0.5, -0.8), (-0.6, 0.1, 0.7)]:
        d = np.array(d) / np.linalg.norm(d)
        fin, i_in = c.fields([_sph(d * 50e-9 * (1 - 1e-9))])
        fout, i_out = c.fields([_sph(d * 50e-9 * (1 + 1e-9))])
        assert i_in[0] == 0 and i_out[0] == -1                    # Geometry::checkInner
        for t in (0, 1):  # E_FF, H_FF (the SH fields jump: surface source)
            a, b = fin[0][t], fout[0][t]
            ta, tb = a - (a @ d) * d, b - (b @ d) * d
            assert np.abs(ta - tb).max() < 5e-8 * np.abs(tb).max(), (t, d)


def test_vector_spherical_waves_are_divergence_free_and_curl_related():
    # N = curl M / k (finite differences on the oracle's own M): pins compute_Mn against compute_Nn
    k = 2 * np.pi / 800e-9
    p0 = np.array([210e-9, -130e-9, 90e-9])
    h = 1e-11  # central differences: error ~ (k h)^2

    def M(pt, regular):
        return O.aux_coefficients(_sph(pt), k, regular, 4)["M"]

    for regular in (True, False):
        N = O.aux_coefficients(_sph(p0), k, regular, 4)["N"]
        J = np.zeros((3, 24, 3), dtype=complex)  # dM_c / dx_a
        for a in range(3):
            e = np.zeros(3)
            e[a] = h
            J[a] = (M(p0 + e, regular) - M(p0 - e, regular)) / (2 * h)
        curl = np.stack([J[1][:, 2] - J[2][:, 1], J[2][:, 0] - J[0][:, 2], J[0][:, 1] - J[1][:, 0]], 1) / k
        assert np.abs(curl - N).max() < 1e-7 * np.abs(N).max()


def test_grid_enumeration():
    gp = [-100e-9, 100e-9, 5, -80e-9, 80e-9, 3, -50e-9, 250e-9, 4]
    pts = O.grid_points(gp)  # OutputGrid.cpp:132-157: x fastest, +1e-12 on every coordinate
    assert pts.shape == (60, 3)
    x = pts[:, 0] * np.sin(pts[:, 1]) * np.cos(pts[:, 2])
    z = pts[:, 0] * np.cos(pts[:, 1])
    assert np.allclose(x[:5], np.linspace(-100e-9, 100e-9, 5) + 1e-12, atol=1e-20)
    assert np.allclose(z[::15], np.linspace(-50e-9, 250e-9, 4) + 1e-12, atol=1e-20)
