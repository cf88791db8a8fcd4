"""Shared test helpers: case definitions mirrored into the CPU oracle and the GPU context.

The oracle (oracle/) is the checker only; the GPU context gets its inputs either from the product's
own C++ host layer (tests of the host layer) or, for kernel-level parity tests, from the oracle's
host-side scalar preparation (materials / incident coefficients), so that a kernel mismatch cannot
hide behind a host-side one.
"""
import numpy as np

from oracle import oracle as O

EPS0 = 1.0 / (4.0 * np.pi * 1e-7 * 299792458.0 ** 2)
MU0 = 4.0 * np.pi * 1e-7

SI = (O.MODEL_SILICON, [1.0, 0.0])
AU = (O.MODEL_GOLD, [1.0, 0.0, -1.0, 0.0, 1.0, 0.0, 1.0, 0.0])


def fixed(eps, eps_sh, ksippp=1e-19, ksiparppar=2e-19, gamma=0.5e-19, mu=1.0):
    eps, eps_sh, k1, k2, g, mu = map(complex, (eps, eps_sh, ksippp, ksiparppar, gamma, mu))
    return (O.MODEL_FIXED, [eps.real, eps.imag, mu.real, mu.imag, eps_sh.real, eps_sh.imag, k1.real, k1.imag,
                            k2.real, k2.imag, g.real, g.imag])


class Spec:
    def __init__(self, name, xyz_nm, radius_nm, material, nMax, wavelength_nm, theta_deg=45.0, phi_deg=90.0,
                 Eth=1.0, Eph=0.0, sh=True, background=None):
        self.name = name
        self.xyz = np.asarray(xyz_nm, dtype=float) * 1e-9
        self.radius = np.atleast_1d(np.asarray(radius_nm, dtype=float)) * 1e-9
        if self.radius.size == 1:
            self.radius = np.repeat(self.radius, len(self.xyz))
        self.material = material if isinstance(material, list) else [material] * len(self.xyz)
        self.nMax = nMax
        self.wavelength = wavelength_nm * 1e-9
        self.theta = np.deg2rad(theta_deg)
        self.phi = np.deg2rad(phi_deg)
        self.Eth, self.Eph, self.sh = Eth, Eph, sh
        self.background = background


def two_si(nMax=6, wavelength_nm=1240.0):  # examples/TwoParticlesSi.xml
    return Spec("TwoParticlesSi", [[0, 0, 0], [0, 0, 200.0]], 50.0, SI, nMax, wavelength_nm)


def three_au(nMax=3, wavelength_nm=400.0):  # examples/ThreeParticlesAu.xml
    return Spec("ThreeParticlesAu", [[-100, -100, 100], [-150, 150, 100], [200, 200, 100]], [50, 100, 150], AU, nMax,
                wavelength_nm)


def random_cluster(nobj, nMax, seed=1, side_nm=None, radius_nm=50.0, min_dist_nm=150.0, wavelength_nm=800.0,
                   material=SI):
    rng = np.random.RandomState(seed)
    side = side_nm if side_nm is not None else 220.0 * nobj ** (1.0 / 3.0) + 200.0
    pts = []
    while len(pts) < nobj:
        p = rng.uniform(0, side, 3)
        if all(np.linalg.norm(p - q) >= min_dist_nm for q in pts):
            pts.append(p)
    return Spec("random%d" % nobj, pts, radius_nm, material, nMax, wavelength_nm)


def cube_lattice(points, nMax, count=None, d_nm=190.0, radius_nm=50.0, wavelength_nm=800.0, material=SI):
    """First `count` sites of the reference's cube lattice (x fastest; Reader.cpp:150-164)."""
    pts = [[d_nm * i, d_nm * j, d_nm * k] for k in range(points) for j in range(points) for i in range(points)]
    if count is not None:
        pts = pts[:count]
    return Spec("cube%d" % len(pts), pts, radius_nm, material, nMax, wavelength_nm)


def oracle_case(spec):
    c = O.Case()
    for p, r, (model, params) in zip(spec.xyz, spec.radius, spec.material):
        c.add_sphere(list(p), float(r), spec.nMax, model, params)
    if spec.background is not None:
        c.set_background(*spec.background)
    c.set_source(spec.wavelength, spec.theta, spec.phi, spec.Eth, spec.Eph, spec.sh)
    return c


def spherical_roundtrip(xyz):
    """Cartesian -> Tools::toSpherical -> Tools::toCartesian, as the reference stores and re-reads vR."""
    out = np.zeros_like(xyz)
    for i, (x, y, z) in enumerate(xyz):
        r = np.sqrt(x * x + y * y + z * z)
        if r > 0:
            th, ph = np.arccos(z / r), np.arctan2(y, x)
        else:
            th = ph = 0.0
        out[i] = [r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), r * np.cos(th)]
    return out


def configure_ctx(ctx, spec, orc):
    """Feed the GPU context with the same scalars the reference's Geometry/Excitation hold
    (taken from the oracle's restatement of ElectroMagnetic / Excitation::populate)."""
    info = orc.info()
    nobj = info["nobj"]
    ctx.set_cluster(spherical_roundtrip(spec.xyz), spec.radius, spec.nMax)
    mats = [orc.material(j) for j in range(nobj)]
    eps = np.array([m["eps_r"] * EPS0 for m in mats])
    mu = np.array([m["mu_r"] * MU0 for m in mats])
    eps_sh = np.array([m["eps_r_SH"] * EPS0 for m in mats])
    if spec.background is not None:
        eps_b, mu_b = spec.background[0] * EPS0, spec.background[1] * MU0
    else:
        eps_b, mu_b = EPS0, MU0
    ctx.set_frequency(info["omega"], info["waveK"], eps_b, mu_b, eps, mu, eps_sh, mu,
                      [m["ksippp"] for m in mats], [m["ksiparppar"] for m in mats], [m["gamma"] for m in mats])
    a, b = orc.incident()
    ctx.set_incident(a, b)


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    nb = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (nb if nb > 0 else 1.0)


def impose_device_pivots(orc, ctx, harmonics=(1, 2)):
    """Step 2 helper: give the oracle the device's pivots where its own differ; returns (blocks, differing blocks)."""
    total = differ = 0
    for h in harmonics:
        for i in range(ctx.nobj):
            for j in range(ctx.nobj):
                r, _, _, Ig, Jg = ctx.aca_block(h, i, j)
                if r <= 0:
                    continue
                total += 1
                ro, _, _, Io, Jo = orc.aca_block(h, i, j)
                if ro == r and list(Io) == list(Ig) and list(Jo) == list(Jg):
                    orc.force_aca_pivots(h, i, j, None, None)
                else:
                    differ += 1
                    orc.force_aca_pivots(h, i, j, Ig, Jg)
    return total, differ
