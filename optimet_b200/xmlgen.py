"""Emit OPTIMET-3D XML inputs (the schema of srcAna/Reader.cpp) for synthetic clusters, so that every
benchmark/test case is an ordinary OPTIMET input the reference could read as well."""
import numpy as np


def _material(m):
    kind = m[0]
    if kind == "silicon":
        return '    <epsilon type="SiliconModel"/>\n'
    if kind == "gold":
        a, b, d = (complex(x) for x in m[1:4])
        return ('    <epsilon type="GoldModel">\n    <parameters a.real="%r" a.imag="%r" b.real="%r" b.imag="%r" '
                'd.real="%r" d.imag="%r" />\n    </epsilon>\n' % (a.real, a.imag, b.real, b.imag, d.real, d.imag))
    if kind == "fixed":
        eps, eps_sh, k1, k2, g = (complex(x) for x in m[1:6])
        s = '    <epsilon type="relative" value.real="%r" value.imag="%r" />\n' % (eps.real, eps.imag)
        for tag, v in (("epsilon_SH", eps_sh), ("ksippp", k1), ("ksiparppar", k2), ("gamma", g)):
            s += '    <%s value.real="%r" value.imag="%r" />\n' % (tag, v.real, v.imag)
        return s
    raise ValueError(kind)


def cluster_xml(xyz_nm, radius_nm, nMax, wavelength_nm, material=("silicon",), theta_deg=45.0, phi_deg=90.0,
                Eth=1.0, Eph=0.0, sh=True, aca=False, scan=None, belos=None, mu=1.0, background=None, field=None,
                projection=False):
    """field = ((xmin, xmax, nx), (ymin, ymax, ny), (zmin, zmax, nz)) in nm selects <output type="field">."""
    xyz = np.asarray(xyz_nm, dtype=float).reshape(-1, 3)
    rad = np.broadcast_to(np.asarray(radius_nm, dtype=float), (len(xyz),))
    mats = material if isinstance(material, list) else [material] * len(xyz)
    Eth, Eph, mu = complex(Eth), complex(Eph), complex(mu)
    out = ['<simulation>\n  <harmonics nmax="%d" />\n  <ACA compression="%s" />\n</simulation>\n'
           % (nMax, "yes" if aca else "no")]
    if belos is not None:
        out.append('<ParameterList name="Belos">\n')
        for name, typ, val in belos:
            out.append('  <Parameter name="%s" type="%s" value="%s"/>\n' % (name, typ, val))
        out.append('</ParameterList>\n')
    out.append('<source type="planewave">\n  <wavelength value="%r" />\n  <propagation theta="%r" phi="%r" />\n'
               '  <polarization Etheta.real="%r" Etheta.imag="%r" Ephi.real="%r" Ephi.imag="%r" />\n'
               '  <SHsources condition="%s" />\n</source>\n'
               % (float(wavelength_nm), float(theta_deg), float(phi_deg), Eth.real, Eth.imag, Eph.real, Eph.imag,
                  "yes" if sh else "no"))
    out.append('<geometry>\n')
    for p, r, m in zip(xyz, rad, mats):
        out.append('  <object type="sphere">\n    <cartesian x="%r" y="%r" z="%r" />\n    <properties radius="%r" />\n'
                   % (float(p[0]), float(p[1]), float(p[2]), float(r)))
        out.append(_material(m))
        out.append('    <mu type="relative" value.real="%r" value.imag="%r" />\n  </object>\n' % (mu.real, mu.imag))
    if background is not None:
        eb, mb = complex(background[0]), complex(background[1])
        out.append('  <background type="absolute">\n    <epsilon value.real="%r" value.imag="%r" />\n'
                   '    <mu value.real="%r" value.imag="%r" />\n  </background>\n' % (eb.real, eb.imag, mb.real, mb.imag))
    out.append('</geometry>\n')
    if field is not None:
        out.append('<output type="field">\n  <grid type="cartesian">\n')
        for ax, (lo, hi, steps) in zip("xyz", field):
            out.append('    <%s min="%r" max="%r" steps="%d" />\n' % (ax, float(lo), float(hi), int(steps)))
        out.append('  </grid>\n  <projection spherical="%s" />\n</output>\n' % ("true" if projection else "false"))
        return "".join(out)
    if scan is None:
        scan = (wavelength_nm, wavelength_nm + 1, 1)
    out.append('<output type="response">\n  <scan type="A+E">\n    <wavelength initial="%r" final="%r" stepsize="%r" />\n'
               '  </scan>\n</output>\n' % (float(scan[0]), float(scan[1]), scan[2]))
    return "".join(out)


def cube_sites(points, count=None, d_nm=190.0):
    """Sites of the reference's cube lattice, x fastest (Reader.cpp:150-164)."""
    pts = [[d_nm * i, d_nm * j, d_nm * k] for k in range(points) for j in range(points) for i in range(points)]
    return np.array(pts[:count] if count is not None else pts)


class MT19937_64:
    """std::mt19937_64 (Matsumoto & Nishimura 2004, the C++11 parameter set); the 10000th output for the default seed
    5489 is 9981545732273789042 ([rand.predef] of the C++ standard; checked in tests/test_host_abi.py)."""
    NN, MM = 312, 156
    MATRIX_A, UM, LM, MASK = 0xB5026F5AA96619E9, 0xFFFFFFFF80000000, 0x7FFFFFFF, (1 << 64) - 1

    def __init__(self, seed=5489):
        self.mt = [0] * self.NN
        self.mt[0] = seed & self.MASK
        for i in range(1, self.NN):
            self.mt[i] = (6364136223846793005 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 62)) + i) & self.MASK
        self.mti = self.NN

    def next(self):
        mt, NN, MM = self.mt, self.NN, self.MM
        if self.mti >= NN:
            for i in range(NN):
                x = (mt[i] & self.UM) | (mt[(i + 1) % NN] & self.LM)
                mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ (self.MATRIX_A if x & 1 else 0)
            self.mti = 0
        x = mt[self.mti]
        self.mti += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        return x & self.MASK

    def uniform(self, lo, hi):
        """std::uniform_real_distribution<double>(lo, hi) as libstdc++ evaluates it for a 64-bit engine: one draw,
        generate_canonical<double, 53> = draw / 2^64 rounded to nearest (values that round to 1 step down)."""
        u = float(self.next()) * (2.0 ** -64)
        if u >= 1.0:
            u = 1.0 - 2.0 ** -53
        return u * (hi - lo) + lo


def random_sites(nobj, side_nm, min_dist_nm, seed):
    """Sequential rejection sampling in a cube (std::mt19937_64 seed 20261017, side 2200 nm,
    minimum centre distance 150 nm); x, y, z of a candidate are drawn in that order with uniform_real_distribution."""
    rng = MT19937_64(seed)
    pts = np.zeros((nobj, 3))
    n = 0
    while n < nobj:
        p = np.array([rng.uniform(0.0, side_nm), rng.uniform(0.0, side_nm), rng.uniform(0.0, side_nm)])
        if n == 0 or np.min(np.linalg.norm(pts[:n] - p, axis=1)) >= min_dist_nm:
            pts[n] = p
            n += 1
    return pts
