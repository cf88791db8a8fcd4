// C entry points over the reference's own classes, compiled from the reference sources where they lie (oracle/Makefile,
// target ref_path): Coupling (Coupling.h:29-41), ElectroMagnetic (material models), Scatterer (Mie and auxiliary factors)
// and AuxCoefficients (vector spherical waves).  Used by tests/test_reference_build.py to pin the oracle -- and, on the
// GPU box where oracle/_ref travels as a built library, the CUDA path -- against the reference's compiled code.
// Only the calling conventions live here; no arithmetic.
#include "AuxCoefficients.h"
#include "Coupling.h"
#include "ElectroMagnetic.h"
#include "Excitation.h"
#include "Geometry.h"
#include "Scatterer.h"
#include "Tools.h"
#include <boost/math/special_functions/spherical_harmonic.hpp>
#include <cstring>
#include <memory>
using namespace optimet;
typedef std::complex<double> cd;

static ElectroMagnetic make_elmag(int model, const double *p, double lambda) {
  ElectroMagnetic e;
  if(model == 0)
    e.init_r(cd(p[0], p[1]), cd(p[2], p[3]), cd(p[4], p[5]), cd(p[6], p[7]), cd(p[8], p[9]), cd(p[10], p[11]));
  else if(model == 3) { // Reader.cpp:648-651
    e.init_r(0.0, cd(p[6], p[7]), 0.0, 0.0, 0.0, 0.0);
    e.initHydrodynamicModel_r(cd(p[0], p[1]), cd(p[2], p[3]), cd(p[4], p[5]), cd(p[6], p[7]));
  } else { // 4, Reader.cpp:656-659
    e.init_r(0.0, cd(p[0], p[1]), 0.0, 0.0, 0.0, 0.0);
    e.initSiliconModel_r(cd(p[0], p[1]));
  }
  e.update(lambda); // Geometry::update (Geometry.cpp:499-503)
  return e;
}

extern "C" {
int ref_coupling(const double R[3], const double k[2], int nMax, int regular_flag, double *A, double *B) {
  Coupling c(Spherical<t_real>(R[0], R[1], R[2]), t_complex(k[0], k[1]), (t_uint)nMax, regular_flag != 0);
  std::memcpy(A, c.diagonal.data(), sizeof(t_complex) * (size_t)c.diagonal.size());
  std::memcpy(B, c.offdiagonal.data(), sizeof(t_complex) * (size_t)c.offdiagonal.size());
  return 0;
}
// out: eps_r, eps_r_SH, ksippp, ksiparppar, gamma, mu_r (the order of oracle.Case.material)
int ref_material(int model, const double *p, double lambda, double *out) {
  ElectroMagnetic e = make_elmag(model, p, lambda);
  cd v[6] = {e.epsilon_r, e.epsilon_r_SH, e.ksippp, e.ksiparppar, e.gamma, e.mu_r};
  std::memcpy(out, v, sizeof(v));
  return 0;
}
// which: 0 T_FF, 1 T_SH (diagonals), 2 TSH1_outer, 3 TSH2_outer, 4 Iaux, 5 IauxSH1, 6 IauxSH2 -> 2n complex
int ref_particle_factors(int model, const double *p, double radius, int nMax, int nMaxS, double lambda,
                         const double bg_eps_r[2], const double bg_mu_r[2], int which, double *out) {
  ElectroMagnetic bg;
  bg.init_r(cd(bg_eps_r[0], bg_eps_r[1]), cd(bg_mu_r[0], bg_mu_r[1]), 0.0, 0.0, 0.0, 0.0); // Reader.cpp:84-96
  Scatterer s(Spherical<double>(0.0, 0.0, 0.0), make_elmag(model, p, lambda), radius, nMax, nMaxS);
  s.scatterer_type = "sphere"; // Reader.cpp:588
  const double omega = constant::c * 2.0 * constant::pi / lambda; // Excitation::omega() = c k_0 (Excitation.h)
  Vector<t_complex> v;
  if(which == 0 || which == 1) {
    Matrix<t_complex> T;
    if(which == 0)
      s.getTLocal(T, omega, bg);
    else
      s.getTLocalSH(T, omega, bg);
    for(std::ptrdiff_t i = 0; i < T.rows(); ++i)
      ((t_complex *)out)[i] = T(i, i);
    return 0;
  }
  switch(which) {
  case 2: v = s.getTLocalSH1_outer(omega, bg); break;
  case 3: v = s.getTLocalSH2_outer(omega, bg); break;
  case 4: v = s.getIaux(omega, bg); break;
  case 5: v = s.getIauxSH1(omega, bg); break;
  default: v = s.getIauxSH2(omega, bg); break;
  }
  std::memcpy(out, v.data(), sizeof(t_complex) * (size_t)v.size());
  return 0;
}
// AuxCoefficients(R, waveK, regular, nMax): out = 4 x n x 3 complex (M, N, Xm, Xp; Cartesian components)
int ref_aux_coefficients(const double R[3], const double k[2], int regular, int nMax, double *out) {
  AuxCoefficients a(Spherical<t_real>(R[0], R[1], R[2]), t_complex(k[0], k[1]), regular != 0, (t_uint)nMax);
  const int N = nMax * (nMax + 2);
  t_complex *o = (t_complex *)out;
  for(int p = 0; p < N; ++p) {
    SphericalP<t_complex> const v[4] = {a.M(p), a.N(p), a.Xm(p), a.Xp(p)};
    for(int t = 0; t < 4; ++t) {
      o[((size_t)t * N + p) * 3 + 0] = v[t].rrr;
      o[((size_t)t * N + p) * 3 + 1] = v[t].the;
      o[((size_t)t * N + p) * 3 + 2] = v[t].phi;
    }
  }
  return 0;
}
// Excitation as Reader.cpp:800-832 builds it: plane wave (theta, phi), E = Etheta e_theta + Ephi e_phi
static std::shared_ptr<Excitation> make_excitation(double lambda, double theta, double phi, const double Eth[2],
                                                   const double Eph[2], const double bg_eps_r[2], const double bg_mu_r[2],
                                                   int nMax) {
  const cd bgcoeff = std::sqrt(cd(bg_eps_r[0], bg_eps_r[1]) * cd(bg_mu_r[0], bg_mu_r[1]));
  Spherical<double> vKinc(2 * consPi / lambda, theta, phi);
  SphericalP<cd> Eaux(cd(0.0, 0.0), cd(Eth[0], Eth[1]), cd(Eph[0], Eph[1]));
  SphericalP<cd> Einc = Tools::toProjection(Spherical<double>(0.0, vKinc.the, vKinc.phi), Eaux);
  auto e = std::make_shared<Excitation>(0, Einc, true, vKinc, nMax, bgcoeff);
  e->populate();
  return e;
}
// a = dataIncAp, b = dataIncBp (n complex each); waveK returned as (re, im)
int ref_excitation(double lambda, double theta, double phi, const double Eth[2], const double Eph[2],
                   const double bg_eps_r[2], const double bg_mu_r[2], int nMax, double *a, double *b, double waveK[2]) {
  auto e = make_excitation(lambda, theta, phi, Eth, Eph, bg_eps_r, bg_mu_r, nMax);
  const int N = nMax * (nMax + 2);
  std::memcpy(a, e->dataIncAp.data(), sizeof(cd) * (size_t)N);
  std::memcpy(b, e->dataIncBp.data(), sizeof(cd) * (size_t)N);
  waveK[0] = e->waveK.real();
  waveK[1] = e->waveK.imag();
  return 0;
}
// Excitation::getIncLocal at the spherical point R (Excitation.cpp:79-129): 2n complex
int ref_inc_local(double lambda, double theta, double phi, const double Eth[2], const double Eph[2],
                  const double bg_eps_r[2], const double bg_mu_r[2], int nMax, const double R[3], double *out) {
  auto e = make_excitation(lambda, theta, phi, Eth, Eph, bg_eps_r, bg_mu_r, nMax);
  return e->getIncLocal(Spherical<double>(R[0], R[1], R[2]), (cd *)out, nMax);
}
// Spherical<double>::operator- (Spherical.h:129-131) and Tools::toSpherical (Tools.cpp:250-258): relative position of
// two scatterers exactly as PreconditionedMatrix.cpp:384 forms it
int ref_relative_position(const double xyz_i[3], const double xyz_j[3], double out[3]) {
  const Spherical<double> a = Tools::toSpherical(Cartesian<double>(xyz_i[0], xyz_i[1], xyz_i[2]));
  const Spherical<double> b = Tools::toSpherical(Cartesian<double>(xyz_j[0], xyz_j[1], xyz_j[2]));
  const Spherical<double> d = a - b;
  out[0] = d.rrr;
  out[1] = d.the;
  out[2] = d.phi;
  return 0;
}
// the Boost.Math stand-in itself, for its check against scipy
int ref_ynm(int n, int m, double theta, double phi, double out[2]) {
  const cd y = boost::math::spherical_harmonic((unsigned)n, m, theta, phi);
  out[0] = y.real();
  out[1] = y.imag();
  return 0;
}
// ---------------------------------------------------------------------------------------------------------------
// a case over the reference's own Geometry / Excitation (what Reader.cpp builds), for the SH path:
// Geometry::Coefficients (symbol::C_*coeff / W_*coeff), getIncLocalSH (symbol::vp_mn, up_mn, upp_mn), getCabsAux,
// AbsCSSHcoeff (symbol::ACSshcoeff), COEFFpartSH (symbol::CXm1 / CXp1), checkInner
// ---------------------------------------------------------------------------------------------------------------
struct RefCase {
  std::shared_ptr<Geometry> g;
  std::shared_ptr<Excitation> e;
  std::vector<std::vector<double>> tab;
  std::vector<double *> ptrs;
  RefCase() : g(std::make_shared<Geometry>()) {}
};
void *refc_create() { return new RefCase(); }
void refc_destroy(void *h) { delete(RefCase *)h; }
int refc_set_background(void *h, const double eps_r[2], const double mu_r[2]) { // Reader.cpp:84-96
  ((RefCase *)h)->g->bground.init_r(cd(eps_r[0], eps_r[1]), cd(mu_r[0], mu_r[1]), 0.0, 0.0, 0.0, 0.0);
  return 0;
}
int refc_add_sphere(void *h, const double xyz[3], double radius, int nMax, int nMaxS, int model, const double *p) {
  try { // Reader.cpp:575-590: Cartesian -> spherical position, sphere type; the model is updated in refc_set_source
    ElectroMagnetic e;
    if(model == 0)
      e.init_r(cd(p[0], p[1]), cd(p[2], p[3]), cd(p[4], p[5]), cd(p[6], p[7]), cd(p[8], p[9]), cd(p[10], p[11]));
    else if(model == 3) {
      e.init_r(0.0, cd(p[6], p[7]), 0.0, 0.0, 0.0, 0.0);
      e.initHydrodynamicModel_r(cd(p[0], p[1]), cd(p[2], p[3]), cd(p[4], p[5]), cd(p[6], p[7]));
    } else {
      e.init_r(0.0, cd(p[0], p[1]), 0.0, 0.0, 0.0, 0.0);
      e.initSiliconModel_r(cd(p[0], p[1]));
    }
    Scatterer s(Tools::toSpherical(Cartesian<double>(xyz[0], xyz[1], xyz[2])), e, radius, nMax, nMaxS);
    s.scatterer_type = "sphere";
    ((RefCase *)h)->g->pushObject(s);
  } catch(std::exception &) {
    return 1;
  }
  return 0;
}
int refc_set_source(void *h, double lambda, double theta, double phi, const double Eth[2], const double Eph[2], int nMax) {
  RefCase *c = (RefCase *)h;
  const double er[2] = {c->g->bground.epsilon_r.real(), c->g->bground.epsilon_r.imag()};
  const double mr[2] = {c->g->bground.mu_r.real(), c->g->bground.mu_r.imag()};
  c->e = make_excitation(lambda, theta, phi, Eth, Eph, er, mr, nMax);
  c->g->update(c->e); // Geometry.cpp:499-503
  return 0;
}
static std::vector<double *> &tables(RefCase *c) { // Simulation.cpp:613-618 + Geometry::Coefficients
  if(c->tab.empty()) {
    const size_t n = c->g->nMax() * (c->g->nMax() + 2), ns = c->g->nMaxS() * (c->g->nMaxS() + 2);
    c->tab.assign(9, std::vector<double>(ns * n * n));
    for(auto &t : c->tab)
      c->ptrs.push_back(t.data());
    c->g->Coefficients((int)c->g->nMax(), (int)c->g->nMaxS(), c->ptrs);
  }
  return c->ptrs;
}
int refc_cg_table(void *h, int t, double *out) {
  RefCase *c = (RefCase *)h;
  tables(c);
  std::memcpy(out, c->tab[t].data(), c->tab[t].size() * sizeof(double));
  return 0;
}
static Vector<t_complex> to_vec(const double *v, size_t n) {
  Vector<t_complex> r((std::ptrdiff_t)n);
  std::memcpy(r.data(), v, n * sizeof(cd));
  return r;
}
// Geometry::getIncLocalSH: out = 4 nS complex (v', u', 0, u'')
int refc_inc_local_sh(void *h, int obj, const double *internal_FF, double *out) {
  RefCase *c = (RefCase *)h;
  const size_t N = 2 * c->g->nMax() * (c->g->nMax() + 2) * c->g->objects.size();
  Vector<t_complex> xi = to_vec(internal_FF, N);
  return c->g->getIncLocalSH(tables(c), obj, c->e, xi, (int)c->g->nMaxS(), (cd *)out);
}
int refc_cabs_aux(void *h, int obj, double *out) {
  RefCase *c = (RefCase *)h;
  return c->g->getCabsAux(c->e->omega(), obj, (int)c->g->nMax(), out);
}
int refc_abs_sh_coeff(void *h, int obj, const double *internal_FF, const double *internal_SH, double *out) {
  RefCase *c = (RefCase *)h;
  const size_t nobj = c->g->objects.size();
  Vector<t_complex> xi = to_vec(internal_FF, 2 * c->g->nMax() * (c->g->nMax() + 2) * nobj);
  Vector<t_complex> xs = to_vec(internal_SH, 2 * c->g->nMaxS() * (c->g->nMaxS() + 2) * nobj);
  return c->g->AbsCSSHcoeff(tables(c), obj, c->e, xi, xs, (int)c->g->nMaxS(), (cd *)out);
}
int refc_coeff_part_sh(void *h, int obj, const double *internal_FF, double r, double *xmn, double *xpl) {
  RefCase *c = (RefCase *)h;
  Vector<t_complex> xi = to_vec(internal_FF, 2 * c->g->nMax() * (c->g->nMax() + 2) * c->g->objects.size());
  return c->g->COEFFpartSH(obj, c->e, xi, r, (int)c->g->nMaxS(), (cd *)xmn, (cd *)xpl, tables(c));
}
int refc_check_inner(void *h, const double R[3]) { return ((RefCase *)h)->g->checkInner(Spherical<double>(R[0], R[1], R[2])); }
}
