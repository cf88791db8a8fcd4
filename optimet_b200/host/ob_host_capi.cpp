// ob_host_capi.cpp -- C entry points of the C++11 host layer (for ctypes-driven tests and bench.py).
#include "ob_host.hpp"
#include <cstring>

using namespace optimet_b200;

namespace {
struct HostCase {
  Run run;
  std::string err;
};
struct HostSolver {
  std::unique_ptr<solver::B200Matrix> s;
  std::string err;
  // the coefficient vectors of the last step, kept across steps and page-locked once they have their size
  Vector x[4];
  void *pinned[4] = {nullptr, nullptr, nullptr, nullptr};
  void pin() {
    for(int i = 0; i < 4; ++i)
      if(!x[i].empty() && pinned[i] != (void *)x[i].data()) {
        if(pinned[i])
          ob_host_unregister(s->context(), pinned[i]);
        pinned[i] = nullptr;
        if(ob_host_register(s->context(), (void *)x[i].data(), x[i].size() * sizeof(t_complex)) == 0)
          pinned[i] = (void *)x[i].data();
      }
  }
  ~HostSolver() {
    for(int i = 0; i < 4; ++i)
      if(pinned[i] && s)
        ob_host_unregister(s->context(), pinned[i]);
  }
};
void set_err(char *err, int errlen, const char *msg) {
  if(err && errlen > 0) {
    std::strncpy(err, msg, errlen - 1);
    err[errlen - 1] = 0;
  }
}
} // namespace

#define OBH_TRY(obj) try {
#define OBH_CATCH(obj)                                                                                                 \
  }                                                                                                                    \
  catch(std::exception & e) {                                                                                          \
    (obj)->err = e.what();                                                                                             \
    return 1;                                                                                                          \
  }                                                                                                                    \
  return 0;

extern "C" {

void *obh_load_xml(const char *path, char *err, int errlen) {
  try {
    HostCase *c = new HostCase;
    c->run = simulation_input(path);
    return c;
  } catch(std::exception &e) {
    set_err(err, errlen, e.what());
    return nullptr;
  }
}
void *obh_load_xml_string(const char *xml, char *err, int errlen) {
  try {
    HostCase *c = new HostCase;
    c->run = simulation_input_string(xml);
    return c;
  } catch(std::exception &e) {
    set_err(err, errlen, e.what());
    return nullptr;
  }
}
void obh_free(void *h) { delete(HostCase *)h; }
const char *obh_error(void *h) { return ((HostCase *)h)->err.c_str(); }

// info: nobj, nMax, nMaxS, SH_cond, outputType, ACA_cond ; params: Run::params[9] ; lambda: current wavelength (m)
int obh_info(void *h, int info[6], double params[9], double *lambda) {
  HostCase *c = (HostCase *)h;
  info[0] = (int)c->run.geometry->objects.size();
  info[1] = c->run.nMax;
  info[2] = c->run.nMaxS;
  info[3] = c->run.excitation->SH_cond ? 1 : 0;
  info[4] = c->run.outputType;
  info[5] = c->run.geometry->ACA_cond_ ? 1 : 0;
  for(int i = 0; i < 9; ++i)
    params[i] = c->run.params[i];
  *lambda = c->run.excitation->lambda();
  return 0;
}
// Simulation.cpp:648-649
int obh_set_wavelength(void *h, double lambda_m) {
  HostCase *c = (HostCase *)h;
  OBH_TRY(c)
  c->run.excitation->updateWavelength(lambda_m);
  c->run.geometry->update(c->run.excitation);
  OBH_CATCH(c)
}
// the scalars update() pushes through the C ABI: xyz[3 nobj] (m), radius[nobj], mats[7][nobj] complex
// (eps, mu, eps_SH, mu_SH, ksippp, ksiparppar, gamma), a/b[n] complex, scal = omega, waveK, eps_b, mu_b (7 doubles)
int obh_get_arrays(void *h, double *xyz, double *radius, double *mats, double *a, double *b, double scal[7]) {
  HostCase *c = (HostCase *)h;
  Geometry const &g = *c->run.geometry;
  const size_t nobj = g.objects.size();
  for(size_t j = 0; j < nobj; ++j) {
    Scatterer const &s = g.objects[j];
    Cartesian p = toCartesian(s.vR);
    xyz[3 * j] = p.x;
    xyz[3 * j + 1] = p.y;
    xyz[3 * j + 2] = p.z;
    radius[j] = s.radius;
    const t_complex m[7] = {s.elmag.epsilon, s.elmag.mu,         s.elmag.epsilon_SH, s.elmag.mu_SH,
                            s.elmag.ksippp,  s.elmag.ksiparppar, s.elmag.gamma};
    for(int i = 0; i < 7; ++i) {
      mats[2 * (i * nobj + j)] = m[i].real();
      mats[2 * (i * nobj + j) + 1] = m[i].imag();
    }
  }
  Excitation const &e = *c->run.excitation;
  std::memcpy(a, e.dataIncAp.data(), e.dataIncAp.size() * sizeof(t_complex));
  std::memcpy(b, e.dataIncBp.data(), e.dataIncBp.size() * sizeof(t_complex));
  scal[0] = e.omega();
  scal[1] = e.waveK.real();
  scal[2] = e.waveK.imag();
  scal[3] = g.bground.epsilon.real();
  scal[4] = g.bground.epsilon.imag();
  scal[5] = g.bground.mu.real();
  scal[6] = g.bground.mu.imag();
  return 0;
}
int obh_gmres_defaults(void *h, ob_gmres_opts *o) {
  *o = default_gmres(((HostCase *)h)->run);
  return 0;
}

void *obh_solver_create(void *h, int device, char *err, int errlen) {
  try {
    HostSolver *s = new HostSolver;
    s->s.reset(new solver::B200Matrix(((HostCase *)h)->run, device));
    return s;
  } catch(std::exception &e) {
    set_err(err, errlen, e.what());
    return nullptr;
  }
}
void obh_solver_free(void *s) { delete(HostSolver *)s; }
const char *obh_solver_error(void *s) { return ((HostSolver *)s)->err.c_str(); }
ob_ctx *obh_solver_ctx(void *s) { return ((HostSolver *)s)->s->context(); }
int obh_solver_comm(void *s_, const char uid[128], int rank, int world) {
  HostSolver *s = (HostSolver *)s_;
  OBH_TRY(s)
  s->s->set_communicator(uid, rank, world);
  OBH_CATCH(s)
}
int obh_solver_set_gmres(void *s_, const ob_gmres_opts *o) {
  ((HostSolver *)s_)->s->set_gmres(*o);
  return 0;
}
int obh_solver_set_aca_mode(void *s_, int mode) {
  ((HostSolver *)s_)->s->set_aca_mode(mode);
  return 0;
}
// one wavelength: solver->update(run) + solver->solve(...) + cross sections (Simulation.cpp:651-667).
// lambda_m > 0 first moves the run to that wavelength.  Output vectors may be NULL.
int obh_solver_step(void *s_, void *h, double lambda_m, double *X_sca, double *X_int, double *X_sca_SH,
                    double *X_int_SH, double cs[5], int iters[2]) {
  HostSolver *s = (HostSolver *)s_;
  HostCase *c = (HostCase *)h;
  OBH_TRY(s)
  if(lambda_m > 0) {
    c->run.excitation->updateWavelength(lambda_m);
    c->run.geometry->update(c->run.excitation);
  }
  s->s->update(c->run);
  {
    // size the persistent vectors before the solve so that they can be page-locked (solve() keeps vectors of the right
    // size as they are)
    const size_t nobj = c->run.geometry->objects.size();
    const size_t n1 = c->run.geometry->nMax(), n2 = c->run.geometry->nMaxS();
    const size_t N1 = 2 * n1 * (n1 + 2) * nobj, N2 = 2 * n2 * (n2 + 2) * nobj;
    const size_t want[4] = {N1, N1, c->run.excitation->SH_cond ? N2 : 0, c->run.excitation->SH_cond ? N2 : 0};
    for(int i = 0; i < 4; ++i)
      if(s->x[i].size() != want[i])
        s->x[i].assign(want[i], t_complex(0, 0));
    s->pin();
  }
  s->s->solve(s->x[0], s->x[1], s->x[2], s->x[3]);
  auto put = [](Vector const &v, double *dst) { // optional copies into caller buffers (obh_solver_vectors gives views)
    if(dst && !v.empty())
      std::memcpy(dst, v.data(), v.size() * sizeof(t_complex));
  };
  put(s->x[0], X_sca);
  put(s->x[1], X_int);
  put(s->x[2], X_sca_SH);
  put(s->x[3], X_int_SH);
  s->s->cross_sections(cs);
  iters[0] = s->s->iterations(1);
  iters[1] = s->s->iterations(2);
  OBH_CATCH(s)
}
// the solver's own (page-locked) coefficient vectors of the last step: X_sca, X_int, X_sca_SH, X_int_SH
int obh_solver_vectors(void *s_, const double *ptrs[4], long sizes[4]) {
  HostSolver *s = (HostSolver *)s_;
  for(int i = 0; i < 4; ++i) {
    ptrs[i] = s->x[i].empty() ? nullptr : (const double *)s->x[i].data();
    sizes[i] = (long)s->x[i].size();
  }
  return 0;
}
// Simulation::scan_wavelengths; lines: lambda, abs_FF, sca_FF, sca_SH, abs_SH, ext_FF, iters_FF, iters_SH per wavelength
int obh_scan(void *s_, void *h, const char *caseFile, double *lines, int maxlines, int *nlines) {
  HostSolver *s = (HostSolver *)s_;
  HostCase *c = (HostCase *)h;
  OBH_TRY(s)
  std::vector<ScanLine> r = scan_wavelengths(c->run, *s->s, caseFile ? caseFile : "");
  *nlines = (int)r.size();
  for(int i = 0; i < (int)r.size() && i < maxlines; ++i) {
    double *l = lines + 8 * i;
    l[0] = r[i].lambda;
    l[1] = r[i].absorption_FF;
    l[2] = r[i].scattering_FF;
    l[3] = r[i].scattering_SH;
    l[4] = r[i].absorption_SH;
    l[5] = r[i].extinction_FF;
    l[6] = r[i].iters_FF;
    l[7] = r[i].iters_SH;
  }
  OBH_CATCH(s)
}
// Simulation::field_simulation: out = npts x 4 x 3 complex (E_FF, H_FF, E_SH, H_SH), inner = npts, dims = nx, ny, nz
int obh_field_simulation(void *s_, void *h, const char *caseFile, double *out, int *inner, long maxpts, int dims[3]) {
  HostSolver *s = (HostSolver *)s_;
  HostCase *c = (HostCase *)h;
  OBH_TRY(s)
  FieldMap fm = field_simulation(c->run, *s->s, caseFile ? caseFile : "");
  dims[0] = fm.nx;
  dims[1] = fm.ny;
  dims[2] = fm.nz;
  const long npts = (long)fm.nx * fm.ny * fm.nz;
  if(npts > maxpts)
    throw std::runtime_error("obh_field_simulation: output buffer too small");
  t_complex *o = (t_complex *)out;
  std::vector<t_complex> const *src[4] = {&fm.E_FF, &fm.H_FF, &fm.E_SH, &fm.H_SH};
  for(long i = 0; i < npts; ++i) {
    for(int f = 0; f < 4; ++f)
      for(int k = 0; k < 3; ++k)
        o[(i * 4 + f) * 3 + k] = (*src[f])[3 * i + k];
    inner[i] = fm.inner[i];
  }
  OBH_CATCH(s)
}
int obh_grid_points(void *h, double *pts, long maxpts, long *npts) {
  HostCase *c = (HostCase *)h;
  std::vector<double> p = grid_points(c->run.params);
  *npts = (long)(p.size() / 3);
  if(*npts > maxpts)
    return 1;
  std::memcpy(pts, p.data(), p.size() * sizeof(double));
  return 0;
}
} // extern "C"
