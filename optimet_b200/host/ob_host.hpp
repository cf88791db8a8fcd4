// ob_host.hpp -- C++11 host layer above the C ABI (include/optimet_b200.h).
//
// Mirrors the part of OPTIMET-3D's object model the hot path is driven through, with the same names
// and argument meaning, so that the call sites read like the reference's:
//   ElectroMagnetic (srcAna/ElectroMagnetic.{h,cpp})   material models feeding the T-matrix
//   Scatterer / Geometry (srcAna/Scatterer.h, Geometry.{h,cpp})
//   Excitation (srcAna/Excitation.{h,cpp})              plane-wave coefficients, populate()
//   Run + simulation_input (srcAna/Run.h, Reader.cpp)   reads the same XML inputs
//   solver::B200Matrix (drop-in for solver::AbstractSolver, srcAna/Solver.h:62-144)
//   Result (cross sections, srcAna/Result.cpp:557-794)  -> device reductions
//   Simulation::scan_wavelengths (srcAna/Simulation.cpp:604-685)
// No Eigen/Boost/GSL/pugixml: plain std::vector<std::complex<double>> and a small XML reader.
// Everything numerical that is O(N_obj) or larger runs on the GPU through the C ABI; this layer only
// prepares the per-wavelength scalars (materials, incident coefficients) exactly as the reference does.
#pragma once
#include "../../include/optimet_b200.h"
#include <complex>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace optimet_b200 {

typedef double t_real;
typedef std::complex<double> t_complex;
typedef std::vector<t_complex> Vector;

namespace constant {
extern const t_real pi, c, mu0, epsilon0, from_nm_to_m;
}

struct Spherical {
  t_real rrr, the, phi;
  Spherical(t_real r = 0, t_real t = 0, t_real p = 0) : rrr(r), the(t), phi(p) {}
};
struct Cartesian {
  t_real x, y, z;
  Cartesian(t_real x_ = 0, t_real y_ = 0, t_real z_ = 0) : x(x_), y(y_), z(z_) {}
};
Spherical toSpherical(Cartesian const &p); // Tools.cpp:250-258
Cartesian toCartesian(Spherical const &p); // Tools.cpp:38-42

class ElectroMagnetic {
public:
  t_complex epsilon, mu, epsilon_r, mu_r, epsilon_SH, mu_SH, epsilon_r_SH, mu_r_SH, ksippp, ksiparppar, gamma;
  int modelType; // 0 fixed, 3 hydrodynamic (GoldModel), 4 SiliconModel
  ElectroMagnetic();
  void init_r(t_complex epsilon_r_, t_complex mu_r_, t_complex epsilon_r_SH_, t_complex ksippp_, t_complex ksiparppar_,
              t_complex gamma_);
  void initHydrodynamicModel_r(t_complex a_, t_complex b_, t_complex d_, t_complex mu_r_);
  void initSiliconModel_r(t_complex mu_r_);
  void update(t_real lambda_);

private:
  t_complex a_SH, b_SH, d_SH;
  t_real lambda;
  void populateHydrodynamicModel();
  void populateSiliconModel();
};

struct Scatterer {
  Spherical vR;
  ElectroMagnetic elmag;
  t_real radius;
  int nMax, nMaxS;
  std::string scatterer_type;
  Scatterer(int nMax_ = 0, int nMaxS_ = 0) : radius(0), nMax(nMax_), nMaxS(nMaxS_), scatterer_type("sphere") {}
};

class Excitation;

class Geometry {
public:
  std::vector<Scatterer> objects;
  ElectroMagnetic bground;
  bool ACA_cond_;
  Geometry() : ACA_cond_(false) {}
  void pushObject(Scatterer const &object_); // throws on overlap (Geometry.cpp:39-52)
  int nMax() const;
  int nMaxS() const;
  void update(std::shared_ptr<Excitation const> incWave_); // Geometry.cpp:499-503
  void ACAcompression(bool c) { ACA_cond_ = c; }
};

class Excitation {
public:
  t_complex Einc[3]; // Cartesian projection of (0, E_theta, E_phi) (Reader.cpp:819-827)
  Spherical vKInc;
  bool SH_cond;
  int nMax;
  t_complex waveK, bgcoef;
  Vector dataIncAp, dataIncBp;
  Excitation(unsigned long type, const t_complex Einc_[3], bool SH_cond_, Spherical vKInc_, int nMax_, t_complex bgcoeff);
  int populate();                     // Excitation.cpp:48-75
  void updateWavelength(t_real lambda_); // Excitation.cpp:132-137
  t_real lambda() const { return 2 * constant::pi / vKInc.rrr; }
  t_real wavenumber() const { return vKInc.rrr; }
  t_real omega() const { return constant::c * wavenumber(); }
};

// Belos <ParameterList> subset (Reader.cpp:917-928); only the keys the examples use
struct BelosParams {
  std::string solver; // "GMRES" | "scalapack" | "eigen" ("scalapack" when the list is absent, Reader.cpp:925-926)
  double tolerance;
  int max_iterations, num_blocks, block_size, max_restarts, verbosity;
  bool present;
  BelosParams()
      : solver("scalapack"), tolerance(1e-8), max_iterations(1000), num_blocks(300), block_size(1), max_restarts(20),
        verbosity(0), present(false) {}
};

class Run {
public:
  std::shared_ptr<Geometry> geometry;
  std::shared_ptr<Excitation> excitation;
  int nMax, nMaxS;
  double params[9];
  int outputType; // 0 field, 2 coefficients, 11 wavelength scan, 12 radius scan, 112 both
  bool projection; // field output: spherical projection about object 0 (Reader.cpp:860-861)
  BelosParams belos_params;
  Run() : geometry(new Geometry), nMax(0), nMaxS(0), outputType(-1), projection(false) {
    for(int i = 0; i < 9; ++i)
      params[i] = 0;
  }
};

Run simulation_input(std::string const &fileName_);        // Reader.cpp:963-972
Run simulation_input_string(std::string const &xml_text);  // same, from memory

namespace solver {

//! B200 drop-in for optimet::solver::AbstractSolver (Solver.h:62-144): same update()/solve() contract.
class B200Matrix {
public:
  B200Matrix(std::shared_ptr<Geometry> geometry, std::shared_ptr<Excitation const> incWave, int device = 0);
  explicit B200Matrix(Run const &run, int device = 0);
  ~B200Matrix();
  //! multi-GPU: rank/world of this process and the NCCL id broadcast by the host launcher
  void set_communicator(const char uid[128], int rank, int world);
  //! solver options; the default follows the reference's own selection (Solver.cpp:30-54 and
  //! PreconditionedMatrixSolver.h:45-79): Belos list -> Belos GMRES; ACA on -> Gmres_Zcomp constants; else direct solve
  void set_gmres(ob_gmres_opts const &o) { opts = o; }
  ob_gmres_opts const &gmres() const { return opts; }
  //! operator form: -1 (default) follows Geometry::ACA_cond_ (compressed operator when <ACA compression="yes">),
  //! 0 never compress (uncompressed pair / dense form even when the XML asks for ACA), 1 always compress
  void set_aca_mode(int mode) { aca_mode = mode; }

  void update(std::shared_ptr<Geometry> geometry_, std::shared_ptr<Excitation const> incWave_) {
    geometry = geometry_;
    incWave = incWave_;
    update();
  }
  void update(Run const &run) { update(run.geometry, run.excitation); }
  void update(); // pushes geometry + frequency + incident coefficients to the device
  //! CGcoeff: the nine tables in Simulation.cpp:616 order, or empty -> built on the device
  void solve(Vector &X_sca_, Vector &X_int_, Vector &X_sca_SH, Vector &X_int_SH,
             std::vector<double *> CGcoeff = std::vector<double *>()) const;
  size_t scattering_size() const;
  //! Result::setFields on the solution of the last solve(): pts_sph = npts x (r, theta, phi); out = npts x 4 x 3
  //! (E_FF, H_FF, E_SH, H_SH, Cartesian components); inner = Geometry::checkInner per point
  void fields(std::vector<double> const &pts_sph, bool sh, std::vector<t_complex> &out, std::vector<int> &inner) const;
  //! cross sections of the last solve, device reductions: ext, sca, abs(=ext-sca), sca_SH, abs_SH
  void cross_sections(double cs[5]) const { for(int i = 0; i < 5; ++i) cs[i] = last_cs[i]; }
  int iterations(int harmonic) const { return last_iters[harmonic - 1]; }
  ob_ctx *context() const { return ctx; }

protected:
  std::shared_ptr<Geometry> geometry;
  std::shared_ptr<Excitation const> incWave;
  ob_ctx *ctx;
  ob_gmres_opts opts;
  mutable double last_cs[5];
  mutable int last_iters[2];
  mutable bool tables_set;
  int aca_mode;
  bool aca_forced;
  void check(int rc) const;
};

} // namespace solver

//! Result of one wavelength (what Simulation::scan_wavelengths writes, Simulation.cpp:659-667)
struct ScanLine {
  double lambda, absorption_FF, scattering_FF, scattering_SH, absorption_SH, extinction_FF;
  int iters_FF, iters_SH;
};
//! Simulation::scan_wavelengths (Simulation.cpp:604-685); writes the four .dat files when caseFile != ""
std::vector<ScanLine> scan_wavelengths(Run &run, solver::B200Matrix &solver, std::string const &caseFile);
//! Field map of one run (Simulation::field_simulation): grid dims and npts x 3 Cartesian components per field, points
//! in OutputGrid order (x index fastest)
struct FieldMap {
  int nx, ny, nz;
  std::vector<t_complex> E_FF, H_FF, E_SH, H_SH;
  std::vector<int> inner;
};
//! OutputGrid::getPoint enumeration of Run::params (OutputGrid.cpp:132-157): npts x (r, theta, phi)
std::vector<double> grid_points(const double gp[9]);
//! Simulation::field_simulation (Simulation.cpp:319-366); writes <case>_FF.field / <case>_SH.field when caseFile != ""
FieldMap field_simulation(Run &run, solver::B200Matrix &solver, std::string const &caseFile);
void write_field_file(std::string const &path, FieldMap const &fm, std::vector<t_complex> const &E,
                      std::vector<t_complex> const &H);
//! default GMRES options for a Run (see B200Matrix::set_gmres)
ob_gmres_opts default_gmres(Run const &run);

} // namespace optimet_b200
