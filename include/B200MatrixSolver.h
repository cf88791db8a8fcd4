// B200MatrixSolver.h -- the file a maintainer adds to the reference tree (srcAna/) to route solver::factory
// (srcAna/Solver.cpp:30-54) to the B200 path.  It derives from the reference's own AbstractSolver
// (srcAna/Solver.h:62-144), so Simulation::scan_wavelengths (srcAna/Simulation.cpp:643-667) and Result stay untouched.
// tests/test_adaptor_compiles.py compiles it against the reference's headers.
//
// The serial build's factory returns ONE solver object (Solver.cpp:32-34): `devices` lets that object use several GPUs
// of the box through the single-process group of include/optimet_b200.h (ob_create_multi).
#ifndef OPTIMET_B200_MATRIX_SOLVER_H
#define OPTIMET_B200_MATRIX_SOLVER_H

#include "Solver.h"
#include "Tools.h"
#include <optimet_b200.h> // the C ABI
#include <stdexcept>
#include <vector>

namespace optimet {
namespace solver {
class B200Matrix : public AbstractSolver {
public:
  //! `devices`: CUDA ordinals to use (default: device 0 only)
  B200Matrix(Run const &run, std::vector<int> const &devices = std::vector<int>(1, 0))
      : AbstractSolver(run), group(nullptr) {
    if(ob_create_multi((int)devices.size(), devices.data(), &group))
      throw std::runtime_error(ob_last_error(nullptr));
    update();
  }
  ~B200Matrix() { ob_destroy_multi(group); }

  void update() override { // PreconditionedMatrixSolver.h:82-100
    auto const &objs = geometry->objects;
    int const N = objs.size();
    std::vector<double> xyz(3 * N), rad(N);
    std::vector<t_complex> eps(N), mu(N), epsS(N), muS(N), k1(N), k2(N), g(N);
    for(int j = 0; j < N; ++j) {
      auto const c = Tools::toCartesian(objs[j].vR); // Tools.cpp:38-42
      xyz[3 * j] = c.x;
      xyz[3 * j + 1] = c.y;
      xyz[3 * j + 2] = c.z;
      rad[j] = objs[j].radius;
      auto const &e = objs[j].elmag; // ElectroMagnetic.h
      eps[j] = e.epsilon;
      mu[j] = e.mu;
      epsS[j] = e.epsilon_SH;
      muS[j] = e.mu_SH;
      k1[j] = e.ksippp;
      k2[j] = e.ksiparppar;
      g[j] = e.gamma;
    }
    // Scattering_matrix_ACA_FF/_SH when <ACA compression="yes"> (PreconditionedMatrix.cpp:489-551), else the exact
    // rotated-axial form (3): same operator as the dense matrix, 21 KB per particle pair at nMax 10
    check(ob_multi_set_option(group, "operator", geometry->ACA_cond_ ? 2 : 3));
    check(ob_multi_set_cluster(group, N, xyz.data(), rad.data(), geometry->nMax(), geometry->nMaxS()));
    t_complex const k = incWave->waveK, eb = geometry->bground.epsilon, mb = geometry->bground.mu;
    check(ob_multi_set_frequency(group, incWave->omega(), (double const *)&k, (double const *)&eb, (double const *)&mb,
                                 (double const *)eps.data(), (double const *)mu.data(), (double const *)epsS.data(),
                                 (double const *)muS.data(), (double const *)k1.data(), (double const *)k2.data(),
                                 (double const *)g.data()));
    check(ob_multi_set_incident(group, (double const *)incWave->dataIncAp.data(), (double const *)incWave->dataIncBp.data()));
  }

  void solve(Vector<t_complex> &X_sca_, Vector<t_complex> &X_int_, Vector<t_complex> &X_sca_SH,
             Vector<t_complex> &X_int_SH, std::vector<double *> CGcoeff) const override {
    auto const N1 = 2 * Tools::iteratorMax(geometry->nMax()) * geometry->objects.size();
    auto const N2 = 2 * Tools::iteratorMax(geometry->nMaxS()) * geometry->objects.size();
    X_sca_.resize(N1);
    X_int_.resize(N1);
    X_sca_SH.resize(N2);
    X_int_SH.resize(N2);
    // the device builds the nine CG / W tables itself (same values as Simulation.cpp:613-618 passes in CGcoeff:
    // tests/test_gpu_parity.py::test_cg_tables); a host that insists on its own can push them with ob_set_cg_tables
    (void)CGcoeff;
    // PreconditionedMatrixSolver.h:50-58: ACA on -> Gmres_Zcomp over the compressed operator, else the direct solve
    ob_gmres_opts o = {geometry->ACA_cond_ ? OB_GMRES_ZCOMP : OB_SOLVE_DIRECT, 1e-6, 240, 0, 2};
    double cs[5];
    int it[2];
    check(ob_multi_run(group, &o, incWave->SH_cond, (double *)X_sca_.data(), (double *)X_int_.data(),
                       (double *)X_sca_SH.data(), (double *)X_int_SH.data(), cs, it));
  }

private:
  ob_multi *group;
  void check(int rc) const { // the reference reports errors as exceptions
    if(rc)
      throw std::runtime_error(ob_multi_last_error(group));
  }
};
} // namespace solver
} // namespace optimet
#endif
