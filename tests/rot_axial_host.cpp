// Host harness around optimet_b200/csrc/ob_rot_axial.cuh: runs the per-pair axial-only recursion the device warp
// runs (same source, lane 0 of 1) on the CPU, for tests/test_rot_axial_host.py.  Test infrastructure.
#include <cmath>
using std::fabs;
using std::fma;
using std::hypot;
using std::sqrt;
#include "../optimet_b200/csrc/ob_rot_axial.cuh"
#include <vector>

extern "C" int rot_axial_host(int NM, const double k[2], double r, double *A, double *B) {
  std::vector<ob::cplx> buf((size_t)ob::rot_axial_buf_entries(NM), ob::mk(0, 0));
  // two pairs through the same buffers: the second run sees the stale level data a warp sees between pairs
  ob::rot_axial_pair(NM, ob::mk(k[0] * 1.7, k[1] + 0.01 * k[0]), 0.6 * r, buf.data(), (ob::cplx *)A, (ob::cplx *)B, 0, 1);
  ob::rot_axial_pair(NM, ob::mk(k[0], k[1]), r, buf.data(), (ob::cplx *)A, (ob::cplx *)B, 0, 1);
  return ob::rot_offX(NM, NM + 1);
}

// record layout v2: Cp = A + B (all mu), Cm = A - B (mu >= 1, stored from offX(1) = NM^2 on)
extern "C" int rot_axial_host_combined(int NM, const double k[2], double r, double *Cp, double *Cm, int mode) {
  std::vector<ob::cplx> buf((size_t)ob::rot_axial_buf_entries(NM), ob::mk(0, 0));
  ob::rot_axial_pair(NM, ob::mk(k[0], k[1]), r, buf.data(), (ob::cplx *)Cp, (ob::cplx *)Cm, 0, 1, mode);
  return ob::rot_offX(NM, NM + 1);
}

// the tabulated path the kernel runs (coefficients and flags from rot_axial_tables_build, two level buffers), planar
// output; two pairs through the same buffers: the second run sees the stale level data a warp sees between pairs
extern "C" int rot_axial_host_tabulated(int NM, const double k[2], double r, double *Cp, double *Cm) {
  std::vector<double> rec, emit;
  std::vector<int> ridx, eidx, eout;
  ob::rot_axial_tables_build(NM, rec, emit, ridx, eidx, eout);
  ob::RotAxTab tab = {rec.data(), emit.data(), ridx.data(), eidx.data(), eout.data(), (int)ridx.size(), (int)eidx.size()};
  std::vector<ob::cplx> buf((size_t)ob::rot_axial_fast_entries(NM), ob::mk(1e300, -1e300)); // poison: must never be read
  ob::rot_axial_pair_fast(NM, ob::mk(k[0] * 0.7, k[1]), 1.3 * r, buf.data(), Cp, Cm, 0, 1, tab);
  ob::rot_axial_pair_fast(NM, ob::mk(k[0], k[1]), r, buf.data(), Cp, Cm, 0, 1, tab);
  return ob::rot_offX(NM, NM + 1);
}

// fragment order of the record's matrices (ob_rot_axial.cuh), for tests/test_rot_axial_host.py
extern "C" int rot_frag_index_host(int rows, int K, int row, int kcol) { return ob::rot_frag_index(rows, K, row, kcol); }
extern "C" int rot_frag_index_a_host(int n, int a, int kcol) { return ob::rot_frag_index_a(n, a, kcol); }
extern "C" int rot_cidx_host(int NM, int mu, int n, int l) { return ob::rot_cidx(NM, mu, n, l); }
