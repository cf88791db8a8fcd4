// Wigner 3j / 6j / 9j for integer angular momenta (all the reference passes: Symbol.cpp:28-42 doubles its int
// arguments), Racah's single-sum formulas in long double with exact factorials; written for this repo.
#include "gsl/gsl_sf_coupling.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
namespace {
long double lfact(int n) {
  static long double t[200];
  static bool init = false;
  if(!init) {
    t[0] = 1;
    for(int i = 1; i < 200; ++i)
      t[i] = t[i - 1] * (long double)i;
    init = true;
  }
  return t[n];
}
bool tri_bad(int a, int b, int c) { return c < std::abs(a - b) || c > a + b; }
long double tri(int a, int b, int c) { return lfact(a + b - c) * lfact(a - b + c) * lfact(-a + b + c) / lfact(a + b + c + 1); }
double w3j(int j1, int j2, int j3, int m1, int m2, int m3) {
  if(j1 < 0 || j2 < 0 || j3 < 0 || tri_bad(j1, j2, j3) || m1 + m2 + m3 != 0 || std::abs(m1) > j1 || std::abs(m2) > j2 ||
     std::abs(m3) > j3)
    return 0;
  const int kmin = std::max(0, std::max(j2 - j3 - m1, j1 - j3 + m2));
  const int kmax = std::min(j1 + j2 - j3, std::min(j1 - m1, j2 + m2));
  long double sum = 0;
  for(int k = kmin; k <= kmax; ++k) {
    const long double t = 1.0L / (lfact(k) * lfact(j1 + j2 - j3 - k) * lfact(j1 - m1 - k) * lfact(j2 + m2 - k) *
                                  lfact(j3 - j2 + m1 + k) * lfact(j3 - j1 - m2 + k));
    sum += (k & 1) ? -t : t;
  }
  long double r = std::sqrt(tri(j1, j2, j3) * lfact(j1 + m1) * lfact(j1 - m1) * lfact(j2 + m2) * lfact(j2 - m2) *
                            lfact(j3 + m3) * lfact(j3 - m3)) *
                  sum;
  return (double)((std::abs(j1 - j2 - m3) & 1) ? -r : r);
}
double w6j(int j1, int j2, int j3, int j4, int j5, int j6) {
  if(j1 < 0 || j2 < 0 || j3 < 0 || j4 < 0 || j5 < 0 || j6 < 0 || tri_bad(j1, j2, j3) || tri_bad(j1, j5, j6) ||
     tri_bad(j4, j2, j6) || tri_bad(j4, j5, j3))
    return 0;
  const int a1 = j1 + j2 + j3, a2 = j1 + j5 + j6, a3 = j4 + j2 + j6, a4 = j4 + j5 + j3;
  const int b1 = j1 + j2 + j4 + j5, b2 = j2 + j3 + j5 + j6, b3 = j3 + j1 + j6 + j4;
  long double sum = 0;
  for(int k = std::max(std::max(a1, a2), std::max(a3, a4)); k <= std::min(b1, std::min(b2, b3)); ++k) {
    const long double t = lfact(k + 1) / (lfact(k - a1) * lfact(k - a2) * lfact(k - a3) * lfact(k - a4) * lfact(b1 - k) *
                                          lfact(b2 - k) * lfact(b3 - k));
    sum += (k & 1) ? -t : t;
  }
  return (double)(std::sqrt(tri(j1, j2, j3) * tri(j1, j5, j6) * tri(j4, j2, j6) * tri(j4, j5, j3)) * sum);
}
double w9j(int j11, int j12, int j13, int j21, int j22, int j23, int j31, int j32, int j33) {
  if(j11 < 0 || j12 < 0 || j13 < 0 || j21 < 0 || j22 < 0 || j23 < 0 || j31 < 0 || j32 < 0 || j33 < 0 ||
     tri_bad(j11, j12, j13) || tri_bad(j21, j22, j23) || tri_bad(j31, j32, j33) || tri_bad(j11, j21, j31) ||
     tri_bad(j12, j22, j32) || tri_bad(j13, j23, j33))
    return 0;
  long double sum = 0;
  for(int k = std::max(std::abs(j11 - j33), std::max(std::abs(j32 - j21), std::abs(j23 - j12)));
      k <= std::min(j11 + j33, std::min(j32 + j21, j23 + j12)); ++k)
    sum += (long double)(2 * k + 1) * (long double)w6j(j11, j21, j31, j32, j33, k) * (long double)w6j(j12, j22, j32, j21, k, j23) *
           (long double)w6j(j13, j23, j33, k, j11, j12);
  return (double)sum;
}
}
extern "C" {
double gsl_sf_coupling_3j(int a, int b, int c, int d, int e, int f) { return w3j(a / 2, b / 2, c / 2, d / 2, e / 2, f / 2); }
double gsl_sf_coupling_6j(int a, int b, int c, int d, int e, int f) { return w6j(a / 2, b / 2, c / 2, d / 2, e / 2, f / 2); }
double gsl_sf_coupling_9j(int a, int b, int c, int d, int e, int f, int g, int h, int i) {
  return w9j(a / 2, b / 2, c / 2, d / 2, e / 2, f / 2, g / 2, h / 2, i / 2);
}
}
