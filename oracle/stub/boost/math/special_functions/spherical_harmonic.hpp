/* Stand-in for Boost.Math's spherical_harmonic (absent from the image), written for this repo from its published
 * definition (Boost.Math "Spherical Harmonics": Y_n^m(theta, phi) = sqrt((2n+1)/(4 pi) (n-m)!/(n+m)!) P_n^m(cos theta)
 * e^{i m phi} with the Condon-Shortley phase in P_n^m, and Y_n^{-m} = (-1)^m conj(Y_n^m)).  Checked against
 * scipy.special.sph_harm in tests/test_reference_coupling.py. */
#ifndef OB_STUB_BOOST_SPH_HARM
#define OB_STUB_BOOST_SPH_HARM
#include <cassert> /* the real Boost headers pull it in; the reference relies on that (assert, :80, :103) */
#include <cmath>
#include <complex>
namespace boost { namespace math {
inline std::complex<double> spherical_harmonic(unsigned n, int m, double theta, double phi) {
  const int am = m < 0 ? -m : m;
  if((unsigned)am > n)
    return 0.0;
  const double x = std::cos(theta), s = std::sqrt((1.0 - x) * (1.0 + x));
  // normalised associated Legendre functions by the standard stable upward recurrence in the degree
  double pmm = std::sqrt(1.0 / (4.0 * M_PI));
  for(int i = 1; i <= am; ++i)
    pmm *= -std::sqrt((2.0 * i + 1.0) / (2.0 * i)) * s;
  double p = pmm;
  if((int)n > am) {
    double pm1 = pmm, pc = x * std::sqrt(2.0 * am + 3.0) * pmm;
    for(int l = am + 2; l <= (int)n; ++l) {
      const double a = std::sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)am * am));
      const double b = std::sqrt((((double)l - 1.0) * (l - 1.0) - (double)am * am) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
      const double pn = a * (x * pc - b * pm1);
      pm1 = pc;
      pc = pn;
    }
    p = pc;
  }
  std::complex<double> y = p * std::complex<double>(std::cos(am * phi), std::sin(am * phi));
  if(m < 0) {
    y = std::conj(y);
    if(am & 1)
      y = -y;
  }
  return y;
}
}}
#endif
