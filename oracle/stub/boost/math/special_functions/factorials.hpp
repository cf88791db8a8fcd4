/* Stand-in for Boost.Math factorial<T>(unsigned) (absent from the image): the exact product, as Boost's table gives it
 * for the arguments the reference uses (AuxCoefficients.cpp:236-241: (2m)! and m!, m <= nMax). */
#ifndef OB_STUB_BOOST_FACTORIALS
#define OB_STUB_BOOST_FACTORIALS
namespace boost { namespace math {
template <class T> inline T factorial(unsigned n) {
  T r = T(1);
  for(unsigned i = 2; i <= n; ++i)
    r *= T(i);
  return r;
}
}}
#endif
