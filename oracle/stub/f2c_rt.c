/* The two libf2c runtime routines AMOS references (d_sign, pow_dd), restated. */
#include <math.h>
double d_sign(const double *a, const double *b) {
  double x = (*a >= 0 ? *a : -*a);
  return *b >= 0 ? x : -x;
}
double pow_dd(const double *ap, const double *bp) { return pow(*ap, *bp); }
