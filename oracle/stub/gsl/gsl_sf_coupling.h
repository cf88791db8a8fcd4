/* Stand-in for <gsl/gsl_sf_coupling.h> (GSL is absent from the image): Wigner 3j / 6j / 9j symbols with GSL's
 * signatures (arguments are TWICE the angular momenta), implemented in stub/gsl_coupling_min.cpp from the Racah
 * formulas GSL documents; the same formulas are checked against sympy's exact values in tests/test_oracle_kats.py. */
#ifndef OB_STUB_GSL_SF_COUPLING_H
#define OB_STUB_GSL_SF_COUPLING_H
#ifdef __cplusplus
extern "C" {
#endif
double gsl_sf_coupling_3j(int two_ja, int two_jb, int two_jc, int two_ma, int two_mb, int two_mc);
double gsl_sf_coupling_6j(int two_ja, int two_jb, int two_jc, int two_jd, int two_je, int two_jf);
double gsl_sf_coupling_9j(int two_ja, int two_jb, int two_jc, int two_jd, int two_je, int two_jf, int two_jg, int two_jh,
                          int two_ji);
#ifdef __cplusplus
}
#endif
#endif
