// ob_special.cuh -- device special functions (replace AMOS zbesj/zbesh + boost Ynm on the hot path).
//   reference: srcAna/Bessel.h:58-143 (spherical j_n, h1_n and z'_n = -z_{n+1} + (n/z) z_n),
//              srcAna/TranslationAdditionCoefficients.cpp:64-68 (Y_nm, Condon-Shortley).
#pragma once
#include "ob_common.cuh"

namespace ob {

__host__ __device__ inline cplx cexp_i(cplx z) { // exp(i z)
  double e = exp(-z.y), s, c;
#ifdef __CUDA_ARCH__
  sincos(z.x, &s, &c);
#else
  s = sin(z.x);
  c = cos(z.x);
#endif
  return mk(e * c, e * s);
}
__host__ __device__ inline void csincos(cplx z, cplx &s, cplx &c) {
  double sx, cx;
#ifdef __CUDA_ARCH__
  sincos(z.x, &sx, &cx);
#else
  sx = sin(z.x);
  cx = cos(z.x);
#endif
  double ch = cosh(z.y), sh = sinh(z.y);
  s = mk(sx * ch, cx * sh);
  c = mk(cx * ch, -sx * sh);
}

// spherical Hankel of the first kind, orders 0..L, upward recurrence (h1 is the dominant
// solution for n > |z|, neutral below): h_{n+1} = (2n+1)/z h_n - h_{n-1}.  Single-valued in z
// (no branch cut), so h1(-conj(z)) = (-1)^n conj(h1(z)) holds exactly as the reference's
// AMOS + sqrt(pi/2z) path produces it.
// f(n, h_n) is called for n = 0 .. L in order (the orders stay in registers: an array indexed by a run-time order
// lives in local memory on the device)
template <class F> __host__ __device__ inline void sph_hankel1_each(cplx z, int L, F &f) {
  cplx e = cexp_i(z);
  cplx iz = cdiv(mk(1, 0), z);
  cplx h0 = cmul(mk(e.y, -e.x), iz);                          // -i e^{iz} / z
  cplx h1 = cneg(cmul(cmul(e, cadd(z, mk(0, 1))), cmul(iz, iz))); // -e^{iz} (z + i) / z^2
  f(0, h0);
  if(L >= 1)
    f(1, h1);
  for(int n = 1; n < L; ++n) {
    cplx t = csub(cscale(cmul(iz, h1), (double)(2 * n + 1)), h0);
    h0 = h1;
    h1 = t;
    f(n + 1, t);
  }
}
struct SphStoreOrders {
  cplx *out;
  __host__ __device__ void operator()(int n, cplx h) { out[n] = h; }
};
__host__ __device__ inline void sph_hankel1(cplx z, int L, cplx *out) {
  SphStoreOrders st = {out};
  sph_hankel1_each(z, L, st);
}

// spherical Bessel j_0..j_L for complex z != 0: Miller's downward recurrence with rescaling,
// normalised by the larger of j_0 = sin z / z, j_1 = sin z / z^2 - cos z / z.
__host__ __device__ inline void sph_bessel_j(cplx z, int L, cplx *out) {
  double az = cabs_(z);
  double mx = az > L ? az : (double)L;
  int nstart = (int)mx + 40 + (int)sqrt(40.0 * mx);
  cplx iz = cdiv(mk(1, 0), z);
  cplx fnp1 = mk(0, 0), fn = mk(1e-200, 0); // f_{n+1}, f_n at n = nstart
  for(int n = nstart; n >= 1; --n) {
    cplx fnm1 = csub(cscale(cmul(iz, fn), (double)(2 * n + 1)), fnp1);
    fnp1 = fn;
    fn = fnm1; // fn = f_{n-1}
    if(n - 1 <= L)
      out[n - 1] = fn;
    if(fabs(fn.x) + fabs(fn.y) > 1e150) {
      fn = cscale(fn, 1e-150);
      fnp1 = cscale(fnp1, 1e-150);
      for(int q = n - 1; q <= L; ++q)
        out[q] = cscale(out[q], 1e-150);
    }
  }
  cplx s, c;
  csincos(z, s, c);
  cplx j0 = cmul(s, iz);
  cplx j1 = csub(cmul(j0, iz), cmul(c, iz));
  cplx scale;
  if(L == 0 || cabs_(j0) >= cabs_(j1))
    scale = cdiv(j0, out[0]);
  else
    scale = cdiv(j1, out[1]);
  for(int n = 0; n <= L; ++n)
    out[n] = cmul(out[n], scale);
}

// Normalised associated Legendre (Condon-Shortley phase), 0 <= m <= l <= L, for one m:
//   N_l^m(x) = sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) P_l^m(x);  Y_lm(theta, 0) = N_l^m(cos theta).
// out[l] for l = m..L (entries below m untouched).  x = cos(theta), s = sin(theta) >= 0.
__host__ __device__ inline void legendre_norm_m(int m, int L, double x, double s, double *out) {
  double pmm = 0.28209479177387814347; // sqrt(1/(4 pi))
  for(int i = 1; i <= m; ++i)
    pmm *= -sqrt((double)(2 * i + 1) / (double)(2 * i)) * s;
  out[m] = pmm;
  if(L == m)
    return;
  double pmmp1 = x * sqrt((double)(2 * m + 3)) * pmm;
  out[m + 1] = pmmp1;
  for(int l = m + 2; l <= L; ++l) {
    double a = sqrt((double)(4 * l * l - 1) / (double)(l * l - m * m));
    double b = sqrt((double)((l - 1) * (l - 1) - m * m) / (double)(4 * (l - 1) * (l - 1) - 1));
    double pll = a * (x * pmmp1 - b * pmm);
    pmm = pmmp1;
    pmmp1 = pll;
    out[l] = pll;
  }
}

} // namespace ob
