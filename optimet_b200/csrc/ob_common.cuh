// ob_common.cuh -- shared device/host helpers for the B200 multiple-scattering path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace ob {

typedef double2 cplx; // (re, im), 16-byte aligned: one LDG/STG.128 per element

__host__ __device__ __forceinline__ cplx mk(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return mk(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cneg(cplx a) { return mk(-a.x, -a.y); }
__host__ __device__ __forceinline__ double cnorm(cplx a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ __forceinline__ cplx cmuli(cplx a) { return mk(-a.y, a.x); } // i*a
// acc += a*b
__host__ __device__ __forceinline__ void cfma(cplx &acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// Smith's algorithm (robust against overflow in |b|^2)
__host__ __device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  if(fabs(b.x) >= fabs(b.y)) {
    double r = b.y / b.x, d = b.x + b.y * r;
    return mk((a.x + a.y * r) / d, (a.y - a.x * r) / d);
  } else {
    double r = b.x / b.y, d = b.x * r + b.y;
    return mk((a.x * r + a.y) / d, (a.y * r - a.x) / d);
  }
}
__host__ __device__ __forceinline__ double cabs_(cplx a) { return hypot(a.x, a.y); }
__host__ __device__ __forceinline__ cplx csqrt_(cplx z) { // principal branch
  double m = hypot(z.x, z.y);
  if(m == 0.0)
    return mk(0, 0);
  double s = sqrt(0.5 * (m + fabs(z.x)));
  double t = 0.5 * z.y / s;
  if(z.x >= 0)
    return mk(s, t);
  return mk(fabs(t), z.y >= 0 ? s : -s);
}

// flat harmonic index conventions of the reference (CompoundIterator.h:24-31):
//   p = n(n+1) - m - 1  (m descending inside n),  n = floor(sqrt(p+1)),  m = n(n+1) - p - 1
__host__ __device__ __forceinline__ int flat_index(int n, int m) { return n * (n + 1) - m - 1; }
__host__ __device__ __forceinline__ int flat_max(int nMax) { return nMax * (nMax + 2); }
__host__ __device__ __forceinline__ void unflatten(int p, int &n, int &m) {
  n = (int)sqrt((double)p + 1.0);
  while(n * n > p + 1) --n; // guard against rounding
  while((n + 1) * (n + 1) <= p + 1) ++n;
  m = n * (n + 1) - p - 1;
}

struct Error : public std::runtime_error {
  explicit Error(std::string const &m) : std::runtime_error(m) {}
};

// true exactly once per device of this process (per-device state such as __constant__ uploads and kernel attributes
// must be set on every device a single-process multi-GPU group uses: ob_create_multi)
inline bool first_use_on_device(bool (&flags)[64]) {
  int dev = 0;
  if(cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
    return true;
  if(flags[dev])
    return false;
  flags[dev] = true;
  return true;
}

#define OB_CUDA(call)                                                                                                  \
  do {                                                                                                                 \
    cudaError_t e__ = (call);                                                                                          \
    if(e__ != cudaSuccess)                                                                                             \
      throw ob::Error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" +              \
                      std::to_string(__LINE__));                                                                       \
  } while(0)

} // namespace ob
