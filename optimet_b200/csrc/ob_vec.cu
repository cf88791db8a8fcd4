// ob_vec.cu -- Krylov vector kernels for the GMRES drivers (Arnoldi dots / updates / norms).
// All reductions are two-stage with a fixed block order: bit-reproducible run to run and identical on
// every rank of a row-sharded solve (the Krylov vectors are replicated).
//   reference: srcAna/PreconditionedMatrix.cpp:939-947 (modified Gram-Schmidt of Gmres_Zcomp) and the
//   orthogonalisation inside Belos (DGKS, restated).
#include "ob_internal.h"

namespace ob {

static const int kDotThreads = 256;
static inline int dot_blocks(int N) {
  int b = (N + 4 * kDotThreads - 1) / (4 * kDotThreads);
  return std::max(1, std::min(b, 296));
}
size_t vec_scratch_elems(int N, int jmax) { return (size_t)dot_blocks(N) * (jmax + 2) + 16; }

// partial[t * B + b] = sum_{i in chunk b} conj(V_t[i]) w[i]
__global__ void k_dot_partial(const cplx *__restrict__ V, size_t ldv, const cplx *__restrict__ w, int N,
                              cplx *__restrict__ partial) {
  __shared__ double sre[kDotThreads / 32], sim[kDotThreads / 32];
  const int b = blockIdx.x, B = gridDim.x, t = blockIdx.y;
  const int chunk = (N + B - 1) / B;
  const int i0 = b * chunk, i1 = min(N, i0 + chunk);
  const cplx *v = V + (size_t)t * ldv;
  double re = 0, im = 0;
  for(int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    cplx a = v[i], c = w[i];
    re = fma(a.x, c.x, re);
    re = fma(a.y, c.y, re);
    im = fma(a.x, c.y, im);
    im = fma(-a.y, c.x, im);
  }
  for(int o = 16; o > 0; o >>= 1) {
    re += __shfl_down_sync(0xffffffffu, re, o);
    im += __shfl_down_sync(0xffffffffu, im, o);
  }
  if((threadIdx.x & 31) == 0) {
    sre[threadIdx.x >> 5] = re;
    sim[threadIdx.x >> 5] = im;
  }
  __syncthreads();
  if(threadIdx.x == 0) {
    double r = 0, q = 0;
    for(int k = 0; k < kDotThreads / 32; ++k) {
      r += sre[k];
      q += sim[k];
    }
    partial[(size_t)t * B + b] = mk(r, q);
  }
}
__global__ void k_dot_final(const cplx *__restrict__ partial, int B, cplx *__restrict__ h) {
  const int t = blockIdx.x;
  if(threadIdx.x == 0) {
    double r = 0, q = 0;
    for(int b = 0; b < B; ++b) {
      cplx p = partial[(size_t)t * B + b];
      r += p.x;
      q += p.y;
    }
    h[t] = mk(r, q);
  }
}
void launch_multi_dot(const cplx *V, size_t ldv, int j, const cplx *w, int N, cplx *h_dev, cplx *scratch,
                      cudaStream_t st) {
  if(j <= 0)
    return;
  const int B = dot_blocks(N);
  k_dot_partial<<<dim3(B, j), kDotThreads, 0, st>>>(V, ldv, w, N, scratch);
  k_dot_final<<<j, 32, 0, st>>>(scratch, B, h_dev);
  OB_CUDA(cudaGetLastError());
}

__global__ void k_multi_axpy(const cplx *__restrict__ V, size_t ldv, int j, const cplx *__restrict__ h,
                             cplx *__restrict__ w, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= N)
    return;
  cplx acc = w[i];
  for(int t = 0; t < j; ++t) {
    cplx c = h[t], v = V[(size_t)t * ldv + i];
    acc.x = fma(-c.x, v.x, acc.x);
    acc.x = fma(c.y, v.y, acc.x);
    acc.y = fma(-c.x, v.y, acc.y);
    acc.y = fma(-c.y, v.x, acc.y);
  }
  w[i] = acc;
}
void launch_multi_axpy(const cplx *V, size_t ldv, int j, const cplx *h_dev, cplx *w, int N, cudaStream_t st) {
  if(j <= 0)
    return;
  k_multi_axpy<<<(N + 255) / 256, 256, 0, st>>>(V, ldv, j, h_dev, w, N);
  OB_CUDA(cudaGetLastError());
}

void launch_norm2(const cplx *w, int N, double *out_dev, double *scratch, cudaStream_t st) {
  // ||w||^2 as the real part of w^H w; out_dev receives a complex (re, 0) pair in two doubles
  launch_multi_dot(w, 0, 1, w, N, (cplx *)out_dev, (cplx *)scratch, st);
}

__global__ void k_scale_to(const cplx *__restrict__ w, double inv, cplx *__restrict__ v, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < N)
    v[i] = cscale(w[i], inv);
}
void launch_scale_to(const cplx *w, double inv, cplx *v, int N, cudaStream_t st) {
  k_scale_to<<<(N + 255) / 256, 256, 0, st>>>(w, inv, v, N);
  OB_CUDA(cudaGetLastError());
}

__global__ void k_axpby(cplx a, const cplx *__restrict__ x, cplx b, const cplx *__restrict__ y, cplx *__restrict__ z,
                        int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < N) {
    cplx r = mk(0, 0);
    if(x)
      r = cmul(a, x[i]);
    if(y)
      r = cadd(r, cmul(b, y[i]));
    z[i] = r;
  }
}
void launch_axpby(cplx a, const cplx *x, cplx b, const cplx *y, cplx *z, int N, cudaStream_t st) {
  k_axpby<<<(N + 255) / 256, 256, 0, st>>>(a, x, b, y, z, N);
  OB_CUDA(cudaGetLastError());
}

__global__ void k_combine(const cplx *__restrict__ V, size_t ldv, int j, const cplx *__restrict__ c,
                          cplx *__restrict__ x, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= N)
    return;
  cplx acc = x[i];
  for(int t = 0; t < j; ++t)
    cfma(acc, c[t], V[(size_t)t * ldv + i]);
  x[i] = acc;
}
void launch_combine(const cplx *V, size_t ldv, int j, const cplx *coef_dev, cplx *x, int N, cudaStream_t st) {
  if(j <= 0)
    return;
  k_combine<<<(N + 255) / 256, 256, 0, st>>>(V, ldv, j, coef_dev, x, N);
  OB_CUDA(cudaGetLastError());
}

// out = a .* b  (- c if c != null), optionally conjugated  (Solver.cpp:68-69, :107-110 and the
// conj(X_int) fed to the SH sources, PreconditionedMatrixSolver.h:66-68)
__global__ void k_hadamard(const cplx *__restrict__ a, const cplx *__restrict__ b, const cplx *__restrict__ c,
                           cplx *__restrict__ out, int N, int conj_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= N)
    return;
  cplx r = cmul(a[i], b[i]);
  if(c)
    r = csub(r, c[i]);
  out[i] = conj_out ? cconj(r) : r;
}
void launch_hadamard(const cplx *a, const cplx *b, const cplx *c, cplx *out, int N, int conj_out, cudaStream_t st) {
  k_hadamard<<<(N + 255) / 256, 256, 0, st>>>(a, b, c, out, N, conj_out);
  OB_CUDA(cudaGetLastError());
}

} // namespace ob
