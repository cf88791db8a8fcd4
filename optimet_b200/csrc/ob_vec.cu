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


// ---------------------------------------------------------------------------------------------
// Fused Arnoldi step: ONE cooperative launch per GMRES iteration (orthogonalisation of w against v_0..v_j,
// norm, normalisation into v_{j+1}) instead of 4 + 3 (j + 1) small launches.  The grid is at most one CTA per SM
// (co-resident by cudaLaunchCooperativeKernel), each thread keeps its <= AR_EPT elements of w in registers, and
// the grid meets at a self-resetting counter/generation barrier.  All sums have a fixed order (per-thread,
// warp tree, warps in order, blocks by a warp tree over b): bit-reproducible and identical on every rank.
//   mode 0: classical Gram-Schmidt with the DGKS second pass decided on the device (Belos "DGKS", restated);
//   mode 1: modified Gram-Schmidt, one vector at a time (Gmres_Zcomp, srcAna/PreconditionedMatrix.cpp:939-947).
// h_out: [0..j] = h_t, [j+1] = ||w||^2 before (mode 0), [j+2] = ||w||^2 after orthogonalisation.
// ---------------------------------------------------------------------------------------------
#define AR_THREADS 256
#define AR_EPT 8

struct ArnoldiArgs {
  const cplx *V;
  size_t ldv;
  int j, N, mode, n_harm;
  cplx *w, *vnext, *h_out, *partial; // partial: [2][(j + 3) * B], row j+2 = norm partials
  cplx *XP, *XS;                     // optional: pair-operator staging of v_{j+1} (ob_pairs.cu), or null
  unsigned *sync;                    // [0] arrival counter, [1] generation
};

__device__ __forceinline__ void grid_barrier(unsigned *sync, unsigned nblocks) {
  __syncthreads();
  if(threadIdx.x == 0) {
    unsigned gen, prev, cur;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(sync + 1) : "memory");
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(sync) : "memory");
    if(prev == nblocks - 1) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(sync), "r"(0u) : "memory");
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(sync + 1) : "memory");
    } else {
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(sync + 1) : "memory");
      } while(cur == gen);
    }
  }
  __syncthreads();
}

// block-wide sum of NV complex values per thread; result valid in thread 0.. (returned to every thread of warp 0)
template <int NV> __device__ __forceinline__ void block_sum(cplx (&v)[NV], double *sh /* [8][2 NV] */) {
#pragma unroll
  for(int q = 0; q < NV; ++q)
    for(int o = 16; o > 0; o >>= 1) {
      v[q].x += __shfl_down_sync(0xffffffffu, v[q].x, o);
      v[q].y += __shfl_down_sync(0xffffffffu, v[q].y, o);
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads(); // sh reuse
  if(lane == 0)
#pragma unroll
    for(int q = 0; q < NV; ++q) {
      sh[(warp * NV + q) * 2] = v[q].x;
      sh[(warp * NV + q) * 2 + 1] = v[q].y;
    }
  __syncthreads();
  if(threadIdx.x < NV) {
    double r = 0, im = 0;
    for(int wv = 0; wv < AR_THREADS / 32; ++wv) {
      r += sh[(wv * NV + threadIdx.x) * 2];
      im += sh[(wv * NV + threadIdx.x) * 2 + 1];
    }
    v[0] = mk(r, im); // thread q holds the total of value q
  }
}

// sum over blocks of partial[b], b < B, by warp 0 with a fixed tree; result broadcast through sh_out
__device__ __forceinline__ cplx sum_over_blocks(const cplx *partial, int B) {
  // executed by one full warp
  const int lane = threadIdx.x & 31;
  double r = 0, im = 0;
  for(int b = lane; b < B; b += 32) {
    const cplx p = __ldcg(partial + b);
    r += p.x;
    im += p.y;
  }
  for(int o = 16; o > 0; o >>= 1) {
    r += __shfl_down_sync(0xffffffffu, r, o);
    im += __shfl_down_sync(0xffffffffu, im, o);
  }
  r = __shfl_sync(0xffffffffu, r, 0);
  im = __shfl_sync(0xffffffffu, im, 0);
  return mk(r, im);
}

__global__ void __launch_bounds__(AR_THREADS, 1) k_arnoldi_step(ArnoldiArgs a) {
  __shared__ double sh[8 * 2 * 8];
  __shared__ cplx hs[260]; // h_t of the current pass (j + 1 <= 256) + scalars
  __shared__ double s_scal[4];
  const int B = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  const int N = a.N, j = a.j;
  const int chunk = (N + B - 1) / B;
  const int i0 = b * chunk, i1 = min(N, i0 + chunk);
  cplx wv[AR_EPT];
  int idx[AR_EPT];
#pragma unroll
  for(int k = 0; k < AR_EPT; ++k) {
    idx[k] = i0 + tid + k * AR_THREADS;
    wv[k] = idx[k] < i1 ? a.w[idx[k]] : mk(0, 0);
  }
  cplx *part0 = a.partial, *part1 = a.partial + (size_t)(j + 3) * B;
  double nb = 0, na = 0;

  auto local_norm2 = [&]() {
    cplx acc[1] = {mk(0, 0)};
#pragma unroll
    for(int k = 0; k < AR_EPT; ++k)
      if(idx[k] < i1)
        acc[0].x = fma(wv[k].x, wv[k].x, fma(wv[k].y, wv[k].y, acc[0].x));
    block_sum<1>(acc, sh);
    return acc[0];
  };

  if(a.mode == 0) {
    // ===== classical Gram-Schmidt (+ DGKS) =====
    for(int pass = 0; pass < 2; ++pass) {
      const int T = pass == 0 ? j + 2 : j + 1; // pass 0 also takes ||w||^2 as the "dot" with itself
      cplx *part = pass == 0 ? part0 : part1;
      for(int t0 = 0; t0 < T; t0 += 8) {
        cplx acc[8];
#pragma unroll
        for(int q = 0; q < 8; ++q)
          acc[q] = mk(0, 0);
#pragma unroll
        for(int k = 0; k < AR_EPT; ++k)
          if(idx[k] < i1) {
            cplx v[8]; // eight independent loads in flight per element
#pragma unroll
            for(int q = 0; q < 8; ++q) {
              const int t = t0 + q;
              v[q] = t <= j ? a.V[(size_t)t * a.ldv + idx[k]] : (t < T ? wv[k] : mk(0, 0));
            }
#pragma unroll
            for(int q = 0; q < 8; ++q) {
              acc[q].x = fma(v[q].x, wv[k].x, acc[q].x);
              acc[q].x = fma(v[q].y, wv[k].y, acc[q].x);
              acc[q].y = fma(v[q].x, wv[k].y, acc[q].y);
              acc[q].y = fma(-v[q].y, wv[k].x, acc[q].y);
            }
          }
        block_sum<8>(acc, sh);
        if(tid < 8 && t0 + tid < T)
          part[(size_t)(t0 + tid) * B + b] = acc[0];
      }
      grid_barrier(a.sync, B);
      // every block sums all T rows itself (warp w takes rows w, w + 8, ...; fixed tree over b): T * B partials
      // from L2 per block instead of a second grid-wide barrier
      for(int t = tid >> 5; t < T; t += AR_THREADS / 32) {
        const cplx s = sum_over_blocks(part + (size_t)t * B, B);
        if((tid & 31) == 0)
          hs[t] = s;
      }
      __syncthreads();
      if(pass == 0)
        nb = hs[j + 1].x;
      // w -= sum_t h_t v_t  (t ascending, as k_multi_axpy); loads issued eight at a time
#pragma unroll
      for(int k = 0; k < AR_EPT; ++k)
        if(idx[k] < i1) {
          cplx acc = wv[k];
          for(int t0 = 0; t0 <= j; t0 += 8) {
            cplx v[8];
#pragma unroll
            for(int q = 0; q < 8; ++q)
              v[q] = t0 + q <= j ? a.V[(size_t)(t0 + q) * a.ldv + idx[k]] : mk(0, 0);
#pragma unroll
            for(int q = 0; q < 8; ++q)
              if(t0 + q <= j) {
                const cplx c = hs[t0 + q];
                acc.x = fma(-c.x, v[q].x, acc.x);
                acc.x = fma(c.y, v[q].y, acc.x);
                acc.y = fma(-c.x, v[q].y, acc.y);
                acc.y = fma(-c.y, v[q].x, acc.y);
              }
          }
          wv[k] = acc;
        }
      // ||w||^2 after the pass
      {
        const cplx ln = local_norm2();
        cplx *pn = part + (size_t)(j + 2) * B; // own row: row j+1 (||w||^2 before) may still be read by slower blocks
        __syncthreads();
        if(tid == 0)
          pn[b] = ln;
        grid_barrier(a.sync, B);
        if(tid < 32) {
          const cplx s = sum_over_blocks(pn, B);
          if(tid == 0)
            s_scal[0] = s.x;
        }
        __syncthreads();
        na = s_scal[0];
      }
      if(b == 0) { // publish / accumulate h
        for(int t = tid; t <= j; t += AR_THREADS)
          a.h_out[t] = pass == 0 ? hs[t] : cadd(a.h_out[t], hs[t]);
        if(tid == 0) {
          if(pass == 0)
            a.h_out[j + 1] = mk(nb, 0);
          a.h_out[j + 2] = mk(na, 0);
        }
      }
      // DGKS: a second pass when the norm dropped by more than 1/sqrt(2)  (uniform over the grid)
      if(pass == 0 && !(sqrt(na) < 0.70710678118654752440 * sqrt(nb)))
        break;
      // pass 1 works in part1 (nothing of part0 is overwritten) and block 0 accumulates h_out with the same threads
    }
  } else {
    // ===== modified Gram-Schmidt =====
    for(int t = 0; t <= j; ++t) {
      cplx *part = (t & 1) ? part1 : part0;
      cplx acc[1] = {mk(0, 0)};
#pragma unroll
      for(int k = 0; k < AR_EPT; ++k)
        if(idx[k] < i1) {
          const cplx v = a.V[(size_t)t * a.ldv + idx[k]];
          acc[0].x = fma(v.x, wv[k].x, acc[0].x);
          acc[0].x = fma(v.y, wv[k].y, acc[0].x);
          acc[0].y = fma(v.x, wv[k].y, acc[0].y);
          acc[0].y = fma(-v.y, wv[k].x, acc[0].y);
        }
      block_sum<1>(acc, sh);
      if(tid == 0)
        part[b] = acc[0];
      grid_barrier(a.sync, B);
      if(tid < 32) {
        const cplx s = sum_over_blocks(part, B);
        if(tid == 0) {
          hs[0] = s;
          if(b == 0)
            a.h_out[t] = s;
        }
      }
      __syncthreads();
      const cplx c = hs[0];
#pragma unroll
      for(int k = 0; k < AR_EPT; ++k)
        if(idx[k] < i1) {
          const cplx v = a.V[(size_t)t * a.ldv + idx[k]];
          wv[k].x = fma(-c.x, v.x, wv[k].x);
          wv[k].x = fma(c.y, v.y, wv[k].x);
          wv[k].y = fma(-c.x, v.y, wv[k].y);
          wv[k].y = fma(-c.y, v.x, wv[k].y);
        }
      __syncthreads(); // hs[0] reuse
    }
    {
      const cplx ln = local_norm2();
      cplx *pn = (((j + 1) & 1) ? part1 : part0) + (size_t)(j + 2) * B;
      __syncthreads();
      if(tid == 0)
        pn[b] = ln;
      grid_barrier(a.sync, B);
      if(tid < 32) {
        const cplx s = sum_over_blocks(pn, B);
        if(tid == 0)
          s_scal[0] = s.x;
      }
      __syncthreads();
      na = s_scal[0];
      if(b == 0 && tid == 0) {
        a.h_out[j + 1] = mk(na, 0);
        a.h_out[j + 2] = mk(na, 0);
      }
    }
  }
  // ---- w (orthogonalised) back to memory, v_{j+1} = w / ||w|| (+ pair-operator staging of v_{j+1}) ----
  const double inv = 1.0 / sqrt(na);
#pragma unroll
  for(int k = 0; k < AR_EPT; ++k)
    if(idx[k] < i1) {
      a.w[idx[k]] = wv[k];
      const cplx v = cscale(wv[k], inv);
      a.vnext[idx[k]] = v;
      if(a.XP) {
        const int n = a.n_harm, p = idx[k] / (2 * n), e = idx[k] - p * 2 * n, half = e / n, c = e - half * n;
        int l = (int)sqrt((double)c + 1.0);
        while(l * l > c + 1)
          --l;
        while((l + 1) * (l + 1) <= c + 1)
          ++l;
        const double sg = (l & 1) ? -1.0 : 1.0;
        const size_t o = ((size_t)p * n + c) * 2 + half;
        a.XP[o] = v;
        a.XS[o] = mk(sg * v.x, sg * v.y);
      }
    }
}

// FP64 FMA peak of this device, measured (the assembly roofline denominator; MEASURED_PEAKS.json has no FP64 figure):
// 8 independent DFMA chains per thread, 8 CTAs of 256 threads per SM.
__global__ void k_dfma_peak(double *out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double m = 1.0000001, c = 1e-9;
#pragma unroll 4
  for(int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c);
    a1 = fma(a1, m, c);
    a2 = fma(a2, m, c);
    a3 = fma(a3, m, c);
    a4 = fma(a4, m, c);
    a5 = fma(a5, m, c);
    a6 = fma(a6, m, c);
    a7 = fma(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if(s == 12345.678)
    out[0] = s; // never true: keeps the chains alive
}
double measure_fp64_peak(int sm_count, cudaStream_t st) {
  double *d = nullptr;
  OB_CUDA(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  OB_CUDA(cudaEventCreate(&e0));
  OB_CUDA(cudaEventCreate(&e1));
  const int iters = 1 << 15, blocks = sm_count * 8, threads = 256;
  k_dfma_peak<<<blocks, threads, 0, st>>>(d, 1024, 0.5); // warm-up
  OB_CUDA(cudaEventRecord(e0, st));
  k_dfma_peak<<<blocks, threads, 0, st>>>(d, iters, 0.5);
  OB_CUDA(cudaEventRecord(e1, st));
  OB_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  OB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
  return flops / (ms * 1e-3) / 1e12;
}

bool arnoldi_fused_supported(int N, int sm_count) {
  const int B = std::max(1, std::min(sm_count, (N + AR_THREADS - 1) / AR_THREADS));
  return (N + B - 1) / B <= AR_THREADS * AR_EPT;
}
size_t arnoldi_scratch_elems(int N, int jmax, int sm_count) { return (size_t)2 * (jmax + 3) * sm_count + 16; }

void launch_arnoldi_step(const cplx *V, size_t ldv, int j, cplx *w, int N, int mode, cplx *h_out, cplx *partial,
                         unsigned *sync, cplx *vnext, cplx *XP, cplx *XS, int n_harm, int sm_count, cudaStream_t st) {
  if(j + 1 > 256)
    throw Error("fused Arnoldi step: Krylov basis larger than 256");
  ArnoldiArgs a;
  a.V = V;
  a.ldv = ldv;
  a.j = j;
  a.N = N;
  a.mode = mode;
  a.n_harm = n_harm;
  a.w = w;
  a.vnext = vnext;
  a.h_out = h_out;
  a.partial = partial;
  a.XP = XP;
  a.XS = XS;
  a.sync = sync;
  const int B = std::max(1, std::min(sm_count, (N + AR_THREADS - 1) / AR_THREADS));
  void *args[] = {&a};
  OB_CUDA(cudaLaunchCooperativeKernel((const void *)k_arnoldi_step, dim3(B), dim3(AR_THREADS), args, 0, st));
}

} // namespace ob
