// ob_lu.cu -- device-resident dense solve of the preconditioned scattering system (complex FP64).
//
// Replaces the reference's direct-solve routes on this path:
//   X_sca = S.colPivHouseholderQr().solve(Q)   srcAna/PreconditionedMatrixSolver.h:58,75  (serial, ACA off)
//   pzgesv_ on the block-cyclic matrix          srcAna/ScalapackSolver.cpp:54-128, scalapack/LinearSystemSolver.hpp:84-91
// Both are backward-stable dense factorisations; here it is a right-looking blocked LU with partial pivoting
// (pivot = largest modulus, lowest row on ties -- the rule of the CPU oracle's dense_solve) followed by the two
// triangular sweeps, zgesv-style: the right-hand side travels with the row interchanges.
//
// Mapping (NB = 64 columns per panel):
//   k_lu_panel       one cooperative launch per panel.  Rows are owned by threads for the whole panel; a pivot row
//                    "dies" in place instead of being swapped, so a column costs ONE grid barrier (pivot search of
//                    column j+1 is fused into the rank-1 update of column j).  Panel lives in L2.
//   k_lu_swaplist    turns the in-place pivot sequence into LAPACK's sequential interchanges (one thread, NB steps)
//   k_lu_laswp       applies them to all N columns and the right-hand side (one thread per column)
//   k_lu_linv        inverse of the unit-lower diagonal block (one CTA)
//   k_zgemm          register-tiled complex-FP64 GEMM (64x64 CTA tile, 4x4 per thread, BK = 16, register prefetch):
//                    U12 = Linv * A12 (in place) and the trailing update A22 -= L21 * U12 -- FP64-pipe bound
//   k_lu_fwd/bwd     block forward / backward substitution, one launch per diagonal block: every CTA re-solves the
//                    64x64 triangle from shared memory, then updates its own rows of the remaining vector
#include "ob_internal.h"
#include <cooperative_groups.h>

namespace ob {

namespace {

constexpr int NB = 64;
constexpr int PANEL_THREADS = 256;
constexpr int MAX_ROWS_PER_THREAD = 32;

// Grid barrier (all CTAs co-resident: cooperative launch).  sync[0] = arrival counter, sync[1] = generation; the
// caller carries the generation it expects in a register, so a barrier is one atomic plus the polling loads.
__device__ __forceinline__ void lu_grid_barrier(unsigned *sync, unsigned nblocks, unsigned &gen) {
  __syncthreads();
  if(threadIdx.x == 0) {
    unsigned prev, cur;
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
  gen += 1;
  __syncthreads();
}

struct PivCand {
  double v;
  int i;
};
__device__ __forceinline__ PivCand better(PivCand a, PivCand b) {
  if(b.v > a.v || (b.v == a.v && b.i < a.i))
    return b;
  return a;
}

// Panel factorisation of columns [k0, k0 + nb), rows [k0, N), without physical row swaps.
// pivrow[jj] = row chosen at step jj (a row of the panel in its ORIGINAL position); info = first zero pivot + 1.
__global__ void __launch_bounds__(PANEL_THREADS)
k_lu_panel(cplx *__restrict__ A, size_t lda, int N, int k0, int nb, double *part_v, int *part_i, unsigned *sync,
           int *pivrow, int *info) {
  __shared__ cplx u[NB];
  __shared__ double sv[PANEL_THREADS / 32];
  __shared__ int si[PANEL_THREADS / 32];
  __shared__ int s_piv;
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  const int GT = G * PANEL_THREADS;
  const int first = k0 + b * PANEL_THREADS + tid;
  unsigned dead = 0; // bit s: owned row first + s*GT has been used as a pivot
  unsigned gen;      // barrier generation at kernel entry (no CTA can have passed a barrier yet)
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(sync + 1) : "memory");
  // candidate for column 0
  PivCand cand{-1.0, 0x7fffffff};
  for(int s = 0, i = first; i < N; ++s, i += GT) {
    const cplx a = A[(size_t)k0 * lda + i];
    cand = better(cand, PivCand{cabs_(a), i});
  }
  for(int jj = 0; jj < nb; ++jj) {
    // ---- grid-wide arg max of the candidates ----
    for(int o = 16; o > 0; o >>= 1) {
      PivCand oth{__shfl_xor_sync(0xffffffffu, cand.v, o), __shfl_xor_sync(0xffffffffu, cand.i, o)};
      cand = better(cand, oth);
    }
    if((tid & 31) == 0) {
      sv[tid >> 5] = cand.v;
      si[tid >> 5] = cand.i;
    }
    __syncthreads();
    if(tid == 0) {
      PivCand c{sv[0], si[0]};
      for(int w = 1; w < PANEL_THREADS / 32; ++w)
        c = better(c, PivCand{sv[w], si[w]});
      // two slots (parity of jj): the next column's partials may be written while a slow block still reads these
      part_v[(jj & 1) * G + b] = c.v;
      part_i[(jj & 1) * G + b] = c.i;
    }
    lu_grid_barrier(sync, G, gen);
    if(tid < 32) {
      PivCand c{-1.0, 0x7fffffff};
      for(int q = tid; q < G; q += 32)
        c = better(c, PivCand{__ldcg(part_v + (jj & 1) * G + q), __ldcg(part_i + (jj & 1) * G + q)});
      for(int o = 16; o > 0; o >>= 1) {
        PivCand oth{__shfl_xor_sync(0xffffffffu, c.v, o), __shfl_xor_sync(0xffffffffu, c.i, o)};
        c = better(c, oth);
      }
      if(tid == 0)
        s_piv = c.i;
    }
    __syncthreads();
    const int p = s_piv;
    // pivot row (final since the previous step's update; written by another block -> read through L2)
    if(tid < nb && tid >= jj) {
      const double2 *src = reinterpret_cast<const double2 *>(A + (size_t)(k0 + tid) * lda + p);
      u[tid] = __ldcg(src);
    }
    __syncthreads();
    const cplx piv = u[jj];
    const bool singular = (piv.x == 0.0 && piv.y == 0.0);
    if(b == 0 && tid == 0) {
      pivrow[jj] = p;
      if(singular && *info == 0)
        *info = k0 + jj + 1;
    }
    const cplx inv = singular ? mk(0, 0) : cdiv(mk(1, 0), piv);
    cand = PivCand{-1.0, 0x7fffffff};
    for(int s = 0, i = first; i < N; ++s, i += GT) {
      if(i == p)
        dead |= 1u << s;
      if(dead & (1u << s))
        continue;
      cplx *row = A + (size_t)k0 * lda + i;
      // eight independent loads in flight per batch (one L2 round trip per batch instead of one per column); the
      // first batch travels with the multiplier's own element
      cplx a[8];
      const cplx araw = row[(size_t)jj * lda];
#pragma unroll
      for(int q = 0; q < 8; ++q)
        if(jj + 1 + q < nb)
          a[q] = row[(size_t)(jj + 1 + q) * lda];
      const cplx l = cmul(araw, inv);
      row[(size_t)jj * lda] = l;
      const cplx ml = cneg(l);
      for(int c0 = jj + 1; c0 < nb; c0 += 8) {
        if(c0 != jj + 1) {
#pragma unroll
          for(int q = 0; q < 8; ++q)
            if(c0 + q < nb)
              a[q] = row[(size_t)(c0 + q) * lda];
        }
#pragma unroll
        for(int q = 0; q < 8; ++q)
          if(c0 + q < nb) {
            cfma(a[q], ml, u[c0 + q]);
            row[(size_t)(c0 + q) * lda] = a[q];
          }
        if(c0 == jj + 1)
          cand = better(cand, PivCand{cabs_(a[0]), i});
      }
    }
    __syncthreads(); // u is rewritten in the next step
  }
}

// In-place pivot sequence -> sequential interchanges ipiv[k0 + jj] (LAPACK convention, absolute row index).
__global__ void k_lu_swaplist(const int *__restrict__ pivrow, int k0, int nb, int *__restrict__ ipiv) {
  if(threadIdx.x != 0 || blockIdx.x != 0)
    return;
  int where[NB], content[NB]; // where[r]: current position of original row k0+r; content[r]: original row at k0+r
  for(int r = 0; r < nb; ++r) {
    where[r] = k0 + r;
    content[r] = k0 + r;
  }
  for(int jj = 0; jj < nb; ++jj) {
    const int p = pivrow[jj];
    const int q = (p < k0 + nb) ? where[p - k0] : p;
    ipiv[k0 + jj] = q;
    const int o1 = content[jj]; // always an original row of the top block (direct solve)
    if(q != k0 + jj) {
      where[o1 - k0] = q;
      if(q < k0 + nb)
        content[q - k0] = o1;
      content[jj] = p;
      if(p < k0 + nb)
        where[p - k0] = k0 + jj;
    }
  }
}

// Row interchanges of one panel applied to every column (and the right-hand side as column N).
__global__ void k_lu_laswp(cplx *__restrict__ A, size_t lda, int N, int k0, int nb, const int *__restrict__ ipiv,
                           cplx *__restrict__ rhs) {
  __shared__ int sp[NB];
  if(threadIdx.x < nb)
    sp[threadIdx.x] = ipiv[k0 + threadIdx.x];
  __syncthreads();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if(col > N || (col == N && rhs == nullptr))
    return;
  cplx *base = col < N ? A + (size_t)col * lda : rhs;
  for(int jj = 0; jj < nb; ++jj) {
    const int q = sp[jj];
    if(q != k0 + jj) {
      const cplx t = base[k0 + jj];
      base[k0 + jj] = base[q];
      base[q] = t;
    }
  }
}

// Linv = inverse of the unit-lower nb x nb block at A(k0, k0); Linv is NB x NB column-major, zero above the diagonal.
__global__ void __launch_bounds__(256) k_lu_linv(const cplx *__restrict__ A, size_t lda, int k0, int nb,
                                                 cplx *__restrict__ Linv) {
  extern __shared__ cplx sm[];
  cplx *L = sm;            // [NB][NB] column-major
  cplx *X = sm + NB * NB;  // X[r * NB + j]: row r of column j (conflict-free across j)
  const int tid = threadIdx.x;
#pragma unroll 4
  for(int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx % NB, c = idx / NB;
    L[idx] = (r < nb && c < nb) ? A[(size_t)(k0 + c) * lda + k0 + r] : mk(0, 0);
  }
  __syncthreads();
  if(tid < NB) {
    const int j = tid;
    for(int r = 0; r < NB; ++r) {
      cplx x;
      if(r < j || r >= nb || j >= nb)
        x = mk(0, 0);
      else if(r == j)
        x = mk(1, 0);
      else {
        cplx x0 = mk(0, 0), x1 = mk(0, 0), x2 = mk(0, 0), x3 = mk(0, 0); // four independent FMA chains
        int s = j;
        for(; s + 3 < r; s += 4) {
          cfma(x0, L[s * NB + r], X[s * NB + j]);
          cfma(x1, L[(s + 1) * NB + r], X[(s + 1) * NB + j]);
          cfma(x2, L[(s + 2) * NB + r], X[(s + 2) * NB + j]);
          cfma(x3, L[(s + 3) * NB + r], X[(s + 3) * NB + j]);
        }
        for(; s < r; ++s)
          cfma(x0, L[s * NB + r], X[s * NB + j]);
        x = cneg(cadd(cadd(x0, x1), cadd(x2, x3)));
      }
      X[r * NB + j] = x;
    }
  }
  __syncthreads();
  for(int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx % NB, j = idx / NB;
    Linv[idx] = X[r * NB + j];
  }
}

// C (m x n) = alpha * A (m x K) * B (K x n) + beta * C, all column-major.  C may alias B when m <= 64 (one row tile):
// a CTA then reads its whole B panel before it writes.
constexpr int BM = 64, BN = 64, BK = 16;
__global__ void __launch_bounds__(256, 2) k_zgemm(int m, int n, int K, double alpha, const cplx *__restrict__ A,
                                               size_t lda, const cplx *B, size_t ldb, double beta, cplx *C,
                                               size_t ldc) {
  __shared__ cplx As[BK][BM];
  __shared__ cplx Bs[BK][BN + 1]; // + 1: the transposing store of a B chunk is 2-way instead of 16-way conflicted
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  cplx acc[4][4];
#pragma unroll
  for(int i = 0; i < 4; ++i)
#pragma unroll
    for(int j = 0; j < 4; ++j)
      acc[i][j] = mk(0, 0);
  cplx ra[4], rb[4];
  auto load = [&](int kk) {
#pragma unroll
    for(int q = 0; q < 4; ++q) {
      const int idx = tid + 256 * q;
      const int am = idx & 63, ak = idx >> 6;
      ra[q] = (m0 + am < m && kk + ak < K) ? A[(size_t)(kk + ak) * lda + m0 + am] : mk(0, 0);
      const int bk = idx & 15, bn = idx >> 4;
      rb[q] = (n0 + bn < n && kk + bk < K) ? B[(size_t)(n0 + bn) * ldb + kk + bk] : mk(0, 0);
    }
  };
  load(0);
  for(int kk = 0; kk < K; kk += BK) {
    __syncthreads();
#pragma unroll
    for(int q = 0; q < 4; ++q) {
      const int idx = tid + 256 * q;
      As[idx >> 6][idx & 63] = ra[q];
      Bs[idx & 15][idx >> 4] = rb[q];
    }
    __syncthreads();
    if(kk + BK < K)
      load(kk + BK);
#pragma unroll
    for(int k = 0; k < BK; ++k) {
      cplx a[4], bb[4];
#pragma unroll
      for(int i = 0; i < 4; ++i)
        a[i] = As[k][tx + 16 * i];
#pragma unroll
      for(int j = 0; j < 4; ++j)
        bb[j] = Bs[k][ty + 16 * j];
#pragma unroll
      for(int i = 0; i < 4; ++i)
#pragma unroll
        for(int j = 0; j < 4; ++j)
          cfma(acc[i][j], a[i], bb[j]);
    }
  }
#pragma unroll
  for(int j = 0; j < 4; ++j) {
    const int cn = n0 + ty + 16 * j;
    if(cn >= n)
      continue;
#pragma unroll
    for(int i = 0; i < 4; ++i) {
      const int cm = m0 + tx + 16 * i;
      if(cm >= m)
        continue;
      cplx *dst = C + (size_t)cn * ldc + cm;
      cplx r = cscale(acc[i][j], alpha);
      if(beta != 0.0) {
        const cplx old = *dst;
        r.x = fma(beta, old.x, r.x);
        r.y = fma(beta, old.y, r.y);
      }
      *dst = r;
    }
  }
}

// Forward substitution step of block k0: out[k0..k0+nb) = L11^-1 v[k0..k0+nb); v[i] -= L(i, k0..) * that, i >= k0+nb.
__global__ void __launch_bounds__(256) k_lu_fwd(const cplx *__restrict__ A, size_t lda, int N, int k0, int nb,
                                                cplx *v, cplx *__restrict__ out) {
  extern __shared__ cplx sm[];
  cplx *L = sm;           // [NB][NB] column-major
  cplx *y = sm + NB * NB; // [NB]
  const int tid = threadIdx.x;
  for(int idx = tid; idx < nb * NB; idx += 256) {
    const int r = idx % NB, c = idx / NB;
    L[idx] = r < nb ? A[(size_t)(k0 + c) * lda + k0 + r] : mk(0, 0);
  }
  if(tid < nb)
    y[tid] = v[k0 + tid];
  __syncthreads();
  for(int s = 0; s < nb - 1; ++s) {
    if(tid > s && tid < nb)
      cfma(y[tid], cneg(L[s * NB + tid]), y[s]);
    __syncthreads();
  }
  if(blockIdx.x == 0 && tid < nb)
    out[k0 + tid] = y[tid];
  const int i = k0 + nb + blockIdx.x * 256 + tid;
  if(i < N) {
    cplx acc0 = mk(0, 0), acc1 = mk(0, 0);
    const cplx *row = A + (size_t)k0 * lda + i;
    int s = 0;
    for(; s + 1 < nb; s += 2) {
      cfma(acc0, row[(size_t)s * lda], y[s]);
      cfma(acc1, row[(size_t)(s + 1) * lda], y[s + 1]);
    }
    if(s < nb)
      cfma(acc0, row[(size_t)s * lda], y[s]);
    v[i] = csub(v[i], cadd(acc0, acc1));
  }
}

// Backward substitution step of block k0: out[k0..k0+nb) = U11^-1 v[k0..k0+nb); v[i] -= U(i, k0..) * that, i < k0.
__global__ void __launch_bounds__(256) k_lu_bwd(const cplx *__restrict__ A, size_t lda, int k0, int nb, cplx *v,
                                                cplx *__restrict__ out) {
  extern __shared__ cplx sm[];
  cplx *U = sm;
  cplx *x = sm + NB * NB;
  const int tid = threadIdx.x;
  for(int idx = tid; idx < nb * NB; idx += 256) {
    const int r = idx % NB, c = idx / NB;
    U[idx] = r < nb ? A[(size_t)(k0 + c) * lda + k0 + r] : mk(0, 0);
  }
  if(tid < nb)
    x[tid] = v[k0 + tid];
  __syncthreads();
  for(int s = nb - 1; s >= 0; --s) {
    if(tid == s)
      x[s] = cdiv(x[s], U[s * NB + s]);
    __syncthreads();
    if(tid < s)
      cfma(x[tid], cneg(U[s * NB + tid]), x[s]);
    __syncthreads();
  }
  if(blockIdx.x == 0 && tid < nb)
    out[k0 + tid] = x[tid];
  const int i = blockIdx.x * 256 + tid;
  if(i < k0) {
    cplx acc0 = mk(0, 0), acc1 = mk(0, 0);
    const cplx *row = A + (size_t)k0 * lda + i;
    int s = 0;
    for(; s + 1 < nb; s += 2) {
      cfma(acc0, row[(size_t)s * lda], x[s]);
      cfma(acc1, row[(size_t)(s + 1) * lda], x[s + 1]);
    }
    if(s < nb)
      cfma(acc0, row[(size_t)s * lda], x[s]);
    v[i] = csub(v[i], cadd(acc0, acc1));
  }
}

} // namespace

void LuWork::alloc(int N, int sm_count) {
  if(cap >= N && ipiv)
    return;
  release();
  cap = N;
  OB_CUDA(cudaMalloc(&ipiv, sizeof(int) * (size_t)std::max(N, 1)));
  OB_CUDA(cudaMalloc(&pivrow, sizeof(int) * NB));
  OB_CUDA(cudaMalloc(&info, sizeof(int)));
  OB_CUDA(cudaMalloc(&sync, sizeof(unsigned) * 4));
  OB_CUDA(cudaMemset(sync, 0, sizeof(unsigned) * 4));
  const int gmax = std::max(sm_count, 1) * 4;
  OB_CUDA(cudaMalloc(&part_v, sizeof(double) * 2 * gmax));
  OB_CUDA(cudaMalloc(&part_i, sizeof(int) * 2 * gmax));
  OB_CUDA(cudaMalloc(&Linv, sizeof(cplx) * NB * NB));
  OB_CUDA(cudaMalloc(&v1, sizeof(cplx) * (size_t)std::max(N, 1)));
  OB_CUDA(cudaMalloc(&v2, sizeof(cplx) * (size_t)std::max(N, 1)));
}
void LuWork::release() {
  void *ptrs[] = {ipiv, pivrow, info, sync, part_v, part_i, Linv, v1, v2};
  for(void *p : ptrs)
    if(p)
      cudaFree(p);
  ipiv = pivrow = info = nullptr;
  sync = nullptr;
  part_v = nullptr;
  part_i = nullptr;
  Linv = v1 = v2 = nullptr;
  cap = 0;
}

// A (N x N, column-major, overwritten by its LU factors), b -> x (device vectors, may alias).  Returns LAPACK's info
// (0, or the 1-based index of the first exactly-zero pivot).  launches is incremented by the kernels enqueued.
int lu_solve(cplx *A, int N, size_t lda, LuWork &w, const cplx *b, cplx *x, int sm_count, cudaStream_t st,
             long &launches) {
  if(N <= 0)
    return 0;
  w.alloc(N, sm_count);
  static bool attr_set[64] = {false};
  const size_t sm_linv = sizeof(cplx) * 2 * NB * NB, sm_sub = sizeof(cplx) * (NB * NB + NB);
  if(first_use_on_device(attr_set)) {
    OB_CUDA(cudaFuncSetAttribute(k_lu_linv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_linv));
    OB_CUDA(cudaFuncSetAttribute(k_lu_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_sub));
    OB_CUDA(cudaFuncSetAttribute(k_lu_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_sub));
  }
  int occ = 1;
  OB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lu_panel, PANEL_THREADS, 0));
  const int gmax = std::max(1, std::min(occ, 4) * sm_count);
  OB_CUDA(cudaMemsetAsync(w.info, 0, sizeof(int), st));
  OB_CUDA(cudaMemcpyAsync(w.v1, b, sizeof(cplx) * (size_t)N, cudaMemcpyDeviceToDevice, st));
  for(int k0 = 0; k0 < N; k0 += NB) {
    const int nb = std::min(NB, N - k0), M = N - k0;
    // fewest CTAs that keep <= 1 row per thread (cheap barrier); beyond the co-resident limit, several rows per thread
    int G = std::min(gmax, (M + PANEL_THREADS - 1) / PANEL_THREADS);
    if((long)G * PANEL_THREADS * MAX_ROWS_PER_THREAD < M)
      throw Error("direct solve: matrix too large for the panel kernel");
    {
      size_t lda_ = lda;
      int N_ = N, k0_ = k0, nb_ = nb;
      void *args[] = {&A, &lda_, &N_, &k0_, &nb_, &w.part_v, &w.part_i, &w.sync, &w.pivrow, &w.info};
      OB_CUDA(cudaLaunchCooperativeKernel((void *)k_lu_panel, dim3(G), dim3(PANEL_THREADS), args, 0, st));
    }
    k_lu_swaplist<<<1, 32, 0, st>>>(w.pivrow, k0, nb, w.ipiv);
    k_lu_laswp<<<(N + 1 + 127) / 128, 128, 0, st>>>(A, lda, N, k0, nb, w.ipiv, w.v1);
    launches += 3;
    const int rest = N - k0 - nb;
    if(rest > 0) {
      k_lu_linv<<<1, 256, sm_linv, st>>>(A, lda, k0, nb, w.Linv);
      // U12 = Linv * A12 (in place: one row tile)
      cplx *A12 = A + (size_t)(k0 + nb) * lda + k0;
      k_zgemm<<<dim3(1, (rest + BN - 1) / BN), 256, 0, st>>>(nb, rest, nb, 1.0, w.Linv, NB, A12, lda, 0.0, A12, lda);
      // A22 -= L21 * U12
      const cplx *L21 = A + (size_t)k0 * lda + k0 + nb;
      cplx *A22 = A + (size_t)(k0 + nb) * lda + k0 + nb;
      k_zgemm<<<dim3((rest + BM - 1) / BM, (rest + BN - 1) / BN), 256, 0, st>>>(rest, rest, nb, -1.0, L21, lda, A12,
                                                                                lda, 1.0, A22, lda);
      launches += 3;
    }
  }
  OB_CUDA(cudaGetLastError());
  // L y = P b
  for(int k0 = 0; k0 < N; k0 += NB) {
    const int nb = std::min(NB, N - k0), rest = N - k0 - nb;
    k_lu_fwd<<<std::max(1, (rest + 255) / 256), 256, sm_sub, st>>>(A, lda, N, k0, nb, w.v1, w.v2);
    launches += 1;
  }
  // U x = y
  for(int k0 = ((N - 1) / NB) * NB; k0 >= 0; k0 -= NB) {
    const int nb = std::min(NB, N - k0);
    k_lu_bwd<<<std::max(1, (k0 + 255) / 256), 256, sm_sub, st>>>(A, lda, k0, nb, w.v2, x);
    launches += 1;
  }
  OB_CUDA(cudaGetLastError());
  int info = 0;
  OB_CUDA(cudaMemcpyAsync(&info, w.info, sizeof(int), cudaMemcpyDeviceToHost, st));
  OB_CUDA(cudaStreamSynchronize(st));
  return info;
}

} // namespace ob
