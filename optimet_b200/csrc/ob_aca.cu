// ob_aca.cu -- ACA-compressed operator on the device.
//
// Replaces, for <ACA compression="yes">:
//   Scattering_matrix_ACA_FF / _SH   srcAna/PreconditionedMatrix.cpp:489-551, 699-759
//   ACA_compression, getMaxInd        srcAna/PreconditionedMatrix.cpp:760-889
//   matvec (the Gmres_Zcomp operator) srcAna/PreconditionedMatrix.cpp:1058-1085
//
// Each ordered particle pair (i, j) with distance >= 2 (r_i + r_j) is stored as U (dim x r) and V (r x dim) from the
// reference's partially pivoted cross approximation, everything else (near blocks) stays dense and the diagonal is the
// identity.  The compression follows the reference's arithmetic literally -- pivot rule (first largest |.| among the
// rows / columns not used yet), residual updates accumulated from zero in pivot order, the norm recursion with its
// truncated cross term, eps = 1e-3 -- because the pivot sequence decides the result at the 1e-3 level: FMA contraction
// is switched off in the residual updates so that the values the pivot search sees are the ones a CPU run sees.
//
// Build: block-rows are assembled in batches into a dense scratch slab by k_assemble (ob_vtac.cu), one CTA per
// admissible block runs the cross approximation out of that slab into a full-rank scratch, the ranks go to the host,
// and a pack kernel copies U, V (and the dense near blocks) into an exactly sized allocation per batch.
#include "ob_internal.h"
#include <algorithm>
#include <cstring>

namespace ob {

#define ACA_THREADS 256
#define ACA_MAXD (2 * OB_MAX_FLAT)

// ---------------------------------------------------------------------------------------------
// exact (non-contracted) complex helpers: the compiler must not fuse these into FMAs
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ cplx xmul(cplx a, cplx b) {
  return mk(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__device__ __forceinline__ cplx xadd(cplx a, cplx b) { return mk(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }
__device__ __forceinline__ cplx xsub(cplx a, cplx b) { return mk(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y)); }
// std::complex<double> operator/ as the reference's compiler evaluates it: libgcc __divdc3 (Smith's ratio form, this
// operation order) without contraction.  libgcc additionally rescales operands below 2^-52 / above DBL_MAX/2 by exact
// powers of two, which does not change the result in the normal range.
__device__ __forceinline__ cplx xdiv(cplx A, cplx B) {
  const double a = A.x, b = A.y, c = B.x, d = B.y;
  if(fabs(c) < fabs(d)) {
    const double ratio = __ddiv_rn(c, d), denom = __dadd_rn(__dmul_rn(c, ratio), d);
    return mk(__ddiv_rn(__dadd_rn(__dmul_rn(a, ratio), b), denom), __ddiv_rn(__dsub_rn(__dmul_rn(b, ratio), a), denom));
  }
  const double ratio = __ddiv_rn(d, c), denom = __dadd_rn(__dmul_rn(d, ratio), c);
  return mk(__ddiv_rn(__dadd_rn(__dmul_rn(b, ratio), a), denom), __ddiv_rn(__dsub_rn(b, __dmul_rn(a, ratio)), denom));
}
// std::abs(std::complex<double>) = hypot as glibc >= 2.35 evaluates it without FMA (sysdeps/ieee754/dbl-64/e_hypot.c:
// sqrt of the sum of squares plus one correction step); checked bit for bit against glibc 2.39 on 6e7 random pairs.
// The pivot search compares these magnitudes, so they must be the ones a CPU run of the reference sees.
__device__ __forceinline__ double xhypot_kernel(double ax, double ay) {
  double t1, t2;
  double h = __dsqrt_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
  if(h <= __dmul_rn(2.0, ay)) {
    const double delta = __dsub_rn(h, ay);
    t1 = __dmul_rn(ax, __dsub_rn(__dmul_rn(2.0, delta), ax));
    t2 = __dmul_rn(__dsub_rn(delta, __dmul_rn(2.0, __dsub_rn(ax, ay))), delta);
  } else {
    const double delta = __dsub_rn(h, ax);
    t1 = __dmul_rn(__dmul_rn(2.0, delta), __dsub_rn(ax, __dmul_rn(2.0, ay)));
    t2 = __dadd_rn(__dmul_rn(__dsub_rn(__dmul_rn(4.0, delta), ay), ay), __dmul_rn(delta, delta));
  }
  return __dsub_rn(h, __ddiv_rn(__dadd_rn(t1, t2), __dmul_rn(2.0, h)));
}
__device__ __forceinline__ double xhypot(double x, double y) {
  const double SCALE = 0x1p-600, LARGE = 0x1p+511, TINY = 0x1p-459, EPS = 0x1p-54;
  x = fabs(x);
  y = fabs(y);
  const double ax = x < y ? y : x, ay = x < y ? x : y;
  if(!(ax <= 1.7976931348623157e308))
    return hypot(x, y); // inf / nan: never a pivot
  if(ax > LARGE) {
    if(ay <= __dmul_rn(ax, EPS))
      return __dadd_rn(ax, ay);
    return __ddiv_rn(xhypot_kernel(__dmul_rn(ax, SCALE), __dmul_rn(ay, SCALE)), SCALE);
  }
  if(ay < TINY) {
    if(ax >= __ddiv_rn(ay, EPS))
      return __dadd_rn(ax, ay);
    return __dmul_rn(xhypot_kernel(__ddiv_rn(ax, SCALE), __ddiv_rn(ay, SCALE)), SCALE);
  }
  if(ax >= __ddiv_rn(ay, EPS))
    return __dadd_rn(ax, ay);
  return xhypot_kernel(ax, ay);
}

struct Best {
  double v;
  int i;
};
__device__ __forceinline__ Best better(Best a, Best b) { // larger value, then smaller index (= first occurrence)
  if(b.v > a.v || (b.v == a.v && b.i < a.i))
    return b;
  return a;
}
// getMaxInd: first index of the largest |v[i]| > 0 among the entries with used[i] == 0; -1 when there is none
__device__ int block_argmax(const cplx *v, const unsigned char *used, int dim, Best *s_best) {
  Best b;
  b.v = 0.0;
  b.i = 0x7fffffff;
  for(int i = threadIdx.x; i < dim; i += blockDim.x) {
    if(used[i])
      continue;
    double a = xhypot(v[i].x, v[i].y);
    if(a > b.v) { // strict: ascending i inside a thread keeps the first occurrence
      b.v = a;
      b.i = i;
    }
  }
  for(int o = 16; o > 0; o >>= 1) {
    Best t;
    t.v = __shfl_xor_sync(0xffffffffu, b.v, o);
    t.i = __shfl_xor_sync(0xffffffffu, b.i, o);
    b = better(b, t);
  }
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if((threadIdx.x & 31) == 0)
    s_best[w] = b;
  __syncthreads();
  if(threadIdx.x == 0) {
    Best r = s_best[0];
    for(int k = 1; k < nw; ++k)
      r = better(r, s_best[k]);
    s_best[0] = r;
  }
  __syncthreads();
  Best r = s_best[0];
  __syncthreads();
  return (r.v > 0.0 && r.i != 0x7fffffff) ? r.i : -1;
}
__device__ double block_sum(double v, double *s_red) {
  for(int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if((threadIdx.x & 31) == 0)
    s_red[w] = v;
  __syncthreads();
  if(threadIdx.x == 0) {
    double r = 0;
    for(int k = 0; k < nw; ++k)
      r += s_red[k];
    s_red[0] = r;
  }
  __syncthreads();
  double r = s_red[0];
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// ACA_compression (PreconditionedMatrix.cpp:760-859): one CTA per admissible block.
//   C(r, c) = slab[(j dim + c) ld + il dim + r];  U(i, p) = U[p dim + i];  V(p, q) = V[p dim + q]
// rank_out: rank >= 2, or -(k+1) when no pivot qualifies at step k (reference behaviour undefined there).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ACA_THREADS)
k_aca_compress(const cplx *__restrict__ slab, size_t ld, const int2 *__restrict__ jobs, int nobj, int dim, double eps,
               cplx *scrU, cplx *scrV, int *__restrict__ rank_out, int *__restrict__ piv_out) {
  __shared__ cplx s_row[ACA_MAXD], s_col[ACA_MAXD];
  __shared__ double s_pom[ACA_MAXD];
  __shared__ unsigned char usedI[ACA_MAXD], usedJ[ACA_MAXD];
  __shared__ Best s_best[ACA_THREADS / 32];
  __shared__ double s_red[ACA_THREADS / 32];
  __shared__ int s_stop;
  __shared__ double s_norma;
  const int job = blockIdx.x;
  const int il = jobs[job].x, j = jobs[job].y;
  const cplx *C = slab + (size_t)j * dim * ld + (size_t)il * dim;
  cplx *U = scrU + (size_t)job * dim * dim;
  cplx *V = scrV + (size_t)job * dim * dim;
  int *PI = piv_out + ((size_t)il * nobj + j) * 2 * dim, *PJ = PI + dim; // block order
  for(int i = threadIdx.x; i < dim; i += blockDim.x)
    usedI[i] = usedJ[i] = 0;
  __syncthreads();
  int I = 0, rank = 0;
  for(int k = 0; k < dim; ++k) {
    // residual row I: CoupMat.row(I) - sum_p U(I,p) V.row(p), accumulated from zero in pivot order (:795-803)
    for(int q = threadIdx.x; q < dim; q += blockDim.x) {
      cplx c = C[(size_t)q * ld + I];
      if(k > 0) {
        cplx s = mk(0, 0);
        for(int p = 0; p < k; ++p)
          s = xadd(s, xmul(U[(size_t)p * dim + I], V[(size_t)p * dim + q]));
        c = xsub(c, s);
      }
      s_row[q] = c;
    }
    __syncthreads();
    const int J = block_argmax(s_row, usedJ, dim, s_best);
    if(J < 0) {
      rank = -(k + 1);
      break;
    }
    const cplx piv = s_row[J];
    __syncthreads();
    for(int q = threadIdx.x; q < dim; q += blockDim.x) {
      cplx v = xdiv(s_row[q], piv);
      V[(size_t)k * dim + q] = v;
      s_row[q] = v;
    }
    if(threadIdx.x == 0) {
      usedJ[J] = 1;
      usedI[I] = 1;
      PJ[k] = J;
      PI[k] = I;
    }
    __syncthreads();
    // residual column J: CoupMat.col(J) - sum_p V(p,J) U.col(p) (:810-818)
    for(int i = threadIdx.x; i < dim; i += blockDim.x) {
      cplx c = C[(size_t)J * ld + i];
      if(k > 0) {
        cplx s = mk(0, 0);
        for(int p = 0; p < k; ++p)
          s = xadd(s, xmul(V[(size_t)p * dim + J], U[(size_t)p * dim + i]));
        c = xsub(c, s);
      }
      U[(size_t)k * dim + i] = c;
      s_col[i] = c;
    }
    __syncthreads();
    double c2 = 0, r2 = 0;
    for(int i = threadIdx.x; i < dim; i += blockDim.x) {
      c2 += cnorm(s_col[i]);
      r2 += cnorm(s_row[i]);
    }
    c2 = block_sum(c2, s_red);
    r2 = block_sum(r2, s_red);
    const double cn = sqrt(c2), rn = sqrt(r2);
    // cross term as written (:825-838): p = 0 .. k-2, first k entries only, no conjugation
    if(k > 0) {
      for(int p = threadIdx.x; p < k - 1; p += blockDim.x) {
        cplx pom1 = mk(0, 0), pom2 = mk(0, 0);
        for(int tt = 0; tt < k; ++tt) {
          pom1 = xadd(pom1, xmul(U[(size_t)p * dim + tt], s_col[tt]));
          pom2 = xadd(pom2, xmul(V[(size_t)p * dim + tt], s_row[tt]));
        }
        s_pom[p] = __dmul_rn(xhypot(pom1.x, pom1.y), xhypot(pom2.x, pom2.y));
      }
    }
    __syncthreads();
    if(threadIdx.x == 0) {
      if(k == 0) {
        s_norma = (cn * cn) * (rn * rn);
        s_stop = 0;
      } else {
        double sum = 0.0;
        for(int p = 0; p < k - 1; ++p)
          sum = sum + s_pom[p];
        s_norma = s_norma + (cn * cn) * (rn * rn) + 2.0 * sum;
        s_stop = (eps * sqrt(s_norma) >= cn * rn) ? 1 : 0;
      }
    }
    __syncthreads();
    rank = k + 1;
    if(s_stop || k + 1 == dim)
      break;
    I = block_argmax(s_col, usedI, dim, s_best);
    if(I < 0) {
      rank = -(k + 2);
      break;
    }
  }
  if(threadIdx.x == 0)
    rank_out[job] = rank;
}

// copy U (dim x r) and V (r x dim) of every low-rank job out of the full-rank scratch
__global__ void k_aca_pack_lr(const int2 *__restrict__ jobs, const AcaDesc *__restrict__ desc, int nobj, int il0,
                              int dim, const cplx *__restrict__ scrU, const cplx *__restrict__ scrV) {
  const int job = blockIdx.x;
  const AcaDesc d = desc[(size_t)(il0 + jobs[job].x) * nobj + jobs[job].y];
  const size_t cnt = (size_t)d.rank * dim;
  const cplx *su = scrU + (size_t)job * dim * dim, *sv = scrV + (size_t)job * dim * dim;
  for(size_t e = threadIdx.x; e < cnt; e += blockDim.x) {
    d.U[e] = su[e];
    d.V[e] = sv[e];
  }
}
// copy the dense near blocks out of the slab (column-major dim x dim, ld = dim in the packed form)
__global__ void k_aca_pack_dense(const int2 *__restrict__ jobs, const AcaDesc *__restrict__ desc, int nobj, int il0,
                                 int dim, const cplx *__restrict__ slab, size_t ld) {
  const int job = blockIdx.x;
  const int il = jobs[job].x, j = jobs[job].y;
  const AcaDesc d = desc[(size_t)(il0 + il) * nobj + j];
  const cplx *C = slab + (size_t)j * dim * ld + (size_t)il * dim;
  for(int e = threadIdx.x; e < dim * dim; e += blockDim.x) {
    int c = e / dim, r = e - c * dim;
    d.U[e] = C[(size_t)c * ld + r];
  }
}

// ---------------------------------------------------------------------------------------------
// matvec (PreconditionedMatrix.cpp:1058-1085): y_i = sum_j [ U_ij (V_ij x_j) | S_ij x_j | x_i ]
// grid (chunks of j, local block-rows); each thread owns rows tid and tid + 256 of the block-row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ACA_THREADS)
k_matvec_aca(const AcaDesc *__restrict__ desc, int nobj, int dim, int nch, const cplx *__restrict__ x,
             cplx *__restrict__ out, size_t out_stride) {
  __shared__ cplx sx[ACA_MAXD], st[ACA_MAXD];
  const int il = blockIdx.y, ch = blockIdx.x;
  const int j0 = (int)((long)ch * nobj / nch), j1 = (int)((long)(ch + 1) * nobj / nch);
  const int r0 = threadIdx.x, r1 = threadIdx.x + ACA_THREADS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = ACA_THREADS / 32;
  cplx a0 = mk(0, 0), a1 = mk(0, 0);
  for(int j = j0; j < j1; ++j) {
    const AcaDesc d = desc[(size_t)il * nobj + j];
    __syncthreads();
    for(int q = threadIdx.x; q < dim; q += ACA_THREADS)
      sx[q] = x[(size_t)j * dim + q];
    __syncthreads();
    if(d.rank == 0) { // identity diagonal block
      if(r0 < dim)
        a0 = cadd(a0, sx[r0]);
      if(r1 < dim)
        a1 = cadd(a1, sx[r1]);
    } else if(d.rank < 0) { // dense near block
      const cplx *S = d.U;
      for(int q = 0; q < dim; ++q) {
        const cplx xq = sx[q];
        if(r0 < dim)
          cfma(a0, S[(size_t)q * dim + r0], xq);
        if(r1 < dim)
          cfma(a1, S[(size_t)q * dim + r1], xq);
      }
    } else {
      for(int p = w; p < d.rank; p += nw) { // t = V x_j, one warp per row of V
        const cplx *Vp = d.V + (size_t)p * dim;
        cplx s = mk(0, 0);
        for(int q = lane; q < dim; q += 32)
          cfma(s, Vp[q], sx[q]);
        for(int o = 16; o > 0; o >>= 1) {
          s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
          s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        }
        if(lane == 0)
          st[p] = s;
      }
      __syncthreads();
      const cplx *Um = d.U;
      for(int p = 0; p < d.rank; ++p) {
        const cplx tp = st[p];
        if(r0 < dim)
          cfma(a0, Um[(size_t)p * dim + r0], tp);
        if(r1 < dim)
          cfma(a1, Um[(size_t)p * dim + r1], tp);
      }
    }
  }
  cplx *o = out + (size_t)ch * out_stride + (size_t)il * dim;
  if(r0 < dim)
    o[r0] = a0;
  if(r1 < dim)
    o[r1] = a1;
}
__global__ void k_aca_reduce(const cplx *__restrict__ partial, int nch, size_t M, cplx *__restrict__ y) {
  size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(r >= M)
    return;
  cplx s = partial[r];
  for(int c = 1; c < nch; ++c)
    s = cadd(s, partial[(size_t)c * M + r]);
  y[r] = s;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <class T> static T *dev_alloc(size_t n) {
  T *p = nullptr;
  OB_CUDA(cudaMalloc(&p, std::max<size_t>(1, n) * sizeof(T)));
  return p;
}

void AcaScratch::release() {
  void *ptrs[] = {slab, scrU, scrV, d_jobs, d_jobs_dense, d_rank};
  for(void *p : ptrs)
    if(p)
      cudaFree(p);
  slab = scrU = scrV = nullptr;
  d_jobs = d_jobs_dense = nullptr;
  d_rank = nullptr;
  elems = blocks = 0;
}
void AcaOperator::release() {
  for(cplx *p : chunks)
    cudaFree(p);
  chunks.clear();
  chunk_cap.clear();
  void *ptrs[] = {desc, piv, partial};
  for(void *p : ptrs)
    if(p)
      cudaFree(p);
  desc = nullptr;
  piv = nullptr;
  partial = nullptr;
  partial_elems = desc_blocks = 0;
  h_desc.clear();
  built = false;
}

// admissibility criterion of PreconditionedMatrix.cpp:526 / :734 / :1070 with Tools::findDistance (Tools.cpp:34-38)
static bool admissible(const double *xyz, const double *radius, int i, int j) {
  double dx = xyz[3 * j] - xyz[3 * i], dy = xyz[3 * j + 1] - xyz[3 * i + 1], dz = xyz[3 * j + 2] - xyz[3 * i + 2];
  double distance = std::sqrt(std::pow(dx, 2.0) + std::pow(dy, 2.0) + std::pow(dz, 2.0));
  return distance >= 2.0 * (radius[i] + radius[j]);
}

void aca_build(AcaOperator &op, AcaScratch &scr, VtacTableSet const &ts, const double *d_xyz, const cplx *Tdiag, cplx k, int nobj,
               int first, int count, const double *h_xyz, const double *h_radius, double eps, size_t budget_bytes,
               int sm_count, cudaStream_t st, long &launches) {
  // device buffers persist across builds (a wavelength sweep rebuilds the operator at every step): scratch, job lists
  // and the per-batch storage chunks are reused whenever they are large enough
  const int dim = 2 * flat_max(ts.nMax);
  const size_t nblocks = (size_t)count * nobj;
  if(op.dim != dim || op.nobj != nobj || op.count != count)
    op.release();
  op.built = false;
  op.nobj = nobj;
  op.dim = dim;
  op.first = first;
  op.count = count;
  op.h_desc.assign(nblocks, AcaDesc());
  if(op.desc_blocks < nblocks) {
    if(op.desc)
      cudaFree(op.desc);
    if(op.piv)
      cudaFree(op.piv);
    op.desc = dev_alloc<AcaDesc>(nblocks);
    op.piv = dev_alloc<int>(nblocks * 2 * dim);
    op.desc_blocks = nblocks;
  }
  OB_CUDA(cudaMemsetAsync(op.piv, 0xff, nblocks * 2 * dim * sizeof(int), st));
  // ~8 CTAs per SM: the U (V x) blocks are short dependent phases, latency is hidden by residency, not by ILP
  const int nch = std::max(1, std::min(nobj, (8 * sm_count + count - 1) / std::max(1, count)));
  if(op.partial && (nch != op.nch || op.partial_elems < (size_t)nch * count * dim)) {
    cudaFree(op.partial);
    op.partial = nullptr;
  }
  op.nch = nch;
  if(nch > 1 && !op.partial) {
    op.partial_elems = (size_t)nch * count * dim;
    op.partial = dev_alloc<cplx>(op.partial_elems);
  }
  op.stored_elems = 0;
  op.n_lowrank = op.n_dense = 0;
  op.rank_sum = 0;
  op.rank_max = 0;
  // batch of block-rows: slab (rows x N) + two full-rank scratch copies of it
  const size_t blk = (size_t)dim * dim;
  const size_t per_row = 3 * blk * nobj * sizeof(cplx);
  int rows_b = (int)std::max<size_t>(1, std::min<size_t>(count, budget_bytes / std::max<size_t>(1, per_row)));
  if(scr.blocks < (size_t)rows_b * nobj || scr.elems < (size_t)rows_b * nobj * blk) { // shared by both harmonics
    scr.release();
    scr.blocks = (size_t)rows_b * nobj;
    scr.elems = scr.blocks * blk;
    scr.slab = dev_alloc<cplx>(scr.elems);
    scr.scrU = dev_alloc<cplx>(scr.elems);
    scr.scrV = dev_alloc<cplx>(scr.elems);
    scr.d_jobs = dev_alloc<int2>(scr.blocks);
    scr.d_jobs_dense = dev_alloc<int2>(scr.blocks);
    scr.d_rank = dev_alloc<int>(scr.blocks);
  }
  cplx *slab = scr.slab, *scrU = scr.scrU, *scrV = scr.scrV;
  int2 *d_jobs = scr.d_jobs, *d_jobs_dense = scr.d_jobs_dense;
  int *d_rank = scr.d_rank;
  size_t batch_index = 0;
  std::vector<int2> jobs, jobs_dense;
  std::vector<int> ranks;
  std::string fail;
  for(int il0 = 0; il0 < count && fail.empty(); il0 += rows_b) {
    const int rows = std::min(rows_b, count - il0);
    const size_t ld = (size_t)rows * dim;
    launch_assemble(ts, d_xyz, Tdiag, k, nobj, first + il0, rows, slab, ld, st);
    launches += 1;
    jobs.clear();
    jobs_dense.clear();
    for(int il = 0; il < rows; ++il)
      for(int j = 0; j < nobj; ++j) {
        const int i = first + il0 + il;
        if(i == j)
          continue;
        if(admissible(h_xyz, h_radius, i, j))
          jobs.push_back(make_int2(il, j));
        else
          jobs_dense.push_back(make_int2(il, j));
      }
    ranks.assign(jobs.size(), 0);
    if(!jobs.empty()) {
      OB_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
      k_aca_compress<<<(unsigned)jobs.size(), ACA_THREADS, 0, st>>>(slab, ld, d_jobs, nobj, dim, eps, scrU, scrV, d_rank,
                                                                    op.piv + (size_t)il0 * nobj * 2 * dim);
      OB_CUDA(cudaGetLastError());
      launches += 1;
      OB_CUDA(cudaMemcpyAsync(ranks.data(), d_rank, jobs.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    if(!jobs_dense.empty())
      OB_CUDA(cudaMemcpyAsync(d_jobs_dense, jobs_dense.data(), jobs_dense.size() * sizeof(int2),
                              cudaMemcpyHostToDevice, st));
    OB_CUDA(cudaStreamSynchronize(st));
    // exact-size allocation for this batch
    size_t total = jobs_dense.size() * blk;
    for(size_t t = 0; t < jobs.size(); ++t) {
      if(ranks[t] < 2) {
        fail = "ACA_compression: no admissible pivot in block (" + std::to_string(first + il0 + jobs[t].x) + ", " +
               std::to_string(jobs[t].y) + ") at step " + std::to_string(-ranks[t] - 1) +
               " (the reference reads an uninitialised index there)";
        break;
      }
      total += 2 * (size_t)ranks[t] * dim;
    }
    if(!fail.empty())
      break;
    // storage chunk of this batch: reuse the previous build's allocation when it is large enough (5 % slack on new ones)
    if(batch_index < op.chunks.size() && op.chunk_cap[batch_index] < total) {
      cudaFree(op.chunks[batch_index]);
      op.chunks[batch_index] = nullptr;
    }
    if(batch_index >= op.chunks.size()) {
      op.chunks.push_back(nullptr);
      op.chunk_cap.push_back(0);
    }
    if(!op.chunks[batch_index]) {
      op.chunk_cap[batch_index] = total + total / 20;
      op.chunks[batch_index] = dev_alloc<cplx>(op.chunk_cap[batch_index]);
    }
    cplx *chunk = op.chunks[batch_index];
    ++batch_index;
    size_t off = 0;
    for(size_t t = 0; t < jobs.size(); ++t) {
      AcaDesc &d = op.h_desc[(size_t)(il0 + jobs[t].x) * nobj + jobs[t].y];
      d.rank = ranks[t];
      d.U = chunk + off;
      d.V = d.U + (size_t)ranks[t] * dim;
      off += 2 * (size_t)ranks[t] * dim;
      op.n_lowrank += 1;
      op.rank_sum += ranks[t];
      op.rank_max = std::max(op.rank_max, ranks[t]);
    }
    for(size_t t = 0; t < jobs_dense.size(); ++t) {
      AcaDesc &d = op.h_desc[(size_t)(il0 + jobs_dense[t].x) * nobj + jobs_dense[t].y];
      d.rank = -1;
      d.U = chunk + off;
      d.V = nullptr;
      off += blk;
      op.n_dense += 1;
    }
    op.stored_elems += (double)total;
    OB_CUDA(cudaMemcpyAsync(op.desc + (size_t)il0 * nobj, op.h_desc.data() + (size_t)il0 * nobj,
                            (size_t)rows * nobj * sizeof(AcaDesc), cudaMemcpyHostToDevice, st));
    if(!jobs.empty()) {
      k_aca_pack_lr<<<(unsigned)jobs.size(), 256, 0, st>>>(d_jobs, op.desc, nobj, il0, dim, scrU, scrV);
      OB_CUDA(cudaGetLastError());
      launches += 1;
    }
    if(!jobs_dense.empty()) {
      k_aca_pack_dense<<<(unsigned)jobs_dense.size(), 256, 0, st>>>(d_jobs_dense, op.desc, nobj, il0, dim, slab, ld);
      OB_CUDA(cudaGetLastError());
      launches += 1;
    }
    OB_CUDA(cudaStreamSynchronize(st)); // jobs / ranks are reused by the next batch
  }
  while(op.chunks.size() > batch_index) { // fewer batches than the previous build
    cudaFree(op.chunks.back());
    op.chunks.pop_back();
    op.chunk_cap.pop_back();
  }
  if(!fail.empty()) {
    op.release();
    throw Error(fail);
  }
  op.built = true;
}

void launch_matvec_aca(AcaOperator const &op, const cplx *x, cplx *y_slice, cudaStream_t st, cudaEvent_t e0,
                       cudaEvent_t e1) {
  const size_t M = (size_t)op.count * op.dim;
  if(e0)
    cudaEventRecord(e0, st);
  if(op.count > 0) {
    dim3 grid(op.nch, op.count);
    k_matvec_aca<<<grid, ACA_THREADS, 0, st>>>(op.desc, op.nobj, op.dim, op.nch, x, op.nch > 1 ? op.partial : y_slice,
                                               M);
    OB_CUDA(cudaGetLastError());
    if(op.nch > 1) {
      k_aca_reduce<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(op.partial, op.nch, M, y_slice);
      OB_CUDA(cudaGetLastError());
    }
  }
  if(e1)
    cudaEventRecord(e1, st);
}

// unit surface: ACA_compression of one caller-supplied dim x dim column-major block (device pointers)
void aca_compress_single(const cplx *C_dev, int dim, double eps, cplx *U_dev, cplx *V_dev, int *rank_dev, int *piv_dev,
                         cudaStream_t st) {
  int2 job = make_int2(0, 0);
  int2 *d_job = dev_alloc<int2>(1);
  OB_CUDA(cudaMemcpyAsync(d_job, &job, sizeof(int2), cudaMemcpyHostToDevice, st));
  k_aca_compress<<<1, ACA_THREADS, 0, st>>>(C_dev, (size_t)dim, d_job, 1, dim, eps, U_dev, V_dev, rank_dev, piv_dev);
  OB_CUDA(cudaGetLastError());
  OB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_job);
}

} // namespace ob
