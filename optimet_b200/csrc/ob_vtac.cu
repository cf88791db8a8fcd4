// ob_vtac.cu -- kernels built on vtac_block(): block assembly of the preconditioned coupling
// matrix, single-displacement A/B (unit parity with Coupling), regular-translation apply
// (incident-wave localisation) and the |T|^2 |x|^2 scattering reduction.
#include "ob_internal.h"
#include "ob_vtac.cuh"

namespace ob {

// ---------------------------------------------------------------------------------------------
// host: index-only tables (depend on nMax alone)
// ---------------------------------------------------------------------------------------------
static inline bool valid_nm(int n, int m) { return n >= 0 && std::abs(m) <= n; }
// TranslationAdditionCoefficients.cpp:35-61
static double h_a_plus(int n, int m) {
  return valid_nm(n, m) ? -std::sqrt((double)((n + m + 1) * (n - m + 1)) / (double)((2 * n + 1) * (2 * n + 3))) : 0.0;
}
static double h_a_minus(int n, int m) {
  return valid_nm(n, m) ? std::sqrt((double)((n + m) * (n - m)) / (double)((2 * n + 1) * (2 * n - 1))) : 0.0;
}
static double h_b_plus(int n, int m) {
  return valid_nm(n, m) ? std::sqrt((double)((n + m + 2) * (n + m + 1)) / (double)((2 * n + 1) * (2 * n + 3))) : 0.0;
}
static double h_b_minus(int n, int m) {
  return valid_nm(n, m) ? std::sqrt((double)((n - m) * (n - m - 1)) / (double)((2 * n + 1) * (2 * n - 1))) : 0.0;
}

void VtacTableSet::build(int NM) {
  release();
  nMax = NM;
  const int L = 2 * NM;
  std::vector<double> ap((L + 2) * (L + 2), 0.0), am(ap), bp(ap), bm(ap);
  for(int l = 0; l <= L + 1; ++l)
    for(int k = -l; k <= l; ++k) {
      int i = l * (l + 1) + k;
      ap[i] = h_a_plus(l, k);
      am[i] = h_a_minus(l, k);
      bp[i] = h_b_plus(l, k);
      bm[i] = h_b_minus(l, k);
    }
  std::vector<double> inv_ap((NM + 1) * (NM + 1), 0.0), am_nm(inv_ap), inv_bp(NM + 1, 0.0);
  for(int n = 1; n <= NM; ++n) {
    inv_bp[n] = 1.0 / h_b_plus(n - 1, n - 1);
    for(int m = 0; m < n; ++m) {
      inv_ap[n * (NM + 1) + m] = 1.0 / h_a_plus(n - 1, m);
      am_nm[n * (NM + 1) + m] = h_a_minus(n - 1, m);
    }
  }
  // Legendre recurrence coefficients (ob_special.cuh: legendre_norm_m), tabulated
  std::vector<double> legA((L + 1) * (L + 2) / 2 + 1, 0.0), legB(legA), legD(L + 2, 0.0), legF(L + 2, 0.0);
  for(int m = 0; m <= L; ++m) {
    legD[m] = std::sqrt((double)(2 * m + 3));
    if(m >= 1)
      legF[m] = -std::sqrt((double)(2 * m + 1) / (double)(2 * m));
    for(int l = m + 2; l <= L; ++l) {
      legA[l * (l + 1) / 2 + m] = std::sqrt((double)(4 * l * l - 1) / (double)(l * l - m * m));
      legB[l * (l + 1) / 2 + m] = std::sqrt((double)((l - 1) * (l - 1) - m * m) / (double)(4 * (l - 1) * (l - 1) - 1));
    }
  }
  std::vector<int> off(NM + 2, 0);
  for(int m = 0; m <= NM; ++m)
    off[m + 1] = off[m] + (L - m + 1) * (L - m + 1);
  const int nr = flat_max(NM);
  std::vector<double> rowc(8 * nr, 0.0), colc(4 * nr, 0.0);
  for(int p = 0; p < nr; ++p) {
    int l, k;
    unflatten(p, l, k);
    double *rc = &rowc[8 * p];
    rc[0] = 0.5 / std::sqrt((double)(l * (l + 1)));                                           // A factor, l part
    rc[1] = 0.5 * std::sqrt((double)(2 * l + 1) / (double)((2 * l - 1) * l * (l + 1)));       // B factor, l part
    rc[2] = std::sqrt((double)((l - k) * (l + k + 1)));                                       // with beta(n,m+1,l,k+1)
    rc[3] = std::sqrt((double)((l + k) * (l - k + 1)));                                       // with beta(n,m-1,l,k-1)
    rc[4] = std::sqrt((double)((l - k) * (l + k)));                                           // B: beta(n,m,l-1,k)
    rc[5] = std::sqrt((double)((l - k) * (l - k - 1)));                                       // B: beta(n,m+1,l-1,k+1)
    rc[6] = std::sqrt((double)((l + k) * (l + k - 1)));                                       // B: beta(n,m-1,l-1,k-1)
    double *cc = &colc[4 * p]; // same (n, m) parametrisation for columns
    cc[0] = 1.0 / std::sqrt((double)(l * (l + 1)));
    cc[1] = std::sqrt((double)((l - k) * (l + k + 1)));
    cc[2] = std::sqrt((double)((l + k) * (l - k + 1)));
  }
  auto up = [&](const void *src, size_t bytes) {
    void *d = nullptr;
    OB_CUDA(cudaMalloc(&d, bytes));
    OB_CUDA(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    allocs.push_back(d);
    return d;
  };
  tb.NM = NM;
  tb.L = L;
  tb.T = off[NM + 1];
  tb.ap = (const double *)up(ap.data(), ap.size() * 8);
  tb.am = (const double *)up(am.data(), am.size() * 8);
  tb.bp = (const double *)up(bp.data(), bp.size() * 8);
  tb.bm = (const double *)up(bm.data(), bm.size() * 8);
  tb.inv_ap_nm = (const double *)up(inv_ap.data(), inv_ap.size() * 8);
  tb.am_nm = (const double *)up(am_nm.data(), am_nm.size() * 8);
  tb.inv_bp_n = (const double *)up(inv_bp.data(), inv_bp.size() * 8);
  tb.legA = (const double *)up(legA.data(), legA.size() * 8);
  tb.legB = (const double *)up(legB.data(), legB.size() * 8);
  tb.legD = (const double *)up(legD.data(), legD.size() * 8);
  tb.legF = (const double *)up(legF.data(), legF.size() * 8);
  tb.off = (const int *)up(off.data(), off.size() * 4);
  tb.rowc = (const double *)up(rowc.data(), rowc.size() * 8);
  tb.colc = (const double *)up(colc.data(), colc.size() * 8);
  // k_translate_apply reuses the buffers for its [groups][2][rows] reduction (<= THREADS * 2 complex)
  smem = std::max(vtac_smem_bytes(NM), (size_t)OB_VTAC_THREADS * 2 * sizeof(cplx));
}
void VtacTableSet::release() {
  for(void *p : allocs)
    cudaFree(p);
  allocs.clear();
  nMax = -1;
}

// Cartesian displacement -> (r, theta, phi) exactly as Spherical.h:64-71 (no FMA contraction, so the
// arguments of acos/atan2 carry the same roundings as the host code of the reference)
__device__ __forceinline__ void to_spherical(double x, double y, double z, double &r, double &the, double &phi) {
  double r2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
  r = sqrt(r2);
  if(r > 0.0) {
    the = acos(z / r);
    phi = atan2(y, x);
  } else {
    the = 0.0;
    phi = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// K1: block assembly.  grid = (N_obj columns, local rows); one CTA per particle-pair block.
//   S(block i,j) = -T_i [[A^T, B^T],[B^T, A^T]],  A,B = Coupling(R_i - R_j, k, nMax)  (irregular)
//   (srcAna/PreconditionedMatrix.cpp:384-390 FF, :592-598 SH); identity on the diagonal (:379).
// Output: column-major slab, ld = local rows, 16-byte streaming stores coalesced along the row index.
// ---------------------------------------------------------------------------------------------
struct EmitStore {
  cplx *base; // top-left element of this block
  size_t ld;
  int n;
  cplx tTE, tTM; // -T_i for this thread's row (TE, TM halves)
  __device__ __forceinline__ void item(int p, int r, cplx A, cplx B) {
    cplx *c0 = base + (size_t)p * ld + r;
    cplx *c1 = base + (size_t)(p + n) * ld + r;
    __stcs(c0, cmul(tTE, A));
    __stcs(c0 + n, cmul(tTM, B));
    __stcs(c1, cmul(tTE, B));
    __stcs(c1 + n, cmul(tTM, A));
  }
};

__global__ void __launch_bounds__(OB_VTAC_THREADS, 2)
k_assemble(VtacTables tb, const double *__restrict__ xyz, const cplx *__restrict__ Tdiag, cplx k, int row0,
           cplx *__restrict__ S, size_t ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int j = blockIdx.x, il = blockIdx.y, i = row0 + il;
  const int n = flat_max(tb.NM);
  cplx *base = S + (size_t)j * 2 * n * ld + (size_t)il * 2 * n;
  if(i == j) { // PreconditionedMatrix.cpp:379
    for(int e = threadIdx.x; e < 4 * n * n; e += blockDim.x) {
      int c = e / (2 * n), r = e - c * 2 * n;
      __stcs(base + (size_t)c * ld + r, mk(r == c ? 1.0 : 0.0, 0.0));
    }
    return;
  }
  double r, the, phi;
  to_spherical(xyz[3 * i] - xyz[3 * j], xyz[3 * i + 1] - xyz[3 * j + 1], xyz[3 * i + 2] - xyz[3 * j + 2], r, the, phi);
  EmitStore em;
  em.base = base;
  em.ld = ld;
  em.n = n;
  const int gs = (n + 31) & ~31;
  const int row = threadIdx.x % gs;
  em.tTE = em.tTM = mk(0, 0);
  if(row < n) {
    em.tTE = cneg(Tdiag[(size_t)i * 2 * n + row]);
    em.tTM = cneg(Tdiag[(size_t)i * 2 * n + n + row]);
  }
  vtac_block(tb, smem_raw, r, the, phi, k, false, em);
}

// ---------------------------------------------------------------------------------------------
// K1 (pair form): one CTA per local pair (i < j); stores the unscaled A^T, B^T (n x n each, rows = harmonics of
// particle i fastest) -- see ob_pairs.cu.  Same vtac_block, half the CTAs and a quarter of the bytes of k_assemble.
// ---------------------------------------------------------------------------------------------
struct EmitPair {
  cplx *A, *B;
  int n;
  __device__ __forceinline__ void item(int p, int r, cplx a, cplx b) {
    __stcs(A + (size_t)p * n + r, a);
    __stcs(B + (size_t)p * n + r, b);
  }
};
// MINB = resident CTAs per SM the register allocation is bounded for: 3 (56 registers, no spills) when three level-buffer
// sets fit the shared memory of an SM (nMax <= 8: 58 KB per CTA), else 2.
template <int MINB>
__global__ void __launch_bounds__(OB_VTAC_THREADS, MINB)
k_assemble_pairs(VtacTables tb, const double *__restrict__ xyz, cplx k, const int2 *__restrict__ pair_ij,
                 cplx *__restrict__ AB) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int2 ij = pair_ij[blockIdx.x];
  const int i = ij.x, j = ij.y;
  const int n = flat_max(tb.NM);
  double r, the, phi;
  to_spherical(xyz[3 * i] - xyz[3 * j], xyz[3 * i + 1] - xyz[3 * j + 1], xyz[3 * i + 2] - xyz[3 * j + 2], r, the, phi);
  EmitPair em;
  em.A = AB + (size_t)blockIdx.x * 2 * n * n;
  em.B = em.A + (size_t)n * n;
  em.n = n;
  vtac_block(tb, smem_raw, r, the, phi, k, false, em);
}

// ---------------------------------------------------------------------------------------------
// single displacement -> A, B (n x n, column-major, A[p + q n] = Coupling.diagonal(p, q))
// ---------------------------------------------------------------------------------------------
struct EmitAB {
  cplx *A, *B;
  int n;
  __device__ __forceinline__ void item(int p, int r, cplx a, cplx b) {
    A[p + (size_t)r * n] = a;
    B[p + (size_t)r * n] = b;
  }
};
__global__ void __launch_bounds__(OB_VTAC_THREADS)
k_vtac_single(VtacTables tb, double r, double the, double phi, cplx k, int regular, cplx *A, cplx *B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = flat_max(tb.NM);
  if(fabs(r) < 1e-10) { // Coupling.cpp:82-84
    for(int e = threadIdx.x; e < n * n; e += blockDim.x) {
      A[e] = mk((e / n) == (e % n) ? 1.0 : 0.0, 0.0);
      B[e] = mk(0, 0);
    }
    return;
  }
  EmitAB em{A, B, n};
  vtac_block(tb, smem_raw, r, the, phi, k, regular != 0, em);
}

// ---------------------------------------------------------------------------------------------
// K3: y_j = [[A^T B^T],[B^T A^T]](R_j - 0, regular) x,  optionally times diag(scale_j)
//   (Excitation::getIncLocal, srcAna/Excitation.cpp:79-129; source_vector, PreconditionedMatrix.cpp:1339-1341)
// x: 2n vector shared by all particles (x_stride = 0) or per particle (x_stride = 2n).
// grid = local particles; out[(j - j0) * 2n ...].
// ---------------------------------------------------------------------------------------------
struct EmitApply {
  const cplx *x; // 2n, in shared memory
  int n;
  cplx yTE, yTM;
  __device__ __forceinline__ void item(int p, int r, cplx A, cplx B) {
    cplx xa = x[p], xb = x[p + n];
    cfma(yTE, A, xa);
    cfma(yTE, B, xb);
    cfma(yTM, B, xa);
    cfma(yTM, A, xb);
  }
};
__global__ void __launch_bounds__(OB_VTAC_THREADS)
k_translate_apply(VtacTables tb, const double *__restrict__ xyz, cplx k, int j0, const cplx *__restrict__ x,
                  int x_stride, const cplx *__restrict__ scale, cplx *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ cplx xs[2 * OB_MAX_FLAT];
  const int j = j0 + blockIdx.x;
  const int n = flat_max(tb.NM);
  const cplx *xj = x + (size_t)j * x_stride;
  for(int e = threadIdx.x; e < 2 * n; e += blockDim.x)
    xs[e] = xj[e];
  __syncthreads();
  double r, the, phi;
  to_spherical(xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2], r, the, phi);
  cplx *o = out + (size_t)blockIdx.x * 2 * n;
  const cplx *sc = scale ? scale + (size_t)j * 2 * n : nullptr;
  if(fabs(r) < 1e-10) { // Coupling.cpp:82-84: identity
    for(int e = threadIdx.x; e < 2 * n; e += blockDim.x)
      o[e] = sc ? cmul(sc[e], xs[e]) : xs[e];
    return;
  }
  EmitApply em;
  em.x = xs;
  em.n = n;
  em.yTE = em.yTM = mk(0, 0);
  vtac_block(tb, smem_raw, r, the, phi, k, true, em);
  // reduce the column groups (vtac_block ended with a barrier: its buffers are free)
  const int gs = (n + 31) & ~31;
  const int ngroups = blockDim.x / gs;
  const int row = threadIdx.x % gs, grp = threadIdx.x / gs;
  cplx *red = (cplx *)smem_raw; // [ngroups][2][gs]
  if(grp < ngroups) {
    red[(grp * 2 + 0) * gs + row] = em.yTE;
    red[(grp * 2 + 1) * gs + row] = em.yTM;
  }
  __syncthreads();
  for(int e = threadIdx.x; e < 2 * n; e += blockDim.x) {
    int half = e / n, rr = e - half * n;
    cplx s = mk(0, 0);
    for(int g = 0; g < ngroups; ++g)
      s = cadd(s, red[(g * 2 + half) * gs + rr]);
    o[e] = sc ? cmul(sc[e], s) : s;
  }
}

// ---------------------------------------------------------------------------------------------
// K6a: per-particle scattering sum  sum_{p,q} |T_AB[p][q]|^2 |x_q|^2 with the regular block
//   (Result::getScatteringCrossSection[_SH], srcAna/Result.cpp:605-646, 709-760)
//   = sum_{col c,row r} (|A|^2+|B|^2) (|x_c|^2 + |x_{c+n}|^2)
// ---------------------------------------------------------------------------------------------
struct EmitSca {
  const double *w; // |x_c|^2 + |x_{c+n}|^2, n entries in shared memory
  double acc;
  __device__ __forceinline__ void item(int p, int r, cplx A, cplx B) { acc = fma(cnorm(A) + cnorm(B), w[p], acc); }
};
__global__ void __launch_bounds__(OB_VTAC_THREADS)
k_sca_sum(VtacTables tb, const double *__restrict__ xyz, cplx k, int j0, const cplx *__restrict__ x,
          double *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double ws[OB_MAX_FLAT];
  __shared__ double wred[OB_VTAC_THREADS / 32];
  const int j = j0 + blockIdx.x;
  const int n = flat_max(tb.NM);
  const cplx *xj = x + (size_t)j * 2 * n;
  for(int e = threadIdx.x; e < n; e += blockDim.x)
    ws[e] = cnorm(xj[e]) + cnorm(xj[e + n]);
  __syncthreads();
  double r, the, phi;
  to_spherical(xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2], r, the, phi);
  double acc = 0;
  if(fabs(r) < 1e-10) { // identity block: sum_q |x_q|^2
    for(int e = threadIdx.x; e < n; e += blockDim.x)
      acc += ws[e];
  } else {
    EmitSca em;
    em.w = ws;
    em.acc = 0;
    vtac_block(tb, smem_raw, r, the, phi, k, true, em);
    acc = em.acc;
  }
  // deterministic block reduction
  for(int o = 16; o > 0; o >>= 1)
    acc += __shfl_down_sync(0xffffffffu, acc, o);
  if((threadIdx.x & 31) == 0)
    wred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if(threadIdx.x == 0) {
    double s = 0;
    for(int w = 0; w < (int)(blockDim.x >> 5); ++w)
      s += wred[w];
    out[blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static void set_smem(const void *fn, size_t bytes) {
  OB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  // several CTAs per SM overlap one block's serial seed phase with another's level march
  OB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
}

void launch_assemble(VtacTableSet const &ts, const double *xyz, const cplx *Tdiag, cplx k, int nobj, int row0,
                     int nrows, cplx *S, size_t ld, cudaStream_t st) {
  if(nrows <= 0)
    return;
  set_smem((const void *)k_assemble, ts.smem);
  dim3 grid(nobj, nrows);
  k_assemble<<<grid, OB_VTAC_THREADS, ts.smem, st>>>(ts.tb, xyz, Tdiag, k, row0, S, ld);
  OB_CUDA(cudaGetLastError());
}
static int g_pairs_minb = 0; // 0 = auto, 2 / 3 = forced (tuning option "assemble_minb")
void assemble_pairs_tuning(int minb) { g_pairs_minb = minb; }
void launch_assemble_pairs(VtacTableSet const &ts, const double *xyz, cplx k, const int2 *pair_ij, long npairs,
                           cplx *AB, cudaStream_t st) {
  if(npairs <= 0)
    return;
  const bool three = g_pairs_minb == 3 || (g_pairs_minb == 0 && 3 * (ts.smem + 1024) <= 227 * 1024);
  if(three) {
    set_smem((const void *)k_assemble_pairs<3>, ts.smem);
    k_assemble_pairs<3><<<(unsigned)npairs, OB_VTAC_THREADS, ts.smem, st>>>(ts.tb, xyz, k, pair_ij, AB);
  } else {
    set_smem((const void *)k_assemble_pairs<2>, ts.smem);
    k_assemble_pairs<2><<<(unsigned)npairs, OB_VTAC_THREADS, ts.smem, st>>>(ts.tb, xyz, k, pair_ij, AB);
  }
  OB_CUDA(cudaGetLastError());
}
void launch_vtac_single(VtacTableSet const &ts, double r, double the, double phi, cplx k, int regular, cplx *A, cplx *B,
                        cudaStream_t st) {
  set_smem((const void *)k_vtac_single, ts.smem);
  k_vtac_single<<<1, OB_VTAC_THREADS, ts.smem, st>>>(ts.tb, r, the, phi, k, regular, A, B);
  OB_CUDA(cudaGetLastError());
}
void launch_translate_apply(VtacTableSet const &ts, const double *xyz, cplx k, int j0, int count, const cplx *x,
                            int x_stride, const cplx *scale, cplx *out, cudaStream_t st) {
  if(count <= 0)
    return;
  set_smem((const void *)k_translate_apply, ts.smem);
  k_translate_apply<<<count, OB_VTAC_THREADS, ts.smem, st>>>(ts.tb, xyz, k, j0, x, x_stride, scale, out);
  OB_CUDA(cudaGetLastError());
}
void launch_sca_sum(VtacTableSet const &ts, const double *xyz, cplx k, int j0, int count, const cplx *x, double *out,
                    cudaStream_t st) {
  if(count <= 0)
    return;
  set_smem((const void *)k_sca_sum, ts.smem);
  k_sca_sum<<<count, OB_VTAC_THREADS, ts.smem, st>>>(ts.tb, xyz, k, j0, x, out);
  OB_CUDA(cudaGetLastError());
}

} // namespace ob
