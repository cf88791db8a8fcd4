// ob_rot.cu -- rotated-axial form of the preconditioned coupling operator (operator = 3), record layout v2.
//
//   reference operator: S(block i,j) = -T_i [[A^T, B^T],[B^T, A^T]],  A,B = Coupling(R_i - R_j, k, nMax), identity
//   on the diagonal (srcAna/PreconditionedMatrix.cpp:350-400, 555-610); applied by pzgemm_ / matvec
//   (srcAna/scalapack/Belos.hpp:74-90, srcAna/PreconditionedMatrix.cpp:1058-1085).
//
// B200-first representation (not the reference's).  With R = R_i - R_j = (d, theta, phi),
//     [A^T B^T; B^T A^T](R) = P^* D Ax D^T P,   P = diag(exp(i m phi)),  D = blockdiag_n d^n(theta)
// (real Wigner small-d matrices, Varshalovich 4.3.1) and Ax the axial (theta = 0) operator, diagonal in the azimuthal
// order mu with A(-mu) = A(mu), B(-mu) = -B(mu); the reversed direction follows from the parity
// A(-R) = (-1)^(n+l) A(R), B(-R) = (-1)^(n+l+1) B(R).  Two further symmetries halve bytes and flops (tests/rot2_model.py
// is the line-by-line CPU model of this file, held to the oracle's full blocks at 1e-14 by tests/test_oracle_kats.py):
//   * d^n commutes with the flip F e_m = (-1)^m e_-m: in the basis s_0 = e_0, s_a = (e_a + (-1)^a e_-a)/sqrt2,
//     a_a = (e_a - (-1)^a e_-a)/sqrt2 it splits into Ds ((n+1) x (n+1)) and Da (n x n), each with
//     D[a',a] = (-1)^(a'-a) D[a,a'] (one stored array serves D^T and D);
//   * in that basis A keeps the class (s/a) and B swaps it, so the channels (TE_s +- TM_a), (TM_s +- TE_a) diagonalise
//     [A B; B A]: q+ = (A^T + B^T) p+, q- = (A^T - B^T) p-.
// Record per unordered pair (i < j), 16-byte aligned sections:
//   ph[m + NM] = exp(i m phi)                                                      (2 NM + 1 complex)
//   Cp[offX(a) + (n - n0) w + (l - n0)] = A[(n,a),(l,a)] + B[(n,a),(l,a)],  a = 0..NM, n0 = max(a,1), w = NM - n0 + 1
//   Cm[same - NM^2]                     = A - B,                             a = 1..NM   (a = 0: B = 0)
//   Ds[offDs(n) + a' (n + 1) + a],  Da[offDa(n) + (a' - 1) n + (a - 1)]                 (reals)
// 21.4 KB at nMax 10 (v1 30 KB, pair form 460.8 KB), 11.7 KB at nMax 8.
//
// Work decomposition: rows are grouped in blocks of I; a strip (b, j) holds the pairs (i, j), i in block b, i < j.
// Records are stored strip by strip (block-major); ranks and CTAs own contiguous strip ranges of equal pair counts.
// Inside a CTA the row-side sums of the I rows of the current block live in shared memory and the column-side sums
// of the current strip in registers: one column partial per STRIP (not per pair) and one row partial per block
// segment reach HBM, added in a fixed order by k_rot_reduce (deterministic, no atomics).
//
// Apply, per pair and for both directions at once (four vectors: x_j TE/TM, parity-signed x_i TE/TM), CTA of
// ceil32(n) threads, item = (degree, class, a) with the (s_a, a_a) items of one (n, a) in adjacent lanes:
//   P0  t = exp(i m phi) x -> flip basis                 (x prefetched into registers during the previous pair)
//   P1  u = D^T t (thread owns one output, reads its own column of D, broadcast reads of t), channel sums by shuffle
//   P2  q = C p per (a, n) for the eight channels, back to the class vectors
//   P3  w = D v (same array, same access pattern as P1);  P4  flip basis -> m (shuffle), conjugate phase, accumulate
// The record of the next pair is fetched into the other shared-memory slot by one cp.async.bulk (TMA bulk copy,
// mbarrier complete_tx) issued right after P0.
#include "ob_internal.h"
#include "ob_vtac.cuh"
#include "ob_rot_axial.cuh"
#include <algorithm>

namespace ob {

#define ROT_SQH 0.70710678118654752440 // 1 / sqrt(2)

// ---------------------------------------------------------------------------------------------
// layout helpers (host + device), mirrored by tests/rot2_model.py
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int rot_offDs(int n) { return n * (n + 1) * (2 * n + 1) / 6 - 1; } // sum_{j<n} (j+1)^2
__host__ __device__ inline int rot_offDa(int n) { return (n - 1) * n * (2 * n - 1) / 6; }     // sum_{j<n} j^2
__host__ __device__ inline int rot_offF(int n) { return (n - 1) * (n + 2); }
__host__ __device__ inline int rot_offP(int NM, int a) { // channel index of (a, l = n0): sum_{u<a} (NM - max(u,1) + 1)
  return a <= 0 ? 0 : NM + (a - 1) * (NM + 1) - (a - 1) * a / 2;
}
RotLayout rot_layout(int NM) {
  RotLayout L;
  L.NM = NM;
  L.n = flat_max(NM);
  L.X = rot_offX(NM, NM + 1);
  L.nDs = rot_offDs(NM + 1);
  L.nDa = rot_offDa(NM + 1);
  L.LF = NM * (NM + 3);
  L.nh = L.LF / 2;
  L.offCp = (size_t)(2 * NM + 1) * sizeof(cplx);
  L.offCm = L.offCp + (size_t)L.X * sizeof(cplx);
  L.offDs = L.offCm + (size_t)(L.X - NM * NM) * sizeof(cplx);
  L.offDa = (L.offDs + (size_t)L.nDs * sizeof(double) + 15) & ~(size_t)15;
  L.rec_bytes = (L.offDa + (size_t)L.nDa * sizeof(double) + 15) & ~(size_t)15;
  return L;
}

// ---------------------------------------------------------------------------------------------
// assembly 1 (cross-check path, "rot_assembly" = 0): axial A, B out of the shared VTAC block code at theta = phi = 0
// ---------------------------------------------------------------------------------------------
struct EmitAxial {
  cplx *Cp, *Cm;
  int NM;
  // p = flat(n, mu) (first index of Coupling.diagonal), r = flat(l, k): keep mu == k >= 0
  __device__ __forceinline__ void item(int p, int r, cplx a, cplx b) {
    int n, mu, l, k;
    unflatten(p, n, mu);
    unflatten(r, l, k);
    if(mu != k || mu < 0)
      return;
    const int n0 = rot_n0(mu), w = NM - n0 + 1;
    const int e = rot_offX(NM, mu) + (n - n0) * w + (l - n0);
    Cp[e] = cadd(a, b);
    if(mu >= 1)
      Cm[e - NM * NM] = csub(a, b);
  }
};
__global__ void __launch_bounds__(OB_VTAC_THREADS, 2)
k_assemble_axial(VtacTables tb, const double *__restrict__ xyz, cplx k, const int2 *__restrict__ pair_ij,
                 unsigned char *__restrict__ recs, RotLayout L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int2 ij = pair_ij[blockIdx.x];
  const double x = xyz[3 * ij.x] - xyz[3 * ij.y], y = xyz[3 * ij.x + 1] - xyz[3 * ij.y + 1],
               z = xyz[3 * ij.x + 2] - xyz[3 * ij.y + 2];
  const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
  unsigned char *rec = recs + (size_t)blockIdx.x * L.rec_bytes;
  EmitAxial em;
  em.Cp = (cplx *)(rec + L.offCp);
  em.Cm = (cplx *)(rec + L.offCm);
  em.NM = L.NM;
  vtac_block(tb, smem_raw, r, 0.0, 0.0, k, false, em);
}

// ---------------------------------------------------------------------------------------------
// assembly 1b (default, "rot_assembly" = 1): axial-only recursion, one warp per pair.  With theta = 0 the scalar
// coefficients beta(n, m, l, k) vanish unless k = m and the reference's recursion
// (TranslationAdditionCoefficients.cpp:102-124) closes on those entries: O(nMax^3) per pair instead of the O(nMax^4)
// of the full block.  The per-pair body lives in ob_rot_axial.cuh and is ALSO compiled for the host: the very source the
// warp runs is checked on the CPU against the oracle's Coupling (tests/test_rot_axial_host.py, lane 0 of 1) and on the
// GPU against the vtac_block path and the oracle (tests/test_gpu_rot.py).
// ---------------------------------------------------------------------------------------------
#define ROT_AX_WARPS 4
static size_t rot_axial_smem_bytes(int NM) { return (size_t)ROT_AX_WARPS * rot_axial_buf_entries(NM) * sizeof(cplx); }
__global__ void __launch_bounds__(ROT_AX_WARPS * 32)
k_assemble_axial_only(const double *__restrict__ xyz, cplx k, const int2 *__restrict__ pair_ij, long npairs,
                      unsigned char *__restrict__ recs, RotLayout L) {
  extern __shared__ __align__(16) unsigned char smem_ax[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int entries = rot_axial_buf_entries(L.NM);
  cplx *buf = (cplx *)smem_ax + (size_t)warp * entries;
  for(int e = lane; e < entries; e += 32)
    buf[e] = mk(0, 0);
  __syncwarp();
  for(long q = (long)blockIdx.x * ROT_AX_WARPS + warp; q < npairs; q += (long)gridDim.x * ROT_AX_WARPS) {
    const int2 ij = pair_ij[q];
    const double x = xyz[3 * ij.x] - xyz[3 * ij.y], y = xyz[3 * ij.x + 1] - xyz[3 * ij.y + 1],
                 z = xyz[3 * ij.x + 2] - xyz[3 * ij.y + 2];
    const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    unsigned char *rec = recs + (size_t)q * L.rec_bytes;
    rot_axial_pair(L.NM, k, r, buf, (cplx *)(rec + L.offCp), (cplx *)(rec + L.offCm), lane, 32, 1);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// assembly 2: phases and flip-basis Wigner small-d matrices of every local pair.  One thread per (a', a), 0 <= a', a
// <= NM, runs the three-term recurrences in the degree j of d^j_{a' a} and d^j_{a' -a} together (they share the seed
// degree j0 = max(a', a) and the recurrence coefficients up to one sign), written as
//     d_j = (alpha_j c -+ beta_j) d_{j-1} - gamma_j d_{j-2},   c = cos(theta);
// alpha, beta, gamma and the seed constants depend on the indices only and come from a table built once per nMax.
// ---------------------------------------------------------------------------------------------
struct RotDTable {
  int NM = -1, device = -1;
  double *coef = nullptr; // [(j * W + a') * W + a][3], W = NM + 1
  double *seed = nullptr; // [(a' * W + a)][2]: |d^{j0}_{a' a}| and |d^{j0}_{a' -a}| at cos = sin = 1 (signs in the kernel)
};
static RotDTable g_dtab[16][OB_MAX_NMAX + 1];

static RotDTable const &rot_dtable(int NM) {
  int dev = 0;
  OB_CUDA(cudaGetDevice(&dev));
  if(dev < 0 || dev >= 16)
    throw Error("rotated-axial operator: device ordinal out of range");
  RotDTable &t = g_dtab[dev][NM];
  if(t.NM == NM)
    return t;
  const int W = NM + 1;
  std::vector<double> fact(2 * NM + 2, 1.0);
  for(size_t i = 1; i < fact.size(); ++i)
    fact[i] = fact[i - 1] * (double)i;
  std::vector<double> coef((size_t)(NM + 1) * W * W * 3, 0.0), seed((size_t)W * W * 2, 0.0);
  for(int j = 1; j <= NM; ++j)
    for(int mp = 0; mp <= NM; ++mp)
      for(int m = 0; m <= NM; ++m) {
        double *c = &coef[((size_t)(j * W + mp) * W + m) * 3];
        if(mp > j - 1 || m > j - 1)
          continue; // the recurrence into degree j starts from j0 = max(mp, m) <= j - 1
        if(mp == 0 && m == 0) {
          c[0] = (2.0 * j - 1.0) / j;
          c[1] = 0.0;
          c[2] = (j - 1.0) / j;
        } else {
          const double den = (j - 1) * std::sqrt((double)((j * j - mp * mp) * (j * j - m * m)));
          c[0] = (2.0 * j - 1.0) * (double)(j * (j - 1)) / den;
          c[1] = (2.0 * j - 1.0) * (double)(m * mp) / den;
          c[2] = j * std::sqrt((double)(((j - 1) * (j - 1) - mp * mp) * ((j - 1) * (j - 1) - m * m))) / den;
        }
      }
  for(int mp = 0; mp <= NM; ++mp)
    for(int m = 0; m <= NM; ++m) {
      const int j0 = std::max(mp, m);
      for(int s = 0; s < 2; ++s) { // s = 0: d^{j0}_{mp, m};  s = 1: d^{j0}_{mp, -m}
        const int mm = s ? -m : m;
        const int tt = mm - mp > 0 ? mm - mp : 0; // the explicit sum (Varshalovich 4.3.1 (2)) has this single term
        seed[((size_t)mp * W + m) * 2 + s] =
            std::sqrt(fact[j0 + mp] * fact[j0 - mp] * fact[j0 + mm] * fact[j0 - mm]) /
            (fact[j0 + mm - tt] * fact[tt] * fact[mp - mm + tt] * fact[j0 - mp - tt]);
      }
    }
  OB_CUDA(cudaMalloc(&t.coef, coef.size() * sizeof(double)));
  OB_CUDA(cudaMemcpy(t.coef, coef.data(), coef.size() * sizeof(double), cudaMemcpyHostToDevice));
  OB_CUDA(cudaMalloc(&t.seed, seed.size() * sizeof(double)));
  OB_CUDA(cudaMemcpy(t.seed, seed.data(), seed.size() * sizeof(double), cudaMemcpyHostToDevice));
  t.NM = NM;
  t.device = dev;
  return t;
}

#define ROT_TAB_THREADS 128
__global__ void __launch_bounds__(ROT_TAB_THREADS)
k_rot_tables(const double *__restrict__ xyz, const int2 *__restrict__ pair_ij, long npairs,
             unsigned char *__restrict__ recs, RotLayout L, const double *__restrict__ coef,
             const double *__restrict__ seedc) {
  __shared__ double s_cp[2 * OB_MAX_NMAX + 1], s_sp[2 * OB_MAX_NMAX + 1]; // powers of cos, sin(theta / 2)
  const int NM = L.NM, W = NM + 1;
  for(long q = blockIdx.x; q < npairs; q += gridDim.x) {
    const int2 ij = pair_ij[q];
    const double x = xyz[3 * ij.x] - xyz[3 * ij.y], y = xyz[3 * ij.x + 1] - xyz[3 * ij.y + 1],
                 z = xyz[3 * ij.x + 2] - xyz[3 * ij.y + 2];
    const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    double beta = 0.0, phi = 0.0;
    if(r > 0.0) {
      beta = acos(z / r);
      phi = atan2(y, x);
    }
    unsigned char *rec = recs + (size_t)q * L.rec_bytes;
    cplx *ph = (cplx *)rec;
    double *Ds = (double *)(rec + L.offDs), *Da = (double *)(rec + L.offDa);
    double sb, cb;
    sincos(0.5 * beta, &sb, &cb);
    for(int t = threadIdx.x; t < 2 * NM + 1; t += blockDim.x) {
      double s, c;
      sincos((double)(t - NM) * phi, &s, &c);
      ph[t] = mk(c, s); // ph[m + NM] = exp(i m phi)
      s_cp[t] = pow(cb, (double)t);
      s_sp[t] = pow(sb, (double)t);
    }
    __syncthreads();
    const double c = cos(beta);
    for(int t = threadIdx.x; t < W * W; t += blockDim.x) {
      const int mp = t / W, m = t - mp * W;
      const int j0 = mp > m ? mp : m;
      // seeds at j = j0 (one term of the explicit sum): exponents 2 j0 + mm - mp - 2 tt and mp - mm + 2 tt, sign
      // (-1)^(mp - mm + tt), tt = max(0, mm - mp), for mm = +m and mm = -m
      const int tp = m - mp > 0 ? m - mp : 0;
      double dP = seedc[2 * t] * s_cp[2 * j0 + m - mp - 2 * tp] * s_sp[mp - m + 2 * tp]; // d^{j0}_{mp, m}
      if((mp - m + tp) & 1)
        dP = -dP;
      double dM = seedc[2 * t + 1] * s_cp[2 * j0 - m - mp] * s_sp[mp + m]; // d^{j0}_{mp, -m}: tt = 0
      if((mp + m) & 1)
        dM = -dM;
      double pP = 0.0, pM = 0.0;
      const double sa = (m & 1) ? -1.0 : 1.0;
      for(int j = j0; j <= NM; ++j) {
        if(j > j0) {
          const double *cf = coef + ((size_t)(j * W + mp) * W + m) * 3;
          const double al = cf[0] * c, be = cf[1], ga = cf[2];
          const double nP = (al - be) * dP - ga * pP, nM = (al + be) * dM - ga * pM;
          pP = dP;
          pM = dM;
          dP = nP;
          dM = nM;
        }
        if(j < 1)
          continue;
        // flip basis (tests/rot2_model.py build_pair): rows a' = mp, columns a = m
        double vs;
        if(mp == 0 && m == 0)
          vs = dP;
        else if(m == 0 || mp == 0)
          vs = 1.41421356237309504880 * dP;
        else {
          vs = dP + sa * dM;
          Da[rot_offDa(j) + (mp - 1) * j + (m - 1)] = dP - sa * dM;
        }
        Ds[rot_offDs(j) + mp * (j + 1) + m] = vs;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// apply
// ---------------------------------------------------------------------------------------------
struct RotArgs {
  const unsigned char *recs;
  const cplx *x;
  const int4 *pinfo;   // per local pair: (i, j, local strip, flags: 1 = last pair of its strip, 2 = last pair of its segment)
  const int *cta_pair; // [grid + 1] local pair ranges
  const int *cta_seg;  // [grid + 1] segment ranges
  cplx *rowpart, *colpart;
  RotLayout L;
  int I;
};

__device__ __forceinline__ uint32_t r_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void r_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(r_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void r_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(r_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void r_mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "RWAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra RWAIT_DONE;\n"
               "bra RWAIT_LOOP;\n"
               "RWAIT_DONE:\n"
               "}" ::"r"(r_smem_u32(bar)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void r_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(r_smem_u32(dst)), "l"(src), "r"(bytes), "r"(r_smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ cplx shfl_xor1(cplx v) {
  return mk(__shfl_xor_sync(0xffffffffu, v.x, 1), __shfl_xor_sync(0xffffffffu, v.y, 1));
}

// dynamic shared memory: record[2] | bufA[4 LF] | bufB[4 LF] | rowacc[I][2n] | mbarrier[2]
static size_t rot_smem_bytes(RotLayout const &L, int I) {
  return 2 * L.rec_bytes + (size_t)(8 * L.LF + I * 2 * L.n) * sizeof(cplx) + 2 * sizeof(uint64_t);
}
static int rot_block_threads(RotLayout const &L) { return (L.n + 31) & ~31; }

// d-phase inner product (P1 and P3): acc[vv] = sum_t d[t stride] v[4 t + vv], vv = direction * 2 + polarisation
__device__ __forceinline__ void rot_dphase(const double *__restrict__ dp, int stride, const cplx *__restrict__ vp, int cnt,
                                           cplx acc[4]) {
  acc[0] = acc[1] = acc[2] = acc[3] = mk(0, 0);
#pragma unroll 2
  for(int t = 0; t < cnt; ++t) {
    const double d = dp[t * stride];
    const cplx v0 = vp[4 * t], v1 = vp[4 * t + 1], v2 = vp[4 * t + 2], v3 = vp[4 * t + 3];
    acc[0].x = fma(d, v0.x, acc[0].x);
    acc[0].y = fma(d, v0.y, acc[0].y);
    acc[1].x = fma(d, v1.x, acc[1].x);
    acc[1].y = fma(d, v1.y, acc[1].y);
    acc[2].x = fma(d, v2.x, acc[2].x);
    acc[2].y = fma(d, v2.y, acc[2].y);
    acc[3].x = fma(d, v3.x, acc[3].x);
    acc[3].y = fma(d, v3.y, acc[3].y);
  }
}

__global__ void __launch_bounds__(OB_ROT_MAX_THREADS) k_matvec_rot(RotArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const RotLayout L = a.L;
  const int NM = L.NM, nH = L.n, n2 = 2 * nH, I = a.I;
  cplx *bufA = (cplx *)(smem + 2 * L.rec_bytes);
  cplx *bufB = bufA + 4 * L.LF;
  cplx *rowacc = bufB + 4 * L.LF;
  uint64_t *full = (uint64_t *)(rowacc + (size_t)I * n2);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int qbeg = a.cta_pair[blockIdx.x], qend = a.cta_pair[blockIdx.x + 1];
  if(qbeg >= qend)
    return;
  if(tid == 0) {
    r_mbar_init(&full[0], 1);
    r_mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- the thread's item: paired lanes (2k, 2k+1) = (s_a, a_a) of (n, a >= 1); then the s_0 items; then idle lanes ----
  const int npaired = NM * (NM + 1);
  int dn = 1, da = 0, cls = 0, kind = 2; // kind 0: paired, 1: s_0, 2: idle
  if(tid < npaired) {
    const int k = tid >> 1;
    dn = (int)((1.0 + sqrt(1.0 + 8.0 * k)) * 0.5);
    while(dn * (dn - 1) / 2 > k)
      --dn;
    while(dn * (dn + 1) / 2 <= k)
      ++dn;
    da = k - dn * (dn - 1) / 2 + 1;
    cls = tid & 1;
    kind = 0;
  } else if(tid < nH) {
    dn = tid - npaired + 1;
    kind = 1;
  }
  const int offFn = rot_offF(dn);
  const int cnt = kind == 2 ? 0 : (cls ? dn : dn + 1), stride = cls ? dn : dn + 1;
  const size_t dofs = cls ? L.offDa + (size_t)(rot_offDa(dn) + (da - 1)) * sizeof(double)
                          : L.offDs + (size_t)(rot_offDs(dn) + da) * sizeof(double);
  const int vbase = cls ? offFn + dn + 2 : offFn; // first F index the d-loops read
  const int fpos = flat_index(dn, da), fneg = flat_index(dn, -da);
  const double sa = (da & 1) ? -1.0 : 1.0, sn = (dn & 1) ? -1.0 : 1.0;
  const int pidx = rot_offP(NM, da) + (dn - rot_n0(da)); // channel index of (a, l = dn)
  // ---- the thread's P2 item (a2, d2), sorted by a then n ----
  int a2 = 0, d2 = 1;
  const bool p2live = tid < L.nh;
  if(p2live) {
    int rem = tid;
    for(a2 = 0; a2 <= NM; ++a2) {
      const int w = NM - rot_n0(a2) + 1;
      if(rem < w)
        break;
      rem -= w;
    }
    d2 = rot_n0(a2) + rem;
  }
  const int n02 = rot_n0(a2), w2 = NM - n02 + 1;
  const int cofs = rot_offX(NM, a2) + (d2 - n02); // + (l - n0) w2 per step
  const int pbase2 = rot_offP(NM, a2);
  const double sa2 = (a2 & 1) ? -0.5 : 0.5;
  const int fs2 = rot_offF(d2) + a2, fa2 = rot_offF(d2) + d2 + 1 + a2;

  for(int e = tid; e < I * n2; e += nthr)
    rowacc[e] = mk(0, 0);
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  __syncthreads();
  if(tid == 0) {
    r_mbar_expect_tx(&full[0], (uint32_t)L.rec_bytes);
    r_bulk_g2s(smem, a.recs + (size_t)qbeg * L.rec_bytes, (uint32_t)L.rec_bytes, &full[0], pol);
  }
  // x of the pair about to be processed: paired s-lanes hold x_j (direction 0), a-lanes x_i (direction 1), at
  // m = +a (xr[0..1] = TE, TM) and m = -a (xr[2..3]); s_0 lanes hold x_j(0) TE/TM and x_i(0) TE/TM
  cplx xr[4];
  xr[0] = xr[1] = xr[2] = xr[3] = mk(0, 0);
  int4 pi = a.pinfo[qbeg];
  auto load_x = [&](int4 const &p, bool reload_j) {
    if(kind == 0) {
      if(cls == 1 || reload_j) {
        const cplx *xs = a.x + (size_t)(cls ? p.x : p.y) * n2;
        xr[0] = xs[fpos];
        xr[1] = xs[nH + fpos];
        xr[2] = xs[fneg];
        xr[3] = xs[nH + fneg];
      }
    } else if(kind == 1) {
      const cplx *xj = a.x + (size_t)p.y * n2, *xi = a.x + (size_t)p.x * n2;
      if(reload_j) {
        xr[0] = xj[fpos];
        xr[1] = xj[nH + fpos];
      }
      xr[2] = xi[fpos];
      xr[3] = xi[nH + fpos];
    }
  };
  load_x(pi, true);
  cplx colacc[4];
  colacc[0] = colacc[1] = colacc[2] = colacc[3] = mk(0, 0);
  int sg = a.cta_seg[blockIdx.x];
  for(int q = qbeg; q < qend; ++q) {
    const int cur = (q - qbeg) & 1;
    const unsigned char *rec = smem + (size_t)cur * L.rec_bytes;
    const cplx *s_ph = (const cplx *)rec;
    const cplx *s_Cp = (const cplx *)(rec + L.offCp);
    const cplx *s_Cm = (const cplx *)(rec + L.offCm) - NM * NM;
    const double *dp = (const double *)(rec + dofs);
    r_mbar_wait(&full[cur], (uint32_t)(((q - qbeg) >> 1) & 1));
    // ---- P0: phases, parity signs of the reversed direction, flip basis ----
    if(kind == 0) {
      const cplx pp = s_ph[NM + da], pm = s_ph[NM - da];
      const cplx te_p = cmul(pp, xr[0]), tm_p = cmul(pp, xr[1]), te_m = cmul(pm, xr[2]), tm_m = cmul(pm, xr[3]);
      // direction 1 (a-lanes): x_i with (-1)^deg, and -1 on TM
      const double fe = (cls ? sn : 1.0) * ROT_SQH, fm = (cls ? -sn : 1.0) * ROT_SQH;
      cplx *ts = bufA + 4 * (offFn + da) + 2 * cls, *ta = bufA + 4 * (offFn + dn + 1 + da) + 2 * cls;
      ts[0] = mk(fe * (te_p.x + sa * te_m.x), fe * (te_p.y + sa * te_m.y));
      ts[1] = mk(fm * (tm_p.x + sa * tm_m.x), fm * (tm_p.y + sa * tm_m.y));
      ta[0] = mk(fe * (te_p.x - sa * te_m.x), fe * (te_p.y - sa * te_m.y));
      ta[1] = mk(fm * (tm_p.x - sa * tm_m.x), fm * (tm_p.y - sa * tm_m.y));
    } else if(kind == 1) {
      cplx *ts = bufA + 4 * offFn; // exp(i 0 phi) = 1
      ts[0] = xr[0];
      ts[1] = xr[1];
      ts[2] = cscale(xr[2], sn);
      ts[3] = cscale(xr[3], -sn);
    }
    __syncthreads(); // B1: T complete; every thread is past P3/P4 of the previous pair -> the other record slot is free
    int4 pnext = pi;
    if(q + 1 < qend) {
      if(tid == 0) {
        r_mbar_expect_tx(&full[cur ^ 1], (uint32_t)L.rec_bytes);
        r_bulk_g2s(smem + (size_t)(cur ^ 1) * L.rec_bytes, a.recs + (size_t)(q + 1) * L.rec_bytes, (uint32_t)L.rec_bytes,
                   &full[cur ^ 1], pol);
      }
      pnext = a.pinfo[q + 1];
    }
    // ---- P1: u = D^T t, channel combinations ----
    cplx acc[4];
    rot_dphase(dp, stride, bufA + 4 * vbase, cnt, acc);
    if(q + 1 < qend)
      load_x(pnext, pnext.y != pi.y); // next pair's x into registers (L2 hits), consumed by its P0
    {
      // s-lane: p+- = TE_s +- TM_a; a-lane: r+- = TM_s +- TE_a (the partner's TM of both directions)
      const cplx o1 = shfl_xor1(acc[1]), o3 = shfl_xor1(acc[3]);
      if(kind == 0) {
        cplx *P = bufB + 8 * pidx + 2 * cls;
        if(cls == 0) {
          P[0] = cadd(acc[0], o1);
          P[1] = csub(acc[0], o1);
          P[4] = cadd(acc[2], o3);
          P[5] = csub(acc[2], o3);
        } else {
          P[0] = cadd(o1, acc[0]);
          P[1] = csub(o1, acc[0]);
          P[4] = cadd(o3, acc[2]);
          P[5] = csub(o3, acc[2]);
        }
      } else if(kind == 1) {
        cplx *P = bufB + 8 * pidx;
        P[0] = acc[0];
        P[1] = acc[0];
        P[2] = acc[1];
        P[3] = acc[1];
        P[4] = acc[2];
        P[5] = acc[2];
        P[6] = acc[3];
        P[7] = acc[3];
      }
    }
    __syncthreads(); // B2
    // ---- P2: q = C p for the eight channels of (a2, d2), back to the class vectors ----
    if(p2live) {
      cplx qv[8];
#pragma unroll
      for(int c = 0; c < 8; ++c)
        qv[c] = mk(0, 0);
      const cplx *cp = s_Cp + cofs, *cm = (a2 == 0 ? s_Cp : s_Cm) + cofs;
      const cplx *pv = bufB + 8 * pbase2;
      for(int l = 0; l < w2; ++l) {
        const cplx vp = cp[l * w2], vm = cm[l * w2];
#pragma unroll
        for(int c = 0; c < 8; c += 2) {
          cfma(qv[c], vp, pv[8 * l + c]);
          cfma(qv[c + 1], vm, pv[8 * l + c + 1]);
        }
      }
#pragma unroll
      for(int d = 0; d < 2; ++d) {
        const cplx qp = qv[4 * d], qm = qv[4 * d + 1], rp = qv[4 * d + 2], rm = qv[4 * d + 3];
        bufA[4 * fs2 + 2 * d] = mk(sa2 * (qp.x + qm.x), sa2 * (qp.y + qm.y));     // TE_s
        bufA[4 * fs2 + 2 * d + 1] = mk(sa2 * (rp.x + rm.x), sa2 * (rp.y + rm.y)); // TM_s
        if(a2 >= 1) {
          bufA[4 * fa2 + 2 * d] = mk(sa2 * (rp.x - rm.x), sa2 * (rp.y - rm.y));     // TE_a
          bufA[4 * fa2 + 2 * d + 1] = mk(sa2 * (qp.x - qm.x), sa2 * (qp.y - qm.y)); // TM_a
        }
      }
    }
    __syncthreads(); // B3
    // ---- P3: w = D v;  P4: flip basis -> m, conjugate phase, parity signs, accumulate ----
    rot_dphase(dp, stride, bufA + 4 * vbase, cnt, acc);
    {
      // the s-lane finishes direction 0 (needs the partner's w_a of direction 0), the a-lane direction 1
      const cplx r0 = shfl_xor1(cls ? acc[0] : acc[2]), r1 = shfl_xor1(cls ? acc[1] : acc[3]);
      cplx *ra = rowacc + (size_t)(pi.x % I) * n2;
      if(kind == 0) {
        const cplx cpp = cconj(s_ph[NM + da]), cpm = cconj(s_ph[NM - da]);
        if(cls == 0) { // w_s = acc[0..1], w_a = r0, r1; the (-1)^a' of the transposed read: sa on the +a' output
          const double f = sa * ROT_SQH;
          const cplx ep = cmul(cpp, mk(f * (acc[0].x + r0.x), f * (acc[0].y + r0.y)));
          const cplx mp_ = cmul(cpp, mk(f * (acc[1].x + r1.x), f * (acc[1].y + r1.y)));
          const cplx em = cmul(cpm, mk(ROT_SQH * (acc[0].x - r0.x), ROT_SQH * (acc[0].y - r0.y)));
          const cplx mm = cmul(cpm, mk(ROT_SQH * (acc[1].x - r1.x), ROT_SQH * (acc[1].y - r1.y)));
          ra[fpos] = cadd(ra[fpos], ep);
          ra[nH + fpos] = cadd(ra[nH + fpos], mp_);
          ra[fneg] = cadd(ra[fneg], em);
          ra[nH + fneg] = cadd(ra[nH + fneg], mm);
        } else { // w_a = acc[2..3], w_s = r0, r1; direction 1: (-1)^deg, and -1 on TM
          const double fe = sa * sn * ROT_SQH, fm = -fe, ge = sn * ROT_SQH, gm = -ge;
          cfma(colacc[0], cpp, mk(fe * (r0.x + acc[2].x), fe * (r0.y + acc[2].y)));
          cfma(colacc[1], cpp, mk(fm * (r1.x + acc[3].x), fm * (r1.y + acc[3].y)));
          cfma(colacc[2], cpm, mk(ge * (r0.x - acc[2].x), ge * (r0.y - acc[2].y)));
          cfma(colacc[3], cpm, mk(gm * (r1.x - acc[3].x), gm * (r1.y - acc[3].y)));
        }
      } else if(kind == 1) {
        ra[fpos] = cadd(ra[fpos], acc[0]);
        ra[nH + fpos] = cadd(ra[nH + fpos], acc[1]);
        colacc[0] = cadd(colacc[0], cscale(acc[2], sn));
        colacc[1] = cadd(colacc[1], cscale(acc[3], -sn));
      }
    }
    if(pi.w & 1) { // last pair of the strip: the column-side sums of particle j
      cplx *cpart = a.colpart + (size_t)pi.z * n2;
      if(kind == 0 && cls == 1) {
        cpart[fpos] = colacc[0];
        cpart[nH + fpos] = colacc[1];
        cpart[fneg] = colacc[2];
        cpart[nH + fneg] = colacc[3];
      } else if(kind == 1) {
        cpart[fpos] = colacc[0];
        cpart[nH + fpos] = colacc[1];
      }
      colacc[0] = colacc[1] = colacc[2] = colacc[3] = mk(0, 0);
    }
    if(pi.w & 2) { // last pair of the segment: flush the row-side sums of the block's I rows
      __syncthreads();
      cplx *rp = a.rowpart + (size_t)sg * I * n2;
      for(int e = tid; e < I * n2; e += nthr) {
        rp[e] = rowacc[e];
        rowacc[e] = mk(0, 0);
      }
      ++sg;
    }
    pi = pnext;
    cplx *tb = bufA;
    bufA = bufB;
    bufB = tb;
  }
}

// acc_p = row-side partials of the segments of p's block (segment order) + the column-side partials of the local strips
// (b, p), b ascending.  finalize != 0: y_p = x_p - T_p .* acc_p written directly (single rank).
#define ROT_REDUCE_THREADS 1024
__global__ void __launch_bounds__(ROT_REDUCE_THREADS)
k_rot_reduce(const cplx *__restrict__ rowpart, const cplx *__restrict__ colpart, const int *__restrict__ blk_seg,
             const long *__restrict__ blk_strip, long strip0, long nstrips, int I, int n2,
             const cplx *__restrict__ x, const cplx *__restrict__ Tdiag, cplx *__restrict__ out, int finalize) {
  __shared__ cplx sh[ROT_REDUCE_THREADS];
  const int p = blockIdx.x, b = p / I, slot = p - b * I;
  const int parts = (int)blockDim.x / n2; // blockDim.x is a multiple of n2
  const int part = threadIdx.x / n2, e = threadIdx.x - part * n2;
  const int s0 = blk_seg[b], s1 = blk_seg[b + 1];
  const int ncol = (p + I - 1) / I; // blocks bb with I bb < p
  const int nterms = (s1 - s0) + ncol;
  cplx a0 = mk(0, 0), a1 = a0, a2 = a0, a3 = a0;
  auto term = [&](int t) -> cplx {
    if(t < s1 - s0)
      return rowpart[((size_t)(s0 + t) * I + slot) * n2 + e];
    const int bb = t - (s1 - s0);
    const long loc = blk_strip[bb] + (p - I * bb - 1) - strip0;
    if(loc < 0 || loc >= nstrips)
      return mk(0, 0);
    return colpart[(size_t)loc * n2 + e];
  };
  int k = part;
  for(; k + 3 * parts < nterms; k += 4 * parts) {
    const cplx v0 = term(k), v1 = term(k + parts), v2 = term(k + 2 * parts), v3 = term(k + 3 * parts);
    a0 = cadd(a0, v0);
    a1 = cadd(a1, v1);
    a2 = cadd(a2, v2);
    a3 = cadd(a3, v3);
  }
  for(; k < nterms; k += parts)
    a0 = cadd(a0, term(k));
  cplx s = cadd(cadd(a0, a1), cadd(a2, a3));
  sh[threadIdx.x] = s;
  __syncthreads();
  if(part != 0)
    return;
  for(int q = 1; q < parts; ++q)
    s = cadd(s, sh[q * n2 + e]);
  const size_t o = (size_t)p * n2 + e;
  out[o] = finalize ? csub(x[o], cmul(Tdiag[o], s)) : s;
}

// ---------------------------------------------------------------------------------------------
// host: plan
// ---------------------------------------------------------------------------------------------
template <class T> static T *rot_upload(std::vector<T> const &v) {
  T *d = nullptr;
  OB_CUDA(cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(T)));
  if(!v.empty())
    OB_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

void rot_plan_release(RotPlan &p) {
  void *ptrs[] = {p.pair_ij, p.pinfo, p.cta_pair, p.cta_seg, p.blk_seg, p.blk_strip, p.rowpart, p.colpart, p.acc};
  for(void *q : ptrs)
    if(q)
      cudaFree(q);
  p = RotPlan();
}

static int g_rot_rows = 0, g_rot_ctas = 0; // tuning overrides ("rot_rows", "rot_ctas_per_sm"); 0 = auto

void rot_plan_build(RotPlan &p, int nobj, int NM, int world, int rank, int sm_count) {
  rot_plan_release(p);
  const RotLayout L = rot_layout(NM);
  p.nobj = nobj;
  p.n = L.n;
  p.threads = rot_block_threads(L);
  if(p.threads > OB_ROT_MAX_THREADS)
    throw Error("rotated-axial operator: nMax exceeds the kernel's thread bound");
  // rows per block: as many (<= 4) as keep three CTAs per SM resident (the row sums of a block live in shared memory)
  int I = g_rot_rows > 0 ? g_rot_rows : 4;
  const size_t sm_total = (size_t)228 * 1024, cta_max = (size_t)227 * 1024;
  while(I > 1 && 3 * (rot_smem_bytes(L, I) + 1024) > sm_total)
    --I;
  if(rot_smem_bytes(L, I) > cta_max)
    throw Error("rotated-axial operator: record does not fit in shared memory");
  p.I = I;
  p.smem = rot_smem_bytes(L, I);
  p.ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, sm_total / (p.smem + 1024)));
  p.ctas_per_sm = std::max(1, std::min(p.ctas_per_sm, 2048 / p.threads));
  if(g_rot_ctas > 0)
    p.ctas_per_sm = std::min(p.ctas_per_sm, g_rot_ctas);
  p.nblocks = (nobj + I - 1) / I;
  // global strip list, block-major: block b holds the strips j = I b + 1 .. nobj - 1, strip (b, j) the pairs
  // (i, j), I b <= i < min(I b + I, j)
  std::vector<long> blk_strip(p.nblocks + 1, 0);
  for(int b = 0; b < p.nblocks; ++b)
    blk_strip[b + 1] = blk_strip[b] + std::max(0, nobj - 1 - I * b);
  const long S = blk_strip[p.nblocks], P = (long)nobj * (nobj - 1) / 2;
  auto strip_pairs = [&](int b, int j) { return std::min(I, j - I * b); };
  // rank ranges: contiguous strips, cut where the running pair count crosses P r / world
  std::vector<long> cut(world + 1, 0);
  {
    long run = 0, g = 0;
    int r = 1;
    for(int b = 0; b < p.nblocks && r < world; ++b)
      for(int j = I * b + 1; j < nobj && r < world; ++j, ++g) {
        run += strip_pairs(b, j);
        while(r < world && run >= P * r / world)
          cut[r++] = g + 1;
      }
    for(; r < world; ++r)
      cut[r] = S;
    cut[world] = S;
  }
  p.strip0 = cut[rank];
  p.nstrips = cut[rank + 1] - cut[rank];
  // local pairs
  std::vector<int2> ij;
  std::vector<int4> pinfo;
  std::vector<int> strip_blk((size_t)p.nstrips);
  std::vector<long> strip_first((size_t)p.nstrips + 1, 0);
  {
    long g = 0;
    for(int b = 0; b < p.nblocks; ++b)
      for(int j = I * b + 1; j < nobj; ++j, ++g) {
        if(g < p.strip0 || g >= p.strip0 + p.nstrips)
          continue;
        const int loc = (int)(g - p.strip0), cntp = strip_pairs(b, j);
        strip_blk[loc] = b;
        strip_first[loc] = (long)ij.size();
        for(int t = 0; t < cntp; ++t) {
          ij.push_back(make_int2(I * b + t, j));
          pinfo.push_back(make_int4(I * b + t, j, loc, t + 1 == cntp ? 1 : 0));
        }
      }
    strip_first[(size_t)p.nstrips] = (long)ij.size();
  }
  p.npairs = (long)ij.size();
  if(p.npairs >= ((long)1 << 31))
    throw Error("rotated-axial operator: too many local pairs");
  // CTA ranges: contiguous strips of (nearly) equal pair counts; segments = maximal runs inside one block
  p.grid = (int)std::max<long>(1, std::min<long>((long)sm_count * p.ctas_per_sm, p.nstrips));
  std::vector<long> cs(p.grid + 1, 0);
  cs[p.grid] = p.nstrips;
  for(int c = 1; c < p.grid; ++c) {
    long t = std::lower_bound(strip_first.begin(), strip_first.end(), p.npairs * c / p.grid) - strip_first.begin();
    t = std::max(t, cs[c - 1] + 1);
    t = std::min(t, p.nstrips - (p.grid - c));
    cs[c] = t;
  }
  std::vector<int> cta_pair(p.grid + 1, 0), cta_seg(p.grid + 1, 0), seg_blk;
  for(int c = 0; c < p.grid; ++c) {
    cta_pair[c] = (int)strip_first[(size_t)cs[c]];
    cta_seg[c] = (int)seg_blk.size();
    for(long t = cs[c]; t < cs[c + 1]; ++t) {
      if(t == cs[c] || strip_blk[(size_t)t] != strip_blk[(size_t)t - 1])
        seg_blk.push_back(strip_blk[(size_t)t]);
      if(t + 1 == cs[c + 1] || strip_blk[(size_t)t + 1] != strip_blk[(size_t)t])
        pinfo[(size_t)strip_first[(size_t)t + 1] - 1].w |= 2;
    }
  }
  cta_pair[p.grid] = (int)p.npairs;
  cta_seg[p.grid] = (int)seg_blk.size();
  p.nseg = (int)seg_blk.size();
  std::vector<int> blk_seg(p.nblocks + 1, 0);
  {
    size_t k = 0;
    for(int b = 0; b < p.nblocks; ++b) {
      blk_seg[b] = (int)k;
      while(k < seg_blk.size() && seg_blk[k] == b)
        ++k;
    }
    blk_seg[p.nblocks] = (int)seg_blk.size();
  }
  p.pair_ij = rot_upload(ij);
  p.pinfo = rot_upload(pinfo);
  p.cta_pair = rot_upload(cta_pair);
  p.cta_seg = rot_upload(cta_seg);
  p.blk_seg = rot_upload(blk_seg);
  p.blk_strip = rot_upload(blk_strip);
  const size_t n2 = 2 * (size_t)L.n;
  OB_CUDA(cudaMalloc(&p.rowpart, std::max<size_t>(1, (size_t)p.nseg * I) * n2 * sizeof(cplx)));
  OB_CUDA(cudaMalloc(&p.colpart, std::max<size_t>(1, (size_t)p.nstrips) * n2 * sizeof(cplx)));
  OB_CUDA(cudaMalloc(&p.acc, (size_t)nobj * n2 * sizeof(cplx)));
}

// ---------------------------------------------------------------------------------------------
// host: launches
// ---------------------------------------------------------------------------------------------
static int g_rot_assembly = 1; // 1 = axial-only recursion (k_assemble_axial_only, default), 0 = vtac_block at theta = 0
void rot_tuning(int assembly, int rows, int ctas_per_sm) {
  if(assembly >= 0)
    g_rot_assembly = assembly;
  if(rows >= 0)
    g_rot_rows = rows;
  if(ctas_per_sm >= 0)
    g_rot_ctas = ctas_per_sm;
}
void launch_assemble_rot(VtacTableSet const &ts, const double *xyz, cplx k, const int2 *pair_ij, long npairs,
                         unsigned char *recs, RotLayout const &L, int sm_count, cudaStream_t st) {
  if(npairs <= 0)
    return;
  RotDTable const &dt = rot_dtable(L.NM);
  if(g_rot_assembly == 1) {
    const size_t sm = rot_axial_smem_bytes(L.NM);
    OB_CUDA(cudaFuncSetAttribute((const void *)k_assemble_axial_only, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const long ctas = std::min<long>((npairs + ROT_AX_WARPS - 1) / ROT_AX_WARPS, (long)sm_count * 8);
    k_assemble_axial_only<<<(unsigned)ctas, ROT_AX_WARPS * 32, sm, st>>>(xyz, k, pair_ij, npairs, recs, L);
  } else {
    OB_CUDA(cudaFuncSetAttribute((const void *)k_assemble_axial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts.smem));
    OB_CUDA(cudaFuncSetAttribute((const void *)k_assemble_axial, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    k_assemble_axial<<<(unsigned)npairs, OB_VTAC_THREADS, ts.smem, st>>>(ts.tb, xyz, k, pair_ij, recs, L);
  }
  OB_CUDA(cudaGetLastError());
  const long tctas = std::min<long>(npairs, (long)sm_count * 16);
  k_rot_tables<<<(unsigned)tctas, ROT_TAB_THREADS, 0, st>>>(xyz, pair_ij, npairs, recs, L, dt.coef, dt.seed);
  OB_CUDA(cudaGetLastError());
}

void launch_matvec_rot(RotPlan const &p, RotLayout const &L, const unsigned char *recs, const cplx *x, const cplx *Tdiag,
                       cplx *acc_or_y, int finalize, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1) {
  if(e0)
    cudaEventRecord(e0, st);
  if(p.npairs > 0) {
    OB_CUDA(cudaFuncSetAttribute((const void *)k_matvec_rot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    OB_CUDA(cudaFuncSetAttribute((const void *)k_matvec_rot, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    RotArgs a;
    a.recs = recs;
    a.x = x;
    a.pinfo = p.pinfo;
    a.cta_pair = p.cta_pair;
    a.cta_seg = p.cta_seg;
    a.rowpart = p.rowpart;
    a.colpart = p.colpart;
    a.L = L;
    a.I = p.I;
    k_matvec_rot<<<p.grid, p.threads, p.smem, st>>>(a);
    OB_CUDA(cudaGetLastError());
  }
  if(e1)
    cudaEventRecord(e1, st);
  const int n2 = 2 * p.n;
  const int thr = std::max(1, ROT_REDUCE_THREADS / n2) * n2;
  k_rot_reduce<<<p.nobj, thr, 0, st>>>(p.rowpart, p.colpart, p.blk_seg, p.blk_strip, p.strip0, p.nstrips, p.I, n2, x, Tdiag,
                                       acc_or_y, finalize);
  OB_CUDA(cudaGetLastError());
}

} // namespace ob
