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
// Record per unordered pair (i < j):
//   ph[m + NM] = exp(i m phi)                                                      (2 NM + 1 complex)
//   Cp = A + B as two planes of doubles [Re | Im], one w x w block per order a = 0..NM at offX(a), n0 = max(a,1),
//        w = NM - n0 + 1, holding (n,a),(l,a);   Cm = A - B likewise for a = 1..NM (a = 0: B = 0)
//   Ds (n + 1) x (n + 1) at offDs(n),  Da n x n at offDa(n) for every degree n            (reals)
// 21.4 KB at nMax 10 (v1 30 KB, pair form 460.8 KB), 11.7 KB at nMax 8.
// Inside a block the entries are in FRAGMENT ORDER (rot_frag_index / rot_cidx in ob_rot_axial.cuh): every DMMA A
// fragment the apply loads is one contiguous run in lane order, compacted to its valid rows and K entries: same bytes
// as the plain row-major blocks, but 2 shared-memory wavefronts per fragment load instead of 4 (a half-warp reads 16
// consecutive doubles; with the odd leading dimensions of these blocks it hit every bank twice).  The assembly
// kernels write through the transpose symmetries (C(n,l) = (-1)^(n+l) C(l,n), D[a',a] = (-1)^(a'-a) D[a,a']) so that
// the lanes of one recursion level fill runs of consecutive K entries.
//
// Work decomposition: rows are grouped in blocks of I; a strip (b, j) holds the pairs (i, j), i in block b, i < j.
// Records are stored strip by strip (block-major); ranks and CTAs own contiguous strip ranges of equal pair counts.
// Inside a CTA the row-side sums of the I rows of the current block and the column-side sums of the current strip
// live in shared memory: one column partial per STRIP (not per pair) and one row partial per block segment reach HBM,
// added in a fixed order by k_rot_reduce (deterministic, no atomics).
//
// Apply, per pair and for both directions at once: four complex vectors (x_j TE/TM, parity-signed x_i TE/TM) = EIGHT
// REAL COLUMNS, which is exactly the N of the FP64 tensor-core instruction DMMA.8x8x4 (mma.sync m8n8k4 f64; measured
// 37.1 TFLOP/s on B200, the DFMA pipe gives 34.0).  The first version of this kernel used one thread per output and
// DFMA; ncu showed it bound by shared-memory wavefronts (4930 per pair, 89 % of the pipe: a 128-bit shared load costs
// four wavefronts even when every lane reads the same address, so the 9 bytes per FMA of the thread-per-output
// scheme could not be fed).  DMMA shares the operands across the warp in the tensor core: 2 bytes per FMA.
//   P0  t = exp(i m phi) x -> flip basis (thread per (n, a); x prefetched into registers during the previous pair)
//   P1  u = D^T t: per degree n and class, M = outputs a (tiles of 8), K = a' (steps of 4), N = 8 real columns;
//       a warp runs the s and a class of one (n, tile): both results of (n, a) meet in one lane and the channel sums
//       p+- = TE_s +- TM_a, r+- = TM_s +- TE_a need one shuffle
//   P2  q = C p per order a: complex matrix as two real DMMAs (Re C, Im C) on the same B fragment, A + B and A - B
//       channels in the same lanes -> class vectors v without any exchange
//   P3  w = D v (same arrays and access pattern as P1);  P4  flip basis -> m, conjugate phase, parity signs,
//       accumulate (owner lanes add into the shared-memory row / column sums and flush them at strip / segment ends)
// The record of the next pair arrives by cp.async.bulk (TMA bulk copies, mbarrier complete_tx) in ONE shared-memory
// slot that is refilled section by section: small-d part and next phases after P1 (P3 reuses the fragments P1 kept in
// registers), axial part after P2 (ROT_SINGLE_SLOT).  Three barriers per pair, three CTAs per SM at nMax 10.
#include "ob_internal.h"
#include "ob_vtac.cuh"
#include "ob_rot_axial.cuh"
#include <algorithm>
#include <cstring>
#include <utility>

namespace ob {

#define ROT_SQH 0.70710678118654752440 // 1 / sqrt(2)

// ---------------------------------------------------------------------------------------------
// layout helpers (host + device), mirrored by tests/rot2_model.py
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr inline int rot_offDs(int n) { return n * (n + 1) * (2 * n + 1) / 6 - 1; } // sum_{j<n} (j+1)^2
__host__ __device__ constexpr inline int rot_offDa(int n) { return (n - 1) * n * (2 * n - 1) / 6; }     // sum_{j<n} j^2
__host__ __device__ constexpr inline int rot_offF(int n) { return (n - 1) * (n + 2); }
__host__ __device__ constexpr inline int rot_offP(int NM, int a) { // channel index of (a, l = n0): sum_{u<a} (NM - max(u,1) + 1)
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
  L.offCp = (size_t)(2 * NM + 1) * sizeof(cplx);                       // Re plane, then Im plane (X doubles each)
  L.offCm = L.offCp + (size_t)L.X * sizeof(cplx);                      // Re plane, then Im plane (X - NM^2 doubles each)
  L.offDs = L.offCm + (size_t)(L.X - NM * NM) * sizeof(cplx);
  L.offDa = (L.offDs + (size_t)L.nDs * sizeof(double) + 15) & ~(size_t)15;
  L.rec_bytes = (L.offDa + (size_t)L.nDa * sizeof(double) + 15) & ~(size_t)15;
  return L;
}

// ---------------------------------------------------------------------------------------------
// assembly 1 (cross-check path, "rot_assembly" = 0): axial A, B out of the shared VTAC block code at theta = phi = 0
// ---------------------------------------------------------------------------------------------
struct EmitAxial {
  double *Cp, *Cm; // [Re | Im] planes
  int NM, X;
  // p = flat(n, mu) (first index of Coupling.diagonal), r = flat(l, k): keep mu == k >= 0
  __device__ __forceinline__ void item(int p, int r, cplx a, cplx b) {
    int n, mu, l, k;
    unflatten(p, n, mu);
    unflatten(r, l, k);
    if(mu != k || mu < 0)
      return;
    const int e = rot_cidx(NM, mu, n, l); // fragment order (ob_rot_axial.cuh)
    Cp[e] = a.x + b.x;
    Cp[X + e] = a.y + b.y;
    if(mu >= 1) {
      Cm[e - NM * NM] = a.x - b.x;
      Cm[X - NM * NM + e - NM * NM] = a.y - b.y;
    }
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
  em.Cp = (double *)(rec + L.offCp);
  em.Cm = (double *)(rec + L.offCm);
  em.NM = L.NM;
  em.X = L.X;
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
// One persistent CTA per SM with as many warps as fit beside the TABLES IN SHARED MEMORY: the first tabulated version
// read them from global memory with 212 KB of the SM's 256 KB configured as shared memory, so the 80 KB of tables
// thrashed what was left of L1 and every item paid two dependent L2 round trips (8.9 ms per C5 harmonic).
#define ROT_AX_MAX_WARPS 16
struct RotAxSizes {
  int nrec, nem; // table entries (recursion, emission)
};
static RotAxSizes rot_axial_sizes(int NM) { return {rot_axial_offR(NM, NM + 1), rot_axial_offE(NM, NM + 1)}; }
static size_t rot_axial_table_bytes(int NM) {
  const RotAxSizes z = rot_axial_sizes(NM);
  return ((size_t)z.nrec * (4 * sizeof(double) + sizeof(int)) + (size_t)z.nem * (8 * sizeof(double) + 2 * sizeof(int)) + 15) &
         ~(size_t)15;
}
// TAB_SMEM: the tables are read through pointers the compiler can see are shared memory (LDS with 32-bit addresses).
// The first version selected between the global and the shared copy at run time: every table read became a generic LD
// (64-bit address arithmetic, long-scoreboard latency: 40 % of the stall samples, profiles/r4k_assemble_axial_only_full.txt)
template <bool TAB_SMEM>
__global__ void __launch_bounds__(ROT_AX_MAX_WARPS * 32)
k_assemble_axial_only(const double *__restrict__ xyz, cplx k, const int2 *__restrict__ pair_ij, long npairs,
                      unsigned char *__restrict__ recs, RotLayout L, RotAxTab gtab, int nrec, int nem) {
  extern __shared__ __align__(16) unsigned char smem_ax[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  auto pairs = [&](RotAxTab const &tab, unsigned char *wbuf) {
    cplx *buf = (cplx *)wbuf + (size_t)warp * rot_axial_fast_entries(L.NM);
    for(long q = (long)blockIdx.x * nwarps + warp; q < npairs; q += (long)gridDim.x * nwarps) {
      const int2 ij = pair_ij[q];
      const double x = xyz[3 * ij.x] - xyz[3 * ij.y], y = xyz[3 * ij.x + 1] - xyz[3 * ij.y + 1],
                   z = xyz[3 * ij.x + 2] - xyz[3 * ij.y + 2];
      const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
      unsigned char *rec = recs + (size_t)q * L.rec_bytes;
      rot_axial_pair_fast(L.NM, k, r, buf, (double *)(rec + L.offCp), (double *)(rec + L.offCm), lane, 32, tab);
      __syncwarp();
    }
  };
  if constexpr(TAB_SMEM) { // doubles first (16-byte aligned), then the three index arrays, then the warps' level buffers
    double *s_rec = (double *)smem_ax, *s_emit = s_rec + (size_t)nrec * 4;
    int *s_ridx = (int *)(s_emit + (size_t)nem * 8), *s_eidx = s_ridx + nrec, *s_eout = s_eidx + nem;
    for(int e = threadIdx.x; e < nrec * 4; e += blockDim.x)
      s_rec[e] = gtab.rec[e];
    for(int e = threadIdx.x; e < nem * 8; e += blockDim.x)
      s_emit[e] = gtab.emit[e];
    for(int e = threadIdx.x; e < nrec; e += blockDim.x)
      s_ridx[e] = gtab.ridx[e];
    for(int e = threadIdx.x; e < nem; e += blockDim.x) {
      s_eidx[e] = gtab.eidx[e];
      s_eout[e] = gtab.eout[e];
    }
    __syncthreads();
    const RotAxTab tab = {s_rec, s_emit, s_ridx, s_eidx, s_eout, nrec, nem};
    pairs(tab, smem_ax + (((size_t)nrec * (4 * sizeof(double) + sizeof(int)) + (size_t)nem * (8 * sizeof(double) + 2 * sizeof(int)) + 15) &
                          ~(size_t)15));
  } else
    pairs(gtab, smem_ax);
}

// index-only coefficient tables of the axial recursion, one set per (device, nMax)
struct RotAxTabDev {
  int NM = -1;
  RotAxTab t = {nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
};
static RotAxTabDev g_axtab[16][OB_MAX_NMAX + 1];
template <class T> static const T *rot_to_device(std::vector<T> const &v) {
  T *d = nullptr;
  OB_CUDA(cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(T)));
  OB_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
static RotAxTab const &rot_axtab(int NM) {
  int dev = 0;
  OB_CUDA(cudaGetDevice(&dev));
  if(dev < 0 || dev >= 16)
    throw Error("rotated-axial operator: device ordinal out of range");
  RotAxTabDev &e = g_axtab[dev][NM];
  if(e.NM == NM)
    return e.t;
  std::vector<double> rec, emit;
  std::vector<int> ridx, eidx, eout;
  rot_axial_tables_build(NM, rec, emit, ridx, eidx, eout);
  e.t.rec = rot_to_device(rec);
  e.t.emit = rot_to_device(emit);
  e.t.ridx = rot_to_device(ridx);
  e.t.eidx = rot_to_device(eidx);
  e.t.eout = rot_to_device(eout);
  e.t.nrec = (int)ridx.size();
  e.t.nem = (int)eidx.size();
  e.NM = NM;
  return e.t;
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
        // flip basis (tests/rot2_model.py build_pair): D[a' = mp, a = m].  Fragment order (ob_rot_axial.cuh) has row a and
        // K index a'; this thread's (mp, m) would be one K index of row m, 32 bytes from its neighbour's.  With
        // D[a', a] = (-1)^(a' - a) D[a, a'] it writes the entry (a' = m, a = mp) instead: consecutive threads fill
        // consecutive K entries of row mp
        const double st = ((mp - m) & 1) ? -1.0 : 1.0;
        double vs;
        if(mp == 0 && m == 0)
          vs = dP;
        else if(m == 0 || mp == 0)
          vs = 1.41421356237309504880 * dP;
        else {
          vs = dP + sa * dM;
          Da[rot_offDa(j) + rot_frag_index_a(j, mp, m - 1)] = st * (dP - sa * dM);
        }
        Ds[rot_offDs(j) + rot_frag_index(j + 1, j + 1, mp, m)] = st * vs;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// apply
// ---------------------------------------------------------------------------------------------
#ifndef ROT_WARPS
#define ROT_WARPS 4
#endif
// ROT_SINGLE_SLOT = 1: ONE record slot per CTA, refilled section by section as the phases release it (small-d part and
// the next phases after P1, axial part after P2): 21 KB less shared memory per CTA and loop-invariant section pointers
// (3.90 -> 3.78 ms per C5 apply at three CTAs per SM).  0: two whole-record slots.
// ROT_MIN_CTAS: resident CTAs per SM the kernel is compiled for at nMax <= 10.  With the single slot and three rows per
// block a FOURTH CTA fits (55 KB, 128 registers, 24 bytes of spills): measured 3.88 ms, slower than three CTAs at 168
// registers (16 % more instructions under the register cap; the shared-memory pipe is already 60 % busy).
#ifndef ROT_SINGLE_SLOT
#define ROT_SINGLE_SLOT 1
#endif
#ifndef ROT_SPLIT_CHAINS
#define ROT_SPLIT_CHAINS 0 // 1: odd and even K steps of a DMMA chain accumulate separately (shorter dependency chains; measured slower: 3.88 against 3.66 ms per C5 apply, the extra accumulators cost registers)
#endif
#ifndef ROT_MIN_CTAS
#define ROT_MIN_CTAS 3
#endif
__host__ __device__ constexpr inline int rot_min_ctas(int NM) { return NM <= 10 ? ROT_MIN_CTAS : (ROT_MIN_CTAS < 3 ? ROT_MIN_CTAS : 3); }
#define ROT_THREADS (32 * ROT_WARPS)
#define ROT_MAX_UNITS 32
// Static work lists, evaluated at COMPILE TIME per nMax (the kernel is a template on nMax: every offset, stride, K-step
// count and tail condition below is an immediate; the first DMMA version took them from run-time tables and spent
// 33 instructions per DMMA on index arithmetic).  d-phase units: degree n, the s and the a class together, every 8-row tile
// (the tiles of a unit share its B fragments: one load per K step).  P2 units: order a, every 8-row tile, the A + B
// and the A - B channels together.  Units go to the warps by longest-processing-time (cost = DMMAs).
struct RotCT {
  int nd, nc;
  int dn[ROT_MAX_UNITS], dw[ROT_MAX_UNITS], dc[ROT_MAX_UNITS];
  int ca[ROT_MAX_UNITS], cw[ROT_MAX_UNITS], cc[ROT_MAX_UNITS];
};
__host__ __device__ constexpr inline int rot_tiles(int rows) { return (rows + 7) / 8; } // row tiles of 8 (at most two)
__host__ __device__ constexpr inline void rot_ct_assign(int count, const int *cost, int *owner) {
  bool done[ROT_MAX_UNITS] = {};
  int load[ROT_WARPS] = {};
  for(int it = 0; it < count; ++it) {
    int best = -1;
    for(int u = 0; u < count; ++u)
      if(!done[u] && (best < 0 || cost[u] > cost[best]))
        best = u;
    int w = 0;
    for(int i = 1; i < ROT_WARPS; ++i)
      if(load[i] < load[w])
        w = i;
    owner[best] = w;
    load[w] += cost[best];
    done[best] = true;
  }
}
__host__ __device__ constexpr inline RotCT rot_ct(int NM) {
  RotCT T{};
  for(int n = 1; n <= NM; ++n) { // one d-phase unit per degree: all row tiles share the B fragments
    T.dn[T.nd] = n;
    T.dc[T.nd] = rot_tiles(n + 1) * ((n + 1 + 3) / 4 + (n + 3) / 4);
    ++T.nd;
  }
  for(int a = 0; a <= NM; ++a) { // one P2 unit per order
    const int w = NM - rot_n0(a) + 1;
    T.ca[T.nc] = a;
    // DMMAs: two (Re, Im) per K step, channel and full tile; one for a stacked tile of at most four rows
    const int full = w / 8, rest = w - 8 * full;
    T.cc[T.nc] = (2 * full + (rest > 4 ? 2 : rest > 0 ? 1 : 0)) * ((w + 3) / 4) * (a == 0 ? 1 : 2);
    ++T.nc;
  }
  rot_ct_assign(T.nd, T.dc, T.dw);
  rot_ct_assign(T.nc, T.cc, T.cw);
  return T;
}
static_assert(OB_MAX_NMAX + 1 <= 16, "the unit code handles at most two row tiles");
static_assert(rot_ct(OB_MAX_NMAX).nd <= ROT_MAX_UNITS && rot_ct(OB_MAX_NMAX).nc <= ROT_MAX_UNITS, "unit tables too small");

struct RotArgs {
  const unsigned char *recs;
  const unsigned char *geo; // records the phases and small-d sections are fetched from (recs itself, or the other harmonic's)
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
// D (8x8) += A (8x4, row) B (4x8, col), FP64 tensor core.  Fragments: A[lane >> 2][lane & 3], B[lane & 3][lane >> 2],
// D[lane >> 2][2 (lane & 3) + {0, 1}]
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Shared-memory vector buffers (doubles).  Class vectors T / V: [direction][F index][TE re, TE im, TM re, TM im], the
// two directions ("planes") PS doubles apart; channels P: [A+B | A-B][direction][(a, l)][p re, p im, r re, r im].  A B
// fragment (row = lane & 3 consecutive F / channel indices, column = lane >> 2) reads 128 contiguous bytes per
// half-warp: plane = lane >> 4, position inside the 32-byte row = (lane >> 2) & 3.  Rows past the valid K range are
// read (their A entries are forced to zero) and must stay finite: the planes carry eight spare rows and start zeroed.
// A-fragment rows past the valid outputs of a tile are read from wherever they fall inside the CTA's shared memory:
// rows of a matrix product are independent and those results are never stored.
__host__ __device__ constexpr inline int rot_plane_doubles(int LF) { return 4 * (LF + 8); }
// Accumulator rows (complex units): TE at 0, TM at nHp = nH rounded up to 2 mod 8, row stride a multiple of 8; the
// column-sum row sits first and the block's rows start 4 units later.  A quarter-warp of the P4 read-modify-write
// (two consecutive outputs x four vectors = (direction, polarisation)) then touches eight distinct 16-byte bank groups
// (the plain [I + 1][2 nH] layout put the four vectors on the same banks: 16 wavefronts per access instead of 4).
__host__ __device__ constexpr inline int rot_acc_tm(int nH) { return nH + ((2 - nH % 8) + 8) % 8; }
__host__ __device__ constexpr inline int rot_acc_row(int nH) { return (rot_acc_tm(nH) + nH + 7) / 8 * 8; }
__host__ __device__ constexpr inline int rot_acc_units(int nH, int I) { return (I + 1) * rot_acc_row(nH) + 4; }
// dynamic shared memory: record[2] (single slot: record | second phase slot) | bufX[2 planes] | bufY[2 planes] |
// acc[I + 1][2n] (rows of the block, then the column sums of the strip) | mbarrier[2]
static size_t rot_slot_bytes(RotLayout const &L) { return ROT_SINGLE_SLOT ? L.rec_bytes + L.offCp : 2 * L.rec_bytes; }
static size_t rot_smem_bytes(RotLayout const &L, int I) {
  return rot_slot_bytes(L) + (size_t)4 * rot_plane_doubles(L.LF) * sizeof(double) + (size_t)rot_acc_units(L.n, I) * sizeof(cplx) +
         2 * sizeof(uint64_t);
}

// per-lane state of the pair being processed (everything else is an immediate)
struct RotLane {
  int warp, fr, fc, cpol;
  const double *Ds, *Da, *Cp, *Cm; // sections of the current record, already offset by this lane's place in a fragment (+ 4 fr + fc = lane)
  const double *vb;                // class-vector buffer + plane / in-row offset of the B fragment
  const double *pb;                // channel buffer + plane / in-row offset of the B fragment
  double *pst;                     // channel buffer + direction plane + 2 * pol (P1 stores)
  double *vst;                     // class-vector buffer + direction plane + 2 * pol (P2 stores)
  const cplx *ph;                  // phases of the current record + NM
  cplx *dst;                       // accumulator row of this lane's direction + pol * nH (P4)
  double gl;                       // sign of this lane's vector in direction 1 (without (-1)^deg)
  bool dir1;
};

// accumulation chain of KS K-steps over K columns: straight-line loads, then DMMAs; the A fragments of a row tile are
// consecutive runs in the record (fragment order, ob_rot_axial.cuh): pa = first fragment + 4 fr + fc, stepA = 4 Rt doubles
// per step; a last step of Kv < 4 valid K entries is compacted to row stride Kv (pt = its address for this lane) and
// the lane's A entry is zeroed past the K range (compile-time known).  The B fragments bv[KS] are loaded once per unit
// (rot_load_b: 16 doubles per step) and serve every row tile of the unit.
// KEEP: 0 = load the A fragments, 1 = load them and keep them in `keep` (registers), 2 = take them from `keep` (the
// small-d fragments of P1 serve P3 again: D[a', a] = (-1)^(a' - a) D[a, a'] makes both phases read the same entries)
template <int KS> __device__ __forceinline__ void rot_load_b(double *bv, const double *__restrict__ pb) {
#pragma unroll
  for(int s = 0; s < KS; ++s)
    bv[s] = pb[16 * s];
}
template <int KS, int K, int KEEP = 0>
__device__ __forceinline__ void rot_chain(double (&acc)[2], const double *__restrict__ pa, int stepA,
                                          const double *__restrict__ pt, const double *bv, int fc, double *keep = nullptr) {
  double av[KS];
#pragma unroll
  for(int s = 0; s < KS; ++s) {
    if(KEEP == 2)
      av[s] = keep[s];
    else
      av[s] = (4 * KS > K && s == KS - 1) ? pt[0] : pa[s * stepA];
  }
  if(KEEP != 2 && 4 * KS > K)
    av[KS - 1] = 4 * (KS - 1) + fc < K ? av[KS - 1] : 0.0;
  if(KEEP == 1) {
#pragma unroll
    for(int s = 0; s < KS; ++s)
      keep[s] = av[s];
  }
#if ROT_SPLIT_CHAINS
  if constexpr(KS >= 2) { // two independent accumulation chains (odd / even K steps), added at the end
    double acc2[2] = {0.0, 0.0};
#pragma unroll
    for(int s = 0; s < KS; ++s) {
      if(s & 1)
        dmma(acc2, av[s], bv[s]);
      else
        dmma(acc, av[s], bv[s]);
    }
    acc[0] += acc2[0];
    acc[1] += acc2[1];
    return;
  }
#endif
#pragma unroll
  for(int s = 0; s < KS; ++s)
    dmma(acc, av[s], bv[s]);
}
// two chains on the same B fragments: real and imaginary plane of one complex matrix (planes dI doubles apart)
template <int KS, int K>
__device__ __forceinline__ void rot_chain2(double (&accR)[2], double (&accI)[2], const double *__restrict__ pa, int dI,
                                           int stepA, const double *__restrict__ pt, const double *bv, int fc) {
  double ar[KS], ai[KS];
#pragma unroll
  for(int s = 0; s < KS; ++s) {
    const double *q = (4 * KS > K && s == KS - 1) ? pt : pa + s * stepA;
    ar[s] = q[0];
    ai[s] = q[dI];
  }
  if(4 * KS > K) {
    const bool in = 4 * (KS - 1) + fc < K;
    ar[KS - 1] = in ? ar[KS - 1] : 0.0;
    ai[KS - 1] = in ? ai[KS - 1] : 0.0;
  }
#if ROT_SPLIT_CHAINS
  if constexpr(KS >= 2) {
    double r2[2] = {0.0, 0.0}, i2[2] = {0.0, 0.0};
#pragma unroll
    for(int s = 0; s < KS; ++s) {
      if(s & 1) {
        dmma(r2, ar[s], bv[s]);
        dmma(i2, ai[s], bv[s]);
      } else {
        dmma(accR, ar[s], bv[s]);
        dmma(accI, ai[s], bv[s]);
      }
    }
    accR[0] += r2[0];
    accR[1] += r2[1];
    accI[0] += i2[0];
    accI[1] += i2[1];
    return;
  }
#endif
#pragma unroll
  for(int s = 0; s < KS; ++s) {
    dmma(accR, ar[s], bv[s]);
    dmma(accI, ai[s], bv[s]);
  }
}

// d-phase unit U of nMax NM (P1 and P3): degree n, both classes, every row tile (rows a = 8 TILE .. 8 TILE + 7).
//   accS[(a, col)] = sum_{a' = 0..n} Ds[(a', a)] v_s[a'][col],  accA[(a, col)] = sum_{a' = 1..n} Da[(a', a)] v_a[a'][col]
// register slots of the kept A fragments: units of the same warp before U (compile time), tile by tile
#ifndef ROT_KEEP_S
#define ROT_KEEP_S 1 // keep the s-class small-d fragments of P1 in registers for P3
#endif
#ifndef ROT_KEEP_A
#define ROT_KEEP_A 1 // the a-class ones too (168 registers with three CTAs per SM: no spills up to nMax 13)
#endif
static_assert(!ROT_SINGLE_SLOT || (ROT_KEEP_S && ROT_KEEP_A), "the single record slot is refilled during P3: P3 must not read the small-d part");
__host__ __device__ constexpr inline int rot_keep_per_tile(int n) {
  return (ROT_KEEP_S ? (n + 1 + 3) / 4 : 0) + (ROT_KEEP_A ? (n + 3) / 4 : 0);
}
__host__ __device__ constexpr inline int rot_keep_slots(int NM, int W, int U) { // fragments kept by W's units before U
  const RotCT T = rot_ct(NM);
  int k = 0;
  for(int u = 0; u < U; ++u)
    if(T.dw[u] == W)
      k += rot_tiles(T.dn[u] + 1) * rot_keep_per_tile(T.dn[u]);
  return k;
}
__host__ __device__ constexpr inline int rot_keep_total(int NM) {
  const RotCT T = rot_ct(NM);
  int best = 1;
  for(int w = 0; w < ROT_WARPS; ++w) {
    const int k = rot_keep_slots(NM, w, T.nd);
    best = k > best ? k : best;
  }
  return best;
}
// one row tile of a d-phase unit.  PHASE 1: P1 (fragments loaded and kept), 3: P3 (kept fragments reused); bS / bA: the
// unit's B fragments (class vectors of degree n), already in registers
template <int NM, int W, int U, int PHASE, int TILE>
__device__ __forceinline__ void rot_dchains(RotLane const &c, double (&accS)[2], double (&accA)[2], double *keep,
                                            const double *bS, const double *bA) {
  constexpr RotCT T = rot_ct(NM);
  constexpr int n = T.dn[U], m0 = 8 * TILE, n1 = n + 1, slot = rot_keep_slots(NM, W, U) + TILE * rot_keep_per_tile(n);
  constexpr int KS = ROT_KEEP_S ? (PHASE == 1 ? 1 : 2) : 0, KA = ROT_KEEP_A ? (PHASE == 1 ? 1 : 2) : 0;
  accS[0] = accS[1] = accA[0] = accA[1] = 0.0;
  // s class: rows a = m0 .. of the (n + 1) x (n + 1) matrix, Rt valid rows in this tile
  constexpr int RtS = n1 - m0 < 8 ? n1 - m0 : 8, ksS = (n1 + 3) / 4, kvS = n1 - 4 * (ksS - 1);
  const double *paS = c.Ds + rot_offDs(n) + m0 * n1;
  rot_chain<ksS, n1, KS>(accS, paS, 4 * RtS, paS + 4 * RtS * (ksS - 1) - (4 - kvS) * c.fr, bS, c.fc, keep + slot);
  // a class: rows a = max(m0, 1) .. of the n x n matrix (tile 0: a phantom row 0 in front of the seven rows a = 1 .. 7)
  constexpr int firstA = m0 ? m0 : 1, lastA = m0 + 7 < n ? m0 + 7 : n, RtA = lastA - firstA + 1, ksA = (n + 3) / 4,
                kvA = n - 4 * (ksA - 1), radj = m0 ? 0 : -1;
  const double *paA = c.Da + rot_offDa(n) + (firstA - 1) * n + 4 * radj;
  rot_chain<ksA, n, KA>(accA, paA, 4 * RtA, paA + 4 * RtA * (ksA - 1) - (4 - kvA) * c.fr - (4 - kvA) * radj, bA, c.fc,
                        keep + slot + (ROT_KEEP_S ? ksS : 0));
}

// P1: u = D^T t, channel sums p+- = TE_s +- TM_a (TE lanes), r+- = TM_s +- TE_a (TM lanes)
template <int NM, int W, int U, int TILE>
__device__ __forceinline__ void rot_p1_tile(RotLane const &c, double *keep, const double *bS, const double *bA) {
  constexpr RotCT T = rot_ct(NM);
  constexpr int n = T.dn[U], m0 = 8 * TILE, SS = rot_plane_doubles(NM * (NM + 3));
  double aS[2], aA[2];
  rot_dchains<NM, W, U, 1, TILE>(c, aS, aA, keep, bS, bA);
  const int aa = m0 + c.fr;
  if(m0 == 0 && c.fr == 0) // the a class has no a = 0 row (its fragment row was read from outside the block)
    aA[0] = aA[1] = 0.0;
  // lane (row a, vector fc): TE lanes need the partner's u_a of TM and vice versa: ch+- = u_s +- u_a(partner)
  const double ox = __shfl_xor_sync(0xffffffffu, aA[0], 1), oy = __shfl_xor_sync(0xffffffffu, aA[1], 1);
  if(m0 + 7 <= n || aa <= n) { // a = 0: both channels equal u_s
    // channel index of (a, l = n): offP(a) + n - max(a, 1)
    const int idx = (aa == 0 ? 0 : NM + (aa - 1) * (NM + 1) - (aa - 1) * aa / 2) + n - (aa > 1 ? aa : 1);
    double *P = c.pst + 4 * idx;
    *(cplx *)P = mk(aS[0] + ox, aS[1] + oy);
    *(cplx *)(P + SS) = mk(aS[0] - ox, aS[1] - oy);
  }
}
template <int NM, int W, int U> __device__ __forceinline__ void rot_p1_unit(RotLane const &c, double *keep) {
  constexpr RotCT T = rot_ct(NM);
  constexpr int n = T.dn[U];
  if constexpr(T.dw[U] != W)
    return;
  double bS[(n + 4) / 4], bA[(n + 3) / 4]; // class vectors of degree n: one load per K step for all row tiles
  rot_load_b<(n + 4) / 4>(bS, c.vb + 4 * rot_offF(n) + 4 * c.fc);
  rot_load_b<(n + 3) / 4>(bA, c.vb + 4 * (rot_offF(n) + n + 2) + 4 * c.fc);
  rot_p1_tile<NM, W, U, 0>(c, keep, bS, bA);
  if constexpr(rot_tiles(n + 1) > 1)
    rot_p1_tile<NM, W, U, 1>(c, keep, bS, bA);
}

// P2: q = C p for order a, rows n = n0 + 8 TILE .. + 7 (Re and Im of C as two real DMMAs on one B fragment), back to
// the class vectors: v_s = (-1)^a (q+ + q-) / 2, v_a = (-1)^a (q+ - q-) / 2.  bP / bM: channel vectors of the A + B and
// the A - B channel of order a, in registers for every row tile
template <int NM, int W, int U, int TILE>
__device__ __forceinline__ void rot_p2_tile(RotLane const &c, const double *bP, const double *bM) {
  constexpr RotCT T = rot_ct(NM);
  constexpr int a = T.ca[U], m0 = 8 * TILE, n0 = rot_n0(a), w = NM - n0 + 1, ks = (w + 3) / 4;
  constexpr int XC = rot_offX(NM, NM + 1), XM = XC - NM * NM;
  constexpr int Rt = w - m0 < 8 ? w - m0 : 8, kv = w - 4 * (ks - 1); // fragment order: tile m0 / 8 of the w x w block of order a
  constexpr double h = (a & 1) ? -0.5 : 0.5; // (-1)^a of the transposed small-d read, and the 1/2 of the channel split
  if constexpr(Rt <= 4) {
    // STACKED tile (at most four rows): lanes fr < 4 take the Re C rows, lanes fr >= 4 the Im C rows fr - 4 of the same
    // fragment position in the other plane: one DMMA per K step instead of two; the halves meet by one shuffle
    const int hi = c.fr >> 2, r = c.fr & 3;
    const int lofs = rot_offX(NM, a) + m0 * w - 16 * hi; // c.Cp carries + lane = + 16 hi + 4 r + fc
    const int tofs = lofs + 4 * Rt * (ks - 1) - (4 - kv) * r;
    double ap[2] = {0, 0};
    rot_chain<ks, w>(ap, c.Cp + hi * XC + lofs, 4 * Rt, c.Cp + hi * XC + tofs, bP, c.fc);
    const double opx = __shfl_xor_sync(0xffffffffu, ap[0], 16), opy = __shfl_xor_sync(0xffffffffu, ap[1], 16);
    const double qpx = ap[0] - opy, qpy = ap[1] + opx; // valid in the lanes fr < 4 (Re rows own, Im rows from the partner)
    const int n = n0 + m0 + r;
    double *V = c.vst + 4 * ((n - 1) * (n + 2) + a);
    if(a == 0) {
      if(c.fr < Rt)
        *(cplx *)V = mk(qpx, qpy);
    } else {
      double am[2] = {0, 0};
      rot_chain<ks, w>(am, c.Cm + hi * XM + lofs, 4 * Rt, c.Cm + hi * XM + tofs, bM, c.fc);
      const double omx = __shfl_xor_sync(0xffffffffu, am[0], 16), omy = __shfl_xor_sync(0xffffffffu, am[1], 16);
      const double qmx = am[0] - omy, qmy = am[1] + omx;
      if(c.fr < Rt) {
        *(cplx *)V = mk(h * (qpx + qmx), h * (qpy + qmy));
        *(cplx *)(V + 4 * (n + 1) + 2 - 4 * c.cpol) = mk(h * (qpx - qmx), h * (qpy - qmy));
      }
    }
    return;
  }
  double apr[2] = {0, 0}, api[2] = {0, 0};
  const int tofs = rot_offX(NM, a) + m0 * w + 4 * Rt * (ks - 1) - (4 - kv) * c.fr;
  rot_chain2<ks, w>(apr, api, c.Cp + rot_offX(NM, a) + m0 * w, XC, 4 * Rt, c.Cp + tofs, bP, c.fc);
  // complex products: (Re C p_re - Im C p_im, Re C p_im + Im C p_re); lane = (row n, channel fc = direction * 2 + family)
  const double qpx = apr[0] - api[1], qpy = apr[1] + api[0];
  const int row = m0 + c.fr, n = n0 + row;
  double *V = c.vst + 4 * ((n - 1) * (n + 2) + a);
  if(a == 0) { // p- = p+: v_s = q+ (TE_s from the p family, TM_s from the r family), no a class
    if(m0 + 7 < w || row < w)
      *(cplx *)V = mk(qpx, qpy);
  } else {
    double amr[2] = {0, 0}, ami[2] = {0, 0};
    rot_chain2<ks, w>(amr, ami, c.Cm + rot_offX(NM, a) + m0 * w, XM, 4 * Rt, c.Cm + tofs, bM, c.fc);
    const double qmx = amr[0] - ami[1], qmy = amr[1] + ami[0];
    if(m0 + 7 < w || row < w) {
      *(cplx *)V = mk(h * (qpx + qmx), h * (qpy + qmy));                                   // v_s: TE (p) / TM (r)
      *(cplx *)(V + 4 * (n + 1) + 2 - 4 * c.cpol) = mk(h * (qpx - qmx), h * (qpy - qmy)); // v_a: TM (p) / TE (r)
    }
  }
}
template <int NM, int W, int U> __device__ __forceinline__ void rot_p2_unit(RotLane const &c, double *) {
  constexpr RotCT T = rot_ct(NM);
  constexpr int a = T.ca[U], n0 = rot_n0(a), w = NM - n0 + 1, ks = (w + 3) / 4, SS = rot_plane_doubles(NM * (NM + 3));
  if constexpr(T.cw[U] != W)
    return;
  double bP[ks], bM[ks];
  const double *pb = c.pb + 4 * rot_offP(NM, a) + 4 * c.fc;
  rot_load_b<ks>(bP, pb);
  if constexpr(a > 0)
    rot_load_b<ks>(bM, pb + SS);
  rot_p2_tile<NM, W, U, 0>(c, bP, bM);
  if constexpr(rot_tiles(w) > 1)
    rot_p2_tile<NM, W, U, 1>(c, bP, bM);
}

// P3 + P4: w = D v, flip basis -> m, conjugate phase, parity signs of direction 1, accumulate (owner lanes)
template <int NM, int W, int U, int TILE>
__device__ __forceinline__ void rot_p3_tile(RotLane const &c, double *keep, const double *bS, const double *bA) {
  constexpr RotCT T = rot_ct(NM);
  constexpr int n = T.dn[U], m0 = 8 * TILE;
  double aS[2], aA[2];
  rot_dchains<NM, W, U, 3, TILE>(c, aS, aA, keep, bS, bA);
  const int ap = m0 + c.fr;
  if(m0 + 7 > n && ap > n)
    return;
  const double g = (c.dir1 && (n & 1)) ? -c.gl : c.gl; // direction 1: (-1)^deg
  cplx *d = c.dst + (n * (n + 1) - 1) - ap;
  if(m0 == 0 && ap == 0) { // a' = 0: the s class alone, exp(i 0 phi) = 1
    d[0] = cadd(d[0], mk(g * aS[0], g * aS[1]));
    return;
  }
  const double fp = (ap & 1) ? -g * ROT_SQH : g * ROT_SQH, fm = g * ROT_SQH;
  const cplx php = c.ph[ap], phm = c.ph[-ap];
  const double sx = fp * (aS[0] + aA[0]), sy = fp * (aS[1] + aA[1]);
  const double dx = fm * (aS[0] - aA[0]), dy = fm * (aS[1] - aA[1]);
  // conj(phase) * value
  const cplx o1 = d[0], o2 = d[2 * ap];
  d[0] = mk(o1.x + (php.x * sx + php.y * sy), o1.y + (php.x * sy - php.y * sx));
  d[2 * ap] = mk(o2.x + (phm.x * dx + phm.y * dy), o2.y + (phm.x * dy - phm.y * dx));
}
template <int NM, int W, int U> __device__ __forceinline__ void rot_p3_unit(RotLane const &c, double *keep) {
  constexpr RotCT T = rot_ct(NM);
  constexpr int n = T.dn[U];
  if constexpr(T.dw[U] != W)
    return;
  double bS[(n + 4) / 4], bA[(n + 3) / 4];
  rot_load_b<(n + 4) / 4>(bS, c.vb + 4 * rot_offF(n) + 4 * c.fc);
  rot_load_b<(n + 3) / 4>(bA, c.vb + 4 * (rot_offF(n) + n + 2) + 4 * c.fc);
  rot_p3_tile<NM, W, U, 0>(c, keep, bS, bA);
  if constexpr(rot_tiles(n + 1) > 1)
    rot_p3_tile<NM, W, U, 1>(c, keep, bS, bA);
}

// The units of ONE warp, selected at compile time, inlined into one straight-line block: the loads and DMMAs of a warp's
// three or four units interleave (with a run-time owner test per unit every unit was its own basic block).
template <int NM, int W, int... U> __device__ __forceinline__ void rot_p1_warp(RotLane const &c, double *keep, std::integer_sequence<int, U...>) {
  (rot_p1_unit<NM, W, U>(c, keep), ...);
}
template <int NM, int W, int... U> __device__ __forceinline__ void rot_p2_warp(RotLane const &c, double *keep, std::integer_sequence<int, U...>) {
  (rot_p2_unit<NM, W, U>(c, keep), ...);
}
template <int NM, int W, int... U> __device__ __forceinline__ void rot_p3_warp(RotLane const &c, double *keep, std::integer_sequence<int, U...>) {
  (rot_p3_unit<NM, W, U>(c, keep), ...);
}
#define ROT_PER_WARP(fn, seq)                                                                                          \
  switch(c.warp) {                                                                                                     \
  case 0:                                                                                                              \
    fn<NM, 0>(c, keep, seq);                                                                                                 \
    break;                                                                                                             \
  case 1:                                                                                                              \
    fn<NM, 1>(c, keep, seq);                                                                                                 \
    break;                                                                                                             \
  case 2:                                                                                                              \
    fn<NM, 2>(c, keep, seq);                                                                                                 \
    break;                                                                                                             \
  case 3:                                                                                                              \
    fn<NM, 3>(c, keep, seq);                                                                                                 \
    break;                                                                                                             \
  case 4:                                                                                                              \
    fn<NM, 4>(c, keep, seq);                                                                                                 \
    break;                                                                                                             \
  case 5:                                                                                                              \
    fn<NM, 5>(c, keep, seq);                                                                                                 \
    break;                                                                                                             \
  case 6:                                                                                                              \
    fn<NM, 6>(c, keep, seq);                                                                                                 \
    break;                                                                                                             \
  default:                                                                                                             \
    fn<NM, 7>(c, keep, seq);                                                                                                 \
    break;                                                                                                             \
  }
static_assert(ROT_WARPS == 4 || ROT_WARPS == 8, "ROT_PER_WARP enumerates up to eight warps (no unit is owned by a warp >= ROT_WARPS)");

template <int NM>
__global__ void __launch_bounds__(ROT_THREADS, rot_min_ctas(NM)) k_matvec_rot(const __grid_constant__ RotArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr RotCT T = rot_ct(NM);
  constexpr int nH = NM * (NM + 2), n2 = 2 * nH, LF = NM * (NM + 3), NH = LF / 2;
  constexpr int PS = rot_plane_doubles(LF); // plane stride of the class-vector buffers (doubles)
  constexpr int PP = PS / 2;                // channel buffers: plane stride (the A+B and A-B halves are PS apart)
  const RotLayout L = a.L;
  const int I = a.I;
  const uint32_t slot_bytes = ROT_SINGLE_SLOT ? (uint32_t)(L.rec_bytes + L.offCp) : (uint32_t)(2 * L.rec_bytes);
  double *bufX = (double *)(smem + slot_bytes);
  double *bufY = bufX + 2 * PS;
  constexpr int NHP = rot_acc_tm(nH), RS = rot_acc_row(nH);
  cplx *acc = (cplx *)(bufY + 2 * PS); // column sums of the strip, then (4 units later) the I rows of the block
  uint64_t *full = (uint64_t *)(acc + rot_acc_units(nH, I));
  const int tid = threadIdx.x, lane = tid & 31;
  const int qbeg = a.cta_pair[blockIdx.x], qend = a.cta_pair[blockIdx.x + 1];
  if(qbeg >= qend)
    return;
  if(tid == 0) {
    r_mbar_init(&full[0], 1);
    r_mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  { // everything the fragment loads may touch starts finite
    double *z = (double *)smem;
    const int nz = (int)(slot_bytes / sizeof(double)) + 4 * PS + 2 * rot_acc_units(nH, I);
    for(int e = tid; e < nz; e += ROT_THREADS)
      z[e] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes before the bulk copies into the slots
  }
  // ---- P0 item of this thread.  SPLIT (nMax <= 10): one direction per thread, 2 Q items over all four warps, where an
  // item is either (pn, pa >= 1) [x at m = +-pa, TE and TM] or two m = 0 entries of degrees pn, pn + 1 [TE and TM each]:
  // four complex values of x per thread either way.  Otherwise both directions per thread, item (pn, pa >= 0). ----
  constexpr int NA = NM * (NM + 1) / 2, NZ = (NM + 1) / 2, Q = NA + NZ;
  constexpr bool SPLIT = 2 * Q <= ROT_THREADS;
  const bool p0live = SPLIT ? tid < 2 * Q : tid < NH;
  const int pdir = SPLIT ? (tid >= Q ? 1 : 0) : 0; // SPLIT: the direction of this thread
  int pn = 1, pa = 0;
  bool pz = false; // SPLIT: item of two m = 0 entries
  if(p0live) {
    if(SPLIT) {
      const int u = tid - pdir * Q;
      if(u < NA) {
        pn = (int)((1.0 + sqrt(1.0 + 8.0 * u)) * 0.5);
        while(pn * (pn - 1) / 2 > u)
          --pn;
        while(pn * (pn + 1) / 2 <= u)
          ++pn;
        pa = u - pn * (pn - 1) / 2 + 1;
      } else {
        pz = true;
        pn = 2 * (u - NA) + 1;
      }
    } else {
      pn = (int)((-1.0 + sqrt(9.0 + 8.0 * tid)) * 0.5);
      while((pn - 1) * (pn + 2) / 2 > tid)
        --pn;
      while(pn * (pn + 3) / 2 <= tid)
        ++pn;
      pa = tid - (pn - 1) * (pn + 2) / 2;
    }
  }
  const bool pz2 = pz && pn + 1 <= NM; // the second degree of a two-entry item exists
  // x entries this thread reads: (fpos, fneg) = m = +pa, -pa of degree pn, or m = 0 of degrees pn, pn + 1
  const int fpos = flat_index(pn, pa), fneg = pz ? flat_index(pz2 ? pn + 1 : pn, 0) : flat_index(pn, -pa);
  const double psa = (pa & 1) ? -1.0 : 1.0, psn = (pn & 1) ? -1.0 : 1.0;
  const int pfs = rot_offF(pn) + pa, pfa = pfs + pn + 1;
  const int pfs2 = rot_offF(pz2 ? pn + 1 : pn); // F index of s_0 of the second degree
  // Consecutive P0 threads own consecutive 32-byte rows [TE | TM] of the class vectors: a quarter-warp storing the same
  // half of its rows hits every 16-byte bank group twice (ncu: 8.5 - 9 wavefronts per STS.128 where 4 is the minimum).
  // Lanes 4 .. 7 of each quarter therefore carry TM in slot 0 and TE in slot 1 from the x loads on (the flip-basis
  // arithmetic is the same for both polarisations up to the sign gm / ge): every store covers eight bank groups
  const int psw = (lane >> 2) & 1;
  // ---- fragment geometry of this lane ----
  RotLane c;
  c.warp = tid >> 5;
  c.fr = lane >> 2;
  c.fc = lane & 3;
  c.cpol = c.fc & 1;
  c.dir1 = (c.fc >> 1) != 0;
  c.gl = (c.dir1 && c.cpol) ? -1.0 : 1.0; // direction 1 carries -1 on TM (and (-1)^deg, applied per unit)
  const int bofs = (lane >> 4) * PS + ((lane >> 2) & 3);  // B fragment: plane and position in the 32-byte row (class vectors)
  const int bofsP = (lane >> 4) * PP + ((lane >> 2) & 3); // the same for the channel buffers
  const int sofs = (c.fc >> 1) * PS + 2 * c.cpol, sofsP = (c.fc >> 1) * PP + 2 * c.cpol; // D fragment column pair -> store offsets

  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  __syncthreads();
  // one record into a shared-memory slot: a single bulk copy, or three (phases | axial coefficients | small-d) when the
  // geometry sections live in the other harmonic's records; the mbarrier counts the bytes of all of them
#if ROT_SINGLE_SLOT
  // single slot: full[0] counts the small-d part + the phases of a pair (phase slots alternate: the record's own at
  // even local pair numbers, the one behind the record at odd ones), full[1] the axial part; each completes once per pair
  const uint32_t bytesD = (uint32_t)(L.rec_bytes - L.offDs), bytesC = (uint32_t)(L.offDs - L.offCp), bytesP = (uint32_t)L.offCp;
  auto fetch_geo = [&](long qq, int kpar) { // small-d sections + phases of pair qq (kpar = parity of its local number)
    const size_t o = (size_t)qq * L.rec_bytes;
    r_mbar_expect_tx(&full[0], bytesD + bytesP);
    r_bulk_g2s(smem + L.offDs, a.geo + o + L.offDs, bytesD, &full[0], pol);
    r_bulk_g2s(smem + (kpar ? L.rec_bytes : 0), a.geo + o, bytesP, &full[0], pol);
  };
  auto fetch_ax = [&](long qq) { // axial coefficients of pair qq
    r_mbar_expect_tx(&full[1], bytesC);
    r_bulk_g2s(smem + L.offCp, a.recs + (size_t)qq * L.rec_bytes + L.offCp, bytesC, &full[1], pol);
  };
  if(tid == 0) {
    fetch_geo(qbeg, 0);
    fetch_ax(qbeg);
  }
#else
  auto fetch = [&](int slot, long qq) {
    unsigned char *dst = smem + (size_t)slot * L.rec_bytes;
    const size_t o = (size_t)qq * L.rec_bytes;
    r_mbar_expect_tx(&full[slot], (uint32_t)L.rec_bytes);
    if(a.geo == a.recs)
      r_bulk_g2s(dst, a.recs + o, (uint32_t)L.rec_bytes, &full[slot], pol);
    else {
      r_bulk_g2s(dst, a.geo + o, (uint32_t)L.offCp, &full[slot], pol);
      r_bulk_g2s(dst + L.offCp, a.recs + o + L.offCp, (uint32_t)(L.offDs - L.offCp), &full[slot], pol);
      r_bulk_g2s(dst + L.offDs, a.geo + o + L.offDs, (uint32_t)(L.rec_bytes - L.offDs), &full[slot], pol);
    }
  };
  if(tid == 0)
    fetch(0, qbeg);
#endif
  // x of the pair about to be processed, at m = +pa (TE, TM) and m = -pa (TE, TM): xj = x_j (direction 0), xi = x_i
  cplx xj[4], xi[SPLIT ? 1 : 4]; // SPLIT: xj holds the thread's own direction (x_j or x_i)
  int4 pi = a.pinfo[qbeg], pnext = a.pinfo[qbeg + 1 < qend ? qbeg + 1 : qbeg];
  auto load_x = [&](cplx (&xr)[4], int part) {
    const cplx *xs = a.x + (size_t)part * n2;
    xr[0] = xs[fpos + (psw ? nH : 0)]; // slot 0 = TE, slot 1 = TM; swapped in the lanes with psw (see p0_dir)
    xr[1] = xs[fpos + (psw ? 0 : nH)];
    xr[2] = xs[fneg + (psw ? nH : 0)];
    xr[3] = xs[fneg + (psw ? 0 : nH)];
  };
  if(p0live) {
    if(SPLIT)
      load_x(xj, pdir ? pi.x : pi.y);
    else {
      load_x(xj, pi.y);
      load_x(*(cplx(*)[4])xi, pi.x);
    }
  }
  int sg = a.cta_seg[blockIdx.x];
  double keep[rot_keep_total(NM)]; // small-d A fragments of P1, reused by P3 (registers: every index is a constant)
  double *bufA = bufX, *bufB = bufY; // class vectors T / V in bufA, channels in bufB; roles swap every pair
  for(int q = qbeg; q < qend; ++q) {
    const int cur = (q - qbeg) & 1;
#if ROT_SINGLE_SLOT
    const unsigned char *rec = smem;
    const cplx *s_ph = (const cplx *)(smem + (cur ? L.rec_bytes : 0));
#else
    const unsigned char *rec = smem + (size_t)cur * L.rec_bytes;
    const cplx *s_ph = (const cplx *)rec;
#endif
    c.ph = s_ph + NM;
    c.Cp = (const double *)(rec + L.offCp) + lane;
    c.Cm = (const double *)(rec + L.offCm) - NM * NM + lane;
    c.Ds = (const double *)(rec + L.offDs) + lane;
    c.Da = (const double *)(rec + L.offDa) + lane;
    c.vb = bufA + bofs;
    c.pb = bufB + bofsP;
    c.pst = bufB + sofsP;
    c.vst = bufA + sofs;
    c.dst = acc + (c.dir1 ? 0 : RS + 4 + (pi.x % I) * RS) + c.cpol * NHP; // direction 1: column sums of particle j
#if ROT_SINGLE_SLOT
    r_mbar_wait(&full[0], (uint32_t)cur); // small-d part and phases of this pair have landed
#else
    r_mbar_wait(&full[cur], (uint32_t)(((q - qbeg) >> 1) & 1));
#endif
    // ---- P0: phases, parity signs of the reversed direction, flip basis ----
    if(p0live) {
      const cplx pp = s_ph[NM + pa], pm = s_ph[NM - pa];
      // one direction: phases, (-1)^deg and -1 on TM for direction 1 (x_i), flip basis, store
      auto p0_dir = [&](int dir, const cplx *xr) {
        const double ge0 = dir ? psn : 1.0, gm0 = dir ? -psn : 1.0;
        const double ge = psw ? gm0 : ge0, gm = psw ? ge0 : gm0; // signs of slot 0 / slot 1
        const int o1 = 1 - 2 * psw;                              // slot 0 goes to ts[0] (ts already at its half), slot 1 to ts[o1]
        cplx *ts = (cplx *)(bufA + dir * PS + 4 * pfs) + psw, *ta = (cplx *)(bufA + dir * PS + 4 * pfa) + psw;
        if(SPLIT && pz) { // exp(i 0 phi) = 1; the second degree has the opposite parity
          ts[0] = cscale(xr[0], ge);
          ts[o1] = cscale(xr[1], gm);
          if(pz2) {
            cplx *t2 = (cplx *)(bufA + dir * PS + 4 * pfs2) + psw;
            t2[0] = cscale(xr[2], dir ? -ge : ge);
            t2[o1] = cscale(xr[3], dir ? -gm : gm);
          }
        } else if(pa == 0) {
          ts[0] = cscale(xr[0], ge);
          ts[o1] = cscale(xr[1], gm);
        } else {
          const cplx te_p = cmul(pp, xr[0]), tm_p = cmul(pp, xr[1]);
          const cplx te_m = cmul(pm, xr[2]), tm_m = cmul(pm, xr[3]);
          const double fe = ge * ROT_SQH, fm = gm * ROT_SQH;
          ts[0] = mk(fe * (te_p.x + psa * te_m.x), fe * (te_p.y + psa * te_m.y));
          ts[o1] = mk(fm * (tm_p.x + psa * tm_m.x), fm * (tm_p.y + psa * tm_m.y));
          ta[0] = mk(fe * (te_p.x - psa * te_m.x), fe * (te_p.y - psa * te_m.y));
          ta[o1] = mk(fm * (tm_p.x - psa * tm_m.x), fm * (tm_p.y - psa * tm_m.y));
        }
      };
      if(SPLIT)
        p0_dir(pdir, xj);
      else {
        p0_dir(0, xj);
        p0_dir(1, xi);
      }
    }
    __syncthreads(); // B1: T complete; every thread is past P3/P4 of the previous pair -> the other record slot is free
    if(q + 1 < qend) {
#if !ROT_SINGLE_SLOT
      if(tid == 0)
        fetch(cur ^ 1, q + 1);
#endif
      if(p0live) { // next pair's x into registers (L2 hits), consumed by its P0; its (i, j) was fetched a pair ago
        if(SPLIT) {
          if(pdir)
            load_x(xj, pnext.x);
          else if(pnext.y != pi.y)
            load_x(xj, pnext.y);
        } else {
          if(pnext.y != pi.y)
            load_x(xj, pnext.y);
          load_x(*(cplx(*)[4])xi, pnext.x);
        }
      }
    }
    const int4 pnext2 = a.pinfo[q + 2 < qend ? q + 2 : qend - 1]; // (i, j) of the pair after the next one
    ROT_PER_WARP(rot_p1_warp, (std::make_integer_sequence<int, T.nd>{})) // P1: u = D^T t, channel combinations
    __syncthreads();                                                     // B2
#if ROT_SINGLE_SLOT
    // every warp has its small-d fragments in registers (P3 reuses them): the small-d part of the slot and the phase slot
    // of the previous pair are free for the next pair; P2 needs the axial part of this one
    if(tid == 0 && q + 1 < qend)
      fetch_geo(q + 1, cur ^ 1);
    r_mbar_wait(&full[1], (uint32_t)cur);
#endif
    ROT_PER_WARP(rot_p2_warp, (std::make_integer_sequence<int, T.nc>{})) // P2: q = C p, back to the class vectors
    __syncthreads();                                                     // B3
#if ROT_SINGLE_SLOT
    if(tid == 0 && q + 1 < qend) // the axial part is free
      fetch_ax(q + 1);
#endif
    ROT_PER_WARP(rot_p3_warp, (std::make_integer_sequence<int, T.nd>{})) // P3: w = D v; P4: accumulate
    if(pi.w & 3) { // last pair of the strip / of the segment: the finished sums go to HBM
      __syncthreads();
      if(pi.w & 1) {
        cplx *cpart = a.colpart + (size_t)pi.z * n2;
        for(int e = tid; e < n2; e += ROT_THREADS) {
          cplx *src = acc + (e < nH ? e : e - nH + NHP);
          cpart[e] = *src;
          *src = mk(0, 0);
        }
      }
      if(pi.w & 2) {
        cplx *rp = a.rowpart + (size_t)sg * I * n2;
        for(int e = tid; e < I * n2; e += ROT_THREADS) {
          const int r = e / n2, k = e - r * n2;
          cplx *src = acc + RS + 4 + r * RS + (k < nH ? k : k - nH + NHP);
          rp[e] = *src;
          *src = mk(0, 0);
        }
        ++sg;
      }
    }
    pi = pnext;
    pnext = pnext2;
    double *tb = bufA;
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
  p.threads = ROT_THREADS;
  if(L.nh > ROT_THREADS)
    throw Error("rotated-axial operator: nMax exceeds the kernel's thread bound");
  // rows per block: as many (<= 8) as keep rot_min_ctas CTAs per SM resident (the row sums of a block live in shared
  // memory; one column partial per strip of up to I pairs goes to HBM: C5 3.74 ms per apply with 4 rows, 3.69 with 8)
  int I = g_rot_rows > 0 ? g_rot_rows : 8;
  const size_t sm_total = (size_t)228 * 1024, cta_max = (size_t)227 * 1024;
  while(I > 1 && rot_min_ctas(NM) * (rot_smem_bytes(L, I) + 1024) > sm_total)
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
                         unsigned char *recs, RotLayout const &L, int sm_count, cudaStream_t st, bool geometry) {
  if(npairs <= 0)
    return;
  RotDTable const &dt = rot_dtable(L.NM);
  if(g_rot_assembly == 1) {
    const RotAxSizes z = rot_axial_sizes(L.NM);
    const size_t per_warp = (size_t)rot_axial_fast_entries(L.NM) * sizeof(cplx), tb = rot_axial_table_bytes(L.NM);
    const size_t budget = (size_t)224 * 1024;
    int warps, in_smem;
    long ctas;
    size_t sm;
    if(tb + 4 * per_warp <= budget) { // tables in shared memory, one persistent CTA per SM
      in_smem = 1;
      warps = (int)std::min<size_t>(ROT_AX_MAX_WARPS, (budget - tb) / per_warp);
      sm = tb + (size_t)warps * per_warp;
      ctas = std::min<long>((npairs + warps - 1) / warps, (long)sm_count);
    } else { // very large nMax: tables stay in global memory, small CTAs
      in_smem = 0;
      warps = 4;
      sm = (size_t)warps * per_warp;
      const long per_sm = std::max<long>(1, std::min<long>(16, (long)(220 * 1024) / (long)(sm + 1024)));
      ctas = std::min<long>((npairs + warps - 1) / warps, (long)sm_count * per_sm);
    }
    auto kern = in_smem ? k_assemble_axial_only<true> : k_assemble_axial_only<false>;
    OB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<(unsigned)ctas, warps * 32, sm, st>>>(xyz, k, pair_ij, npairs, recs, L, rot_axtab(L.NM), z.nrec, z.nem);
  } else {
    OB_CUDA(cudaFuncSetAttribute((const void *)k_assemble_axial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts.smem));
    OB_CUDA(cudaFuncSetAttribute((const void *)k_assemble_axial, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    k_assemble_axial<<<(unsigned)npairs, OB_VTAC_THREADS, ts.smem, st>>>(ts.tb, xyz, k, pair_ij, recs, L);
  }
  OB_CUDA(cudaGetLastError());
  if(!geometry) // the apply reads phases and small-d matrices from the other harmonic's records
    return;
  const long tctas = std::min<long>(npairs, (long)sm_count * 16);
  k_rot_tables<<<(unsigned)tctas, ROT_TAB_THREADS, 0, st>>>(xyz, pair_ij, npairs, recs, L, dt.coef, dt.seed);
  OB_CUDA(cudaGetLastError());
}

typedef void (*RotKernel)(const RotArgs);
template <int NM> static RotKernel rot_kernel_for(int nMax) {
  if(nMax == NM)
    return k_matvec_rot<NM>;
  if constexpr(NM > 1)
    return rot_kernel_for<NM - 1>(nMax);
  else
    throw Error("rotated-axial operator: nMax out of range");
}

void launch_matvec_rot(RotPlan const &p, RotLayout const &L, const unsigned char *recs, const unsigned char *geo, const cplx *x, const cplx *Tdiag,
                       cplx *acc_or_y, int finalize, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1) {
  if(e0)
    cudaEventRecord(e0, st);
  if(p.npairs > 0) {
    RotKernel kern = rot_kernel_for<OB_MAX_NMAX>(L.NM); // one instantiation per nMax: every index is an immediate
    OB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    OB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    RotArgs a;
    a.recs = recs;
    a.geo = geo ? geo : recs;
    a.x = x;
    a.pinfo = p.pinfo;
    a.cta_pair = p.cta_pair;
    a.cta_seg = p.cta_seg;
    a.rowpart = p.rowpart;
    a.colpart = p.colpart;
    a.L = L;
    a.I = p.I;
    kern<<<p.grid, p.threads, p.smem, st>>>(a);
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
