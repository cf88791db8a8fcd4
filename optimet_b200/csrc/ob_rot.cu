// ob_rot.cu -- rotated-axial form of the preconditioned coupling operator (operator = 3).
//
//   reference operator: S(block i,j) = -T_i [[A^T, B^T],[B^T, A^T]],  A,B = Coupling(R_i - R_j, k, nMax), identity
//   on the diagonal (srcAna/PreconditionedMatrix.cpp:350-400, 555-610); applied by pzgemm_ / matvec
//   (srcAna/scalapack/Belos.hpp:74-90, srcAna/PreconditionedMatrix.cpp:1058-1085).
//
// B200-first representation (not the reference's).  With R = (d, theta, phi),
//     A(R) = U A(d z) U^-1,  B(R) = U B(d z) U^-1,   U = diag(exp(i m phi)) d(theta),
// where d(theta) is block diagonal over the degree n with the real Wigner small-d matrices d^n_{m' mu}(theta)
// (Varshalovich 4.3.1) and the axial coefficients A(d z), B(d z) couple equal azimuthal orders only, with
// A(n,-mu,l,-mu) = A(n,mu,l,mu) and B(n,-mu,l,-mu) = -B(n,mu,l,mu).  The reversed direction follows from the parity
// A(-R) = (-1)^(n+l) A(R), B(-R) = (-1)^(n+l+1) B(R).  Per unordered pair only
//     (2 nMax + 1) phases + sum_mu (nMax - max(mu,1) + 1)^2 complex for A and for B + sum_n (2n+1)^2 reals
// are stored: 16.6 KB at nMax 8 (pair form 204.8 KB, dense 4 x 409.6 KB), 30 KB at nMax 10 (460.8 KB).  The identity
// and this exact data layout are checked on the CPU against the oracle's full blocks to 1e-15
// (scratch/proto_rot_layout.py, tests/test_oracle_kats.py::test_rotation_axial_factorisation).
//
// Apply, per pair and for both directions at once (four vectors: x_j TE/TM, parity-signed x_i TE/TM):
//   t = exp(i m phi) x;  u = d^T t;  v = A^T u + (+-) B^T u';  w = d v;  result = exp(-i m phi) w
// The row-side sums go to rowpart[segment], the column-side ones to colpart[pair]; k_pairs_reduce (ob_pairs.cu) adds
// them in a fixed order and applies y = x - T .* acc, exactly as for the pair form.
#include "ob_internal.h"
#include "ob_vtac.cuh"
#include "ob_rot_axial.cuh"
#include <algorithm>
#include <cuda_pipeline.h>

namespace ob {

#define ROT_THREADS 256
#define ROT_PI 3.14159265358979323846

// ---------------------------------------------------------------------------------------------
// layout helpers (host + device), mirrored by scratch/proto_rot_layout.py
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int rot_offD(int n) { // sum_{j<n} (2j+1)^2
  const int t = n - 1;
  return 4 * (t * (t + 1) * (2 * t + 1) / 6) + 4 * (t * (t + 1) / 2) + t;
}
RotLayout rot_layout(int NM) {
  RotLayout L;
  L.NM = NM;
  L.n = flat_max(NM);
  L.X = rot_offX(NM, NM + 1);
  L.Dn = rot_offD(NM + 1);
  L.offA = (size_t)(2 * NM + 1) * sizeof(cplx);
  L.offB = L.offA + (size_t)L.X * sizeof(cplx);
  L.offD = L.offB + (size_t)L.X * sizeof(cplx);
  L.rec_bytes = (L.offD + (size_t)L.Dn * sizeof(double) + 15) & ~(size_t)15;
  return L;
}

// ---------------------------------------------------------------------------------------------
// assembly 1: axial A, B of every local pair out of the shared VTAC block code (theta = phi = 0)
// ---------------------------------------------------------------------------------------------
struct EmitAxial {
  cplx *A, *B;
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
    A[e] = a;
    B[e] = b;
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
  em.A = (cplx *)(rec + L.offA);
  em.B = (cplx *)(rec + L.offB);
  em.NM = L.NM;
  vtac_block(tb, smem_raw, r, 0.0, 0.0, k, false, em);
}

// ---------------------------------------------------------------------------------------------
// assembly 1b (opt-in, "rot_assembly" = 1): axial-only recursion, one warp per pair.  With theta = 0 the scalar
// coefficients beta(n, m, l, k) vanish unless k = m and the reference's recursion
// (TranslationAdditionCoefficients.cpp:102-124) closes on those entries: O(nMax^3) per pair instead of the O(nMax^4)
// of the full block.  The per-pair body lives in ob_rot_axial.cuh and is ALSO compiled for the host: the very source the
// warp runs is checked on the CPU against the oracle's Coupling (tests/test_rot_axial_host.py, lane 0 of 1).  NOT yet
// run on a GPU: the round-1 GPU budget was spent when it was written, so it stays off by default until
// tests/test_gpu_rot.py has been run with OB_VALIDATE_PENDING=1.
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
    rot_axial_pair(L.NM, k, r, buf, (cplx *)(rec + L.offA), (cplx *)(rec + L.offB), lane, 32);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// assembly 2: phases and Wigner small-d matrices of every local pair.  One thread per (m', m): three-term recurrence
// in the degree j, seeded at j0 = max(|m'|, |m|) where the explicit sum has a single term.
// ---------------------------------------------------------------------------------------------
__constant__ double c_fact[2 * OB_MAX_NMAX + 2];

__global__ void k_rot_tables(const double *__restrict__ xyz, const int2 *__restrict__ pair_ij,
                             unsigned char *__restrict__ recs, RotLayout L) {
  const int NM = L.NM;
  const int2 ij = pair_ij[blockIdx.x];
  const double x = xyz[3 * ij.x] - xyz[3 * ij.y], y = xyz[3 * ij.x + 1] - xyz[3 * ij.y + 1],
               z = xyz[3 * ij.x + 2] - xyz[3 * ij.y + 2];
  const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
  double beta = 0.0, phi = 0.0;
  if(r > 0.0) {
    beta = acos(z / r);
    phi = atan2(y, x);
  }
  unsigned char *rec = recs + (size_t)blockIdx.x * L.rec_bytes;
  cplx *ph = (cplx *)rec;
  double *dmat = (double *)(rec + L.offD);
  const int W = 2 * NM + 1;
  for(int t = threadIdx.x; t < W; t += blockDim.x) { // ph[m + NM] = exp(i m phi)
    double s, c;
    sincos((double)(t - NM) * phi, &s, &c);
    ph[t] = mk(c, s);
  }
  double sb, cb;
  sincos(0.5 * beta, &sb, &cb);
  const double c = cos(beta);
  for(int t = threadIdx.x; t < W * W; t += blockDim.x) {
    const int mp = t / W - NM, m = t % W - NM;
    const int amp = mp < 0 ? -mp : mp, am = m < 0 ? -m : m;
    const int j0 = amp > am ? amp : am;
    // seed (Varshalovich 4.3.1 (2) at j = j0: one term, t = max(0, m - mp))
    const int tt = m - mp > 0 ? m - mp : 0;
    double seed = sqrt(c_fact[j0 + mp] * c_fact[j0 - mp] * c_fact[j0 + m] * c_fact[j0 - m]) /
                  (c_fact[j0 + m - tt] * c_fact[tt] * c_fact[mp - m + tt] * c_fact[j0 - mp - tt]);
    seed *= pow(cb, (double)(2 * j0 + m - mp - 2 * tt)) * pow(sb, (double)(mp - m + 2 * tt));
    if((mp - m + tt) & 1)
      seed = -seed;
    double dm1 = 0.0, dcur = seed;
    if(j0 >= 1)
      dmat[rot_offD(j0) + (j0 - mp) * (2 * j0 + 1) + (j0 - m)] = seed;
    for(int j = j0 + 1; j <= NM; ++j) {
      double dn;
      if(mp == 0 && m == 0)
        dn = ((2 * j - 1) * c * dcur - (j - 1) * dm1) / j;
      else
        dn = ((2 * j - 1) * ((double)(j * (j - 1)) * c - (double)(m * mp)) * dcur -
              j * sqrt((double)(((j - 1) * (j - 1) - mp * mp) * ((j - 1) * (j - 1) - m * m))) * dm1) /
             ((j - 1) * sqrt((double)((j * j - mp * mp) * (j * j - m * m))));
      dm1 = dcur;
      dcur = dn;
      dmat[rot_offD(j) + (j - mp) * (2 * j + 1) + (j - m)] = dn;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// apply
// ---------------------------------------------------------------------------------------------
struct RotArgs {
  const unsigned char *recs;
  const cplx *x;
  const int2 *pair_ij;
  const int4 *segs;
  const int *cta_seg;
  cplx *rowpart, *colpart;
  RotLayout L;
};

// dynamic shared memory: record[2] (double buffered) | bufP[4][n] | bufQ[4][n] | rowacc[2][n] | deg[n] (bytes)
static size_t rot_smem_bytes(RotLayout const &L) {
  return 2 * L.rec_bytes + (size_t)(4 + 4 + 2) * L.n * sizeof(cplx) + (((size_t)L.n + 15) & ~(size_t)15);
}
static int rot_block_threads(RotLayout const &L) { return std::min(1024, (2 * L.n + 31) & ~31); }

__device__ __forceinline__ void rot_prefetch(unsigned char *dst, const unsigned char *src, int words) {
  for(int wd = threadIdx.x; wd < words; wd += blockDim.x)
    __pipeline_memcpy_async(dst + 16 * (size_t)wd, src + 16 * (size_t)wd, 16);
  __pipeline_commit();
}

// One thread per (direction, harmonic) and phase, carrying both polarisations (every matrix element read from shared
// memory serves two or four products: the kernel is bound by shared-memory wavefronts, not by FP64 or HBM).  The CTA
// walks the contiguous range of the pair list its segments cover; the record of the next pair is copied into the
// other shared-memory buffer (cp.async) while the current one is applied.
// buffers: buf[(dir * 2 + pol) * n + e]
__global__ void __launch_bounds__(1024) k_matvec_rot(RotArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const RotLayout L = a.L;
  const int NM = L.NM, n = L.n, n2 = 2 * n, n4 = 4 * n;
  cplx *bufP = (cplx *)(smem + 2 * L.rec_bytes);
  cplx *bufQ = bufP + n4;
  cplx *rowacc = bufQ + n4;
  unsigned char *s_deg = (unsigned char *)(rowacc + 2 * n);
  const int tid = threadIdx.x, nthr = blockDim.x;
  for(int e = tid; e < n; e += nthr) {
    int nn, m;
    unflatten(e, nn, m);
    s_deg[e] = (unsigned char)nn;
  }
  const int seg0 = a.cta_seg[blockIdx.x], seg1 = a.cta_seg[blockIdx.x + 1];
  if(seg0 >= seg1)
    return;
  const int qbeg = a.segs[seg0].y, qend = a.segs[seg1 - 1].z; // the CTA's pairs are contiguous in the list
  const int words = (int)(L.rec_bytes / 16);
  rot_prefetch(smem, a.recs + (size_t)qbeg * L.rec_bytes, words);
  int sg = seg0;
  int4 seg = a.segs[sg];
  for(int e = tid; e < 2 * n; e += nthr)
    rowacc[e] = mk(0, 0);
  for(int q = qbeg; q < qend; ++q) {
    const int cur = (q - qbeg) & 1;
    unsigned char *rec = smem + (size_t)cur * L.rec_bytes;
    const cplx *s_ph = (const cplx *)rec;
    const cplx *s_A = (const cplx *)(rec + L.offA);
    const cplx *s_B = (const cplx *)(rec + L.offB);
    const double *s_d = (const double *)(rec + L.offD);
    const int i = seg.x, j = a.pair_ij[q].y;
    __pipeline_wait_prior(0);
    __syncthreads(); // record `cur` complete; the other buffer and bufP/bufQ are free (end-of-pair barrier below)
    if(q + 1 < qend)
      rot_prefetch(smem + (size_t)(cur ^ 1) * L.rec_bytes, a.recs + (size_t)(q + 1) * L.rec_bytes, words);
    // phase 0: t = exp(i m phi) x; direction 0 = x_j (row side), direction 1 = (-1)^deg x_i (column side)
    for(int it = tid; it < n2; it += nthr) {
      const int dir = it / n, e = it - dir * n;
      const int nn = s_deg[e], m = nn * (nn + 1) - e - 1;
      const cplx *xs = a.x + (size_t)(dir ? i : j) * 2 * n + e;
      cplx xe = xs[0], xm = xs[n];
      if(dir && (nn & 1)) {
        xe = cneg(xe);
        xm = cneg(xm);
      }
      const cplx phs = s_ph[m + NM];
      bufP[(dir * 2) * n + e] = cmul(phs, xe);
      bufP[(dir * 2 + 1) * n + e] = cmul(phs, xm);
    }
    __syncthreads();
    // phase 1: u[(n, mu)] = sum_{m'} d^n[m', mu] t[(n, m')], both polarisations
    for(int it = tid; it < n2; it += nthr) {
      const int dir = it / n, e = it - dir * n;
      const int nn = s_deg[e], b = e - (nn * nn - 1), w = 2 * nn + 1; // mu = nn - b
      const double *dd = s_d + rot_offD(nn) + b;
      const cplx *se = bufP + (dir * 2) * n + (nn * nn - 1), *sm = se + n;
      double ex = 0, ey = 0, mx = 0, my = 0;
#pragma unroll 3
      for(int aa = 0; aa < w; ++aa) {
        const double dv = dd[aa * w];
        const cplx te = se[aa], tm = sm[aa];
        ex = fma(dv, te.x, ex);
        ey = fma(dv, te.y, ey);
        mx = fma(dv, tm.x, mx);
        my = fma(dv, tm.y, my);
      }
      bufQ[(dir * 2) * n + e] = mk(ex, ey);
      bufQ[(dir * 2 + 1) * n + e] = mk(mx, my);
    }
    __syncthreads();
    // phase 2: v_TE = A^T u_TE + sB B^T u_TM, v_TM = sB B^T u_TE + A^T u_TM with A^T[(n,mu),(l,mu)] = A[(l,|mu|),(n,|mu|)],
    //          sB = (+1 row side | -1 column side) * sign(mu)   (B(-mu) = -B(mu))
    for(int it = tid; it < n2; it += nthr) {
      const int dir = it / n, e = it - dir * n;
      const int nn = s_deg[e], mu = nn * (nn + 1) - e - 1, am = mu < 0 ? -mu : mu;
      const int n0 = rot_n0(am), w = NM - n0 + 1;
      const int base = rot_offX(NM, am) + (nn - n0);
      const double sB = ((dir != 0) != (mu < 0)) ? -1.0 : 1.0;
      const cplx *ue = bufQ + (dir * 2) * n, *um = ue + n;
      cplx ae = mk(0, 0), amm = mk(0, 0), be = mk(0, 0), bm = mk(0, 0); // A u_TE, A u_TM, B u_TE, B u_TM
      for(int l = n0; l <= NM; ++l) {
        const int ea = base + (l - n0) * w, sl = l * (l + 1) - mu - 1;
        const cplx av = s_A[ea], bv = s_B[ea], xe = ue[sl], xm = um[sl];
        cfma(ae, av, xe);
        cfma(amm, av, xm);
        cfma(be, bv, xe);
        cfma(bm, bv, xm);
      }
      bufP[(dir * 2) * n + e] = mk(fma(sB, bm.x, ae.x), fma(sB, bm.y, ae.y));
      bufP[(dir * 2 + 1) * n + e] = mk(fma(sB, be.x, amm.x), fma(sB, be.y, amm.y));
    }
    __syncthreads();
    // phase 3 + 4: w[(n, m)] = sum_mu d^n[m, mu] v[(n, mu)];  result = exp(-i m phi) w
    for(int it = tid; it < n2; it += nthr) {
      const int dir = it / n, e = it - dir * n;
      const int nn = s_deg[e], aa = e - (nn * nn - 1), w = 2 * nn + 1, m = nn - aa;
      const double *dd = s_d + rot_offD(nn) + aa * w;
      const cplx *se = bufP + (dir * 2) * n + (nn * nn - 1), *sm = se + n;
      double ex = 0, ey = 0, mx = 0, my = 0;
#pragma unroll 3
      for(int b = 0; b < w; ++b) {
        const double dv = dd[b];
        const cplx te = se[b], tm = sm[b];
        ex = fma(dv, te.x, ex);
        ey = fma(dv, te.y, ey);
        mx = fma(dv, tm.x, mx);
        my = fma(dv, tm.y, my);
      }
      const cplx cph = cconj(s_ph[m + NM]);
      cplx ve = cmul(cph, mk(ex, ey)), vm = cmul(cph, mk(mx, my));
      if(dir == 0) { // the same thread owns elements e and n + e of rowacc for every pair
        rowacc[e] = cadd(rowacc[e], ve);
        rowacc[n + e] = cadd(rowacc[n + e], vm);
      } else {
        if(nn & 1) {
          ve = cneg(ve);
          vm = cneg(vm);
        }
        a.colpart[(size_t)q * 2 * n + e] = ve;
        a.colpart[(size_t)q * 2 * n + n + e] = vm;
      }
    }
    __syncthreads();
    if(q + 1 == seg.z) { // end of the row segment: flush the row-side sums (same element ownership as above)
      for(int it = tid; it < n; it += nthr) {
        a.rowpart[(size_t)sg * 2 * n + it] = rowacc[it];
        a.rowpart[(size_t)sg * 2 * n + n + it] = rowacc[n + it];
        rowacc[it] = mk(0, 0);
        rowacc[n + it] = mk(0, 0);
      }
      if(++sg < seg1)
        seg = a.segs[sg];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static int g_rot_assembly = 0; // 0 = vtac_block at theta = 0 (validated), 1 = axial-only recursion (see k_assemble_axial_only)
void rot_tuning(int assembly) { g_rot_assembly = assembly; }
void launch_assemble_rot(VtacTableSet const &ts, const double *xyz, cplx k, const int2 *pair_ij, long npairs,
                         unsigned char *recs, RotLayout const &L, cudaStream_t st) {
  if(npairs <= 0)
    return;
  static bool fact_set = false;
  if(!fact_set) {
    double f[2 * OB_MAX_NMAX + 2];
    f[0] = 1.0;
    for(int i = 1; i < 2 * OB_MAX_NMAX + 2; ++i)
      f[i] = f[i - 1] * (double)i;
    OB_CUDA(cudaMemcpyToSymbol(c_fact, f, sizeof(f)));
    fact_set = true;
  }
  if(g_rot_assembly == 1) {
    const size_t sm = rot_axial_smem_bytes(L.NM);
    OB_CUDA(cudaFuncSetAttribute((const void *)k_assemble_axial_only, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const long ctas = std::min<long>((npairs + ROT_AX_WARPS - 1) / ROT_AX_WARPS, 148L * 8);
    k_assemble_axial_only<<<(unsigned)ctas, ROT_AX_WARPS * 32, sm, st>>>(xyz, k, pair_ij, npairs, recs, L);
  } else {
    OB_CUDA(cudaFuncSetAttribute((const void *)k_assemble_axial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts.smem));
    OB_CUDA(cudaFuncSetAttribute((const void *)k_assemble_axial, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    k_assemble_axial<<<(unsigned)npairs, OB_VTAC_THREADS, ts.smem, st>>>(ts.tb, xyz, k, pair_ij, recs, L);
  }
  OB_CUDA(cudaGetLastError());
  k_rot_tables<<<(unsigned)npairs, 256, 0, st>>>(xyz, pair_ij, recs, L);
  OB_CUDA(cudaGetLastError());
}

void launch_matvec_rot(PairPlan const &p, RotLayout const &L, const unsigned char *recs, const cplx *x, const cplx *Tdiag,
                       cplx *acc_or_y, int finalize, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1) {
  if(e0)
    cudaEventRecord(e0, st);
  if(p.npairs > 0) {
    const size_t smem = rot_smem_bytes(L);
    OB_CUDA(cudaFuncSetAttribute((const void *)k_matvec_rot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RotArgs a;
    a.recs = recs;
    a.x = x;
    a.pair_ij = p.pair_ij;
    a.segs = p.segs;
    a.cta_seg = p.cta_seg;
    a.rowpart = p.rowpart;
    a.colpart = p.colpart;
    a.L = L;
    k_matvec_rot<<<p.grid, rot_block_threads(L), smem, st>>>(a);
    OB_CUDA(cudaGetLastError());
  }
  if(e1)
    cudaEventRecord(e1, st);
  launch_pairs_reduce(p, x, Tdiag, acc_or_y, finalize, st);
}

} // namespace ob
