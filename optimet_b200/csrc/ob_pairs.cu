// ob_pairs.cu -- compact ("pair") form of the preconditioned coupling operator and its TMA-streamed matvec.
//
//   reference operator: S(block i,j) = -T_i [[A^T, B^T],[B^T, A^T]],  A,B = Coupling(R_i - R_j, k, nMax), identity
//   on the diagonal (srcAna/PreconditionedMatrix.cpp:350-400, 555-610); applied by pzgemm_ / matvec
//   (srcAna/scalapack/Belos.hpp:74-90, srcAna/PreconditionedMatrix.cpp:1058-1085).
//
// B200-first representation (not the reference's): the dense block carries A and B twice and the row scaling T_i;
// moreover A(-R) = (-1)^(n+l) A(R), B(-R) = (-1)^(n+l+1) B(R) (tests/test_oracle_kats.py::test_inversion_parity), so
// block (j,i) is block (i,j) with a sign pattern.  Only the unscaled A^T, B^T (n x n each) of the pairs i < j are
// stored: 32 n^2 bytes per unordered pair = 1/4 of the dense bytes.  The matvec streams every stored element once
// and uses it for four products (y_i TE/TM and y_j TE/TM): 32 DFMA per 32 bytes, still below the FP64 ridge.
//
//   y_p = x_p - T_p .* ( sum_{j>p} [A B;B A]_pj x_j  +  s .* sum_{i<p} [A -B;-B A]_ip (s .* x_i) ),  s_c = (-1)^deg(c)
//
// Kernel structure: persistent CTAs (one per SM), each owning a contiguous range of the (i,j)-ordered pair list split
// into row segments; one producer warp streams column panels of A and B (one cp.async.bulk of KB*n*16 bytes each, plus
// the two 32-byte-per-column x slices) through an NS-stage shared-memory ring with mbarrier full/empty signalling;
// the consumer threads are laid out as (row r, column group g), keep the row-side sums in registers across a whole
// segment and the column-side sums across one pair, and reduce the groups through shared memory.  Row-side partials
// go to rowpart[segment], column-side partials to colpart[pair]; k_pairs_reduce adds them in a fixed order
// (deterministic: no atomics), across ranks an NCCL all-reduce adds the per-rank sums, then y = x - T .* acc.
#include "ob_internal.h"
#include <algorithm>

namespace ob {

__device__ __forceinline__ uint32_t p_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void p_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void p_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(p_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(p_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void p_mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "PWAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra PWAIT_DONE;\n"
               "bra PWAIT_LOOP;\n"
               "PWAIT_DONE:\n"
               "}" ::"r"(p_smem_u32(bar)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ uint64_t p_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t p_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void p_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(p_smem_u32(dst)), "l"(src), "r"(bytes), "r"(p_smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void consumer_barrier(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}
__device__ __forceinline__ int degree_of(int p) { // p = l(l+1) - m - 1  ->  l
  int l = (int)sqrt((double)p + 1.0);
  while(l * l > p + 1)
    --l;
  while((l + 1) * (l + 1) <= p + 1)
    ++l;
  return l;
}

#define OB_PAIR_CONSUMERS 384
#define OB_PAIR_THREADS (OB_PAIR_CONSUMERS + 32)

struct PairKernelArgs {
  const cplx *AB;       // [P_loc][2][n*n]
  const cplx *XP, *XS;  // [nobj][n][2]: (x_TE, x_TM) and s_c (x_TE, x_TM)
  const int2 *pair_ij;  // [P_loc]
  const int4 *segs;     // [nseg] {i, q0, q1, 0}
  const int *cta_seg;   // [grid + 1]
  cplx *rowpart;        // [nseg][2n]
  cplx *colpart;        // [P_loc][2n]
  int n, KB, NS, G;
};

__global__ void __launch_bounds__(OB_PAIR_THREADS, 1) k_matvec_pairs(PairKernelArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int n = a.n, KB = a.KB, NS = a.NS, G = a.G;
  const size_t stage_elems = (size_t)2 * KB * n + (size_t)4 * KB; // A panel, B panel, XP slice, XS slice
  cplx *stage_base = (cplx *)smem;
  cplx *scratch = stage_base + (size_t)NS * stage_elems;          // [2][G][2][n]
  uint64_t *full = (uint64_t *)(scratch + (size_t)2 * G * 2 * n);
  uint64_t *empty = full + NS;
  const int tid = threadIdx.x;
  if(tid == 0) {
    for(int s = 0; s < NS; ++s) {
      p_mbar_init(&full[s], 1);
      p_mbar_init(&empty[s], OB_PAIR_CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int seg0 = a.cta_seg[blockIdx.x], seg1 = a.cta_seg[blockIdx.x + 1];
  const int npanels = (n + KB - 1) / KB;

  if(tid >= OB_PAIR_CONSUMERS) {
    // ===== producer warp: one elected lane issues the bulk copies =====
    if(tid == OB_PAIR_CONSUMERS) {
      const uint64_t pol_s = p_policy_evict_first(), pol_x = p_policy_evict_last();
      int s = 0;
      uint32_t ph = 0;
      for(int sg = seg0; sg < seg1; ++sg) {
        const int4 seg = a.segs[sg];
        const cplx *xs_i = a.XS + (size_t)seg.x * n * 2;
        for(int q = seg.y; q < seg.z; ++q) {
          const int j = a.pair_ij[q].y;
          const cplx *Aq = a.AB + (size_t)q * 2 * n * n;
          const cplx *xp_j = a.XP + (size_t)j * n * 2;
          for(int pn = 0; pn < npanels; ++pn) {
            const int c0 = pn * KB, nc = min(KB, n - c0);
            p_mbar_wait(&empty[s], ph ^ 1);
            cplx *dst = stage_base + (size_t)s * stage_elems;
            const uint32_t mb = (uint32_t)((size_t)nc * n * sizeof(cplx)), xb = (uint32_t)(nc * 2 * sizeof(cplx));
            p_mbar_expect_tx(&full[s], 2 * mb + 2 * xb);
            p_bulk_g2s(dst, Aq + (size_t)c0 * n, mb, &full[s], pol_s);
            p_bulk_g2s(dst + (size_t)KB * n, Aq + (size_t)n * n + (size_t)c0 * n, mb, &full[s], pol_s);
            p_bulk_g2s(dst + (size_t)2 * KB * n, xp_j + (size_t)c0 * 2, xb, &full[s], pol_x);
            p_bulk_g2s(dst + (size_t)2 * KB * n + 2 * KB, xs_i + (size_t)c0 * 2, xb, &full[s], pol_x);
            if(++s == NS) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
    return;
  }

  // ===== consumers: thread -> (row r, column group g) =====
  const int r = tid % n, g = tid / n;
  const bool active = g < G;
  const int n2 = 2 * n; // reduction role: thread o (o < 2n, strided) sums output (h, r) = (o / n, o % n) over the groups
  int s = 0, ev = 0;
  uint32_t ph = 0;
  for(int sg = seg0; sg < seg1; ++sg) {
    const int4 seg = a.segs[sg];
    double rTEx = 0, rTEy = 0, rTMx = 0, rTMy = 0; // row-side sums (y_i), kept across the segment
    for(int q = seg.y; q < seg.z; ++q) {
      double cTEx = 0, cTEy = 0, cTMx = 0, cTMy = 0; // column-side sums (y_j), one pair
      for(int pn = 0; pn < npanels; ++pn) {
        const int nc = min(KB, n - pn * KB);
        p_mbar_wait(&full[s], ph);
        if(active) {
          const cplx *As = stage_base + (size_t)s * stage_elems;
          const cplx *Bs = As + (size_t)KB * n;
          const cplx *xp = As + (size_t)2 * KB * n;
          const cplx *xs = xp + 2 * KB;
#pragma unroll 2
          for(int c = g; c < nc; c += G) {
            const cplx av = As[(size_t)c * n + r], bv = Bs[(size_t)c * n + r];
            const cplx xje = xp[2 * c], xjm = xp[2 * c + 1], xie = xs[2 * c], xim = xs[2 * c + 1];
            // y_i.TE += A xj.TE + B xj.TM ; y_i.TM += B xj.TE + A xj.TM
            rTEx = fma(av.x, xje.x, rTEx);
            rTEy = fma(av.x, xje.y, rTEy);
            rTMx = fma(av.x, xjm.x, rTMx);
            rTMy = fma(av.x, xjm.y, rTMy);
            rTEx = fma(-av.y, xje.y, rTEx);
            rTEy = fma(av.y, xje.x, rTEy);
            rTMx = fma(-av.y, xjm.y, rTMx);
            rTMy = fma(av.y, xjm.x, rTMy);
            rTEx = fma(bv.x, xjm.x, rTEx);
            rTEy = fma(bv.x, xjm.y, rTEy);
            rTMx = fma(bv.x, xje.x, rTMx);
            rTMy = fma(bv.x, xje.y, rTMy);
            rTEx = fma(-bv.y, xjm.y, rTEx);
            rTEy = fma(bv.y, xjm.x, rTEy);
            rTMx = fma(-bv.y, xje.y, rTMx);
            rTMy = fma(bv.y, xje.x, rTMy);
            // y_j.TE += A (s xi.TE) - B (s xi.TM) ; y_j.TM += A (s xi.TM) - B (s xi.TE)
            cTEx = fma(av.x, xie.x, cTEx);
            cTEy = fma(av.x, xie.y, cTEy);
            cTMx = fma(av.x, xim.x, cTMx);
            cTMy = fma(av.x, xim.y, cTMy);
            cTEx = fma(-av.y, xie.y, cTEx);
            cTEy = fma(av.y, xie.x, cTEy);
            cTMx = fma(-av.y, xim.y, cTMx);
            cTMy = fma(av.y, xim.x, cTMy);
            cTEx = fma(-bv.x, xim.x, cTEx);
            cTEy = fma(-bv.x, xim.y, cTEy);
            cTMx = fma(-bv.x, xie.x, cTMx);
            cTMy = fma(-bv.x, xie.y, cTMy);
            cTEx = fma(bv.y, xim.y, cTEx);
            cTEy = fma(-bv.y, xim.x, cTEy);
            cTMx = fma(bv.y, xie.y, cTMx);
            cTMy = fma(-bv.y, xie.x, cTMy);
          }
        }
        __syncwarp();
        if((tid & 31) == 0)
          p_mbar_arrive(&empty[s]);
        if(++s == NS) {
          s = 0;
          ph ^= 1;
        }
      }
      // ---- column-side reduction over the groups -> colpart[q] (sign s_r applied) ----
      {
        cplx *buf = scratch + (size_t)(ev & 1) * G * 2 * n;
        if(active) {
          buf[(size_t)(g * 2 + 0) * n + r] = mk(cTEx, cTEy);
          buf[(size_t)(g * 2 + 1) * n + r] = mk(cTMx, cTMy);
        }
        consumer_barrier(OB_PAIR_CONSUMERS);
        for(int o = tid; o < n2; o += OB_PAIR_CONSUMERS) {
          const int out_h = o / n, out_r = o - out_h * n;
          cplx sum = buf[(size_t)out_h * n + out_r];
          for(int gg = 1; gg < G; ++gg)
            sum = cadd(sum, buf[(size_t)(gg * 2 + out_h) * n + out_r]);
          const double out_sign = (degree_of(out_r) & 1) ? -1.0 : 1.0;
          a.colpart[(size_t)q * n2 + o] = mk(out_sign * sum.x, out_sign * sum.y);
        }
        ++ev;
      }
    }
    // ---- row-side reduction -> rowpart[segment] ----
    {
      cplx *buf = scratch + (size_t)(ev & 1) * G * 2 * n;
      if(active) {
        buf[(size_t)(g * 2 + 0) * n + r] = mk(rTEx, rTEy);
        buf[(size_t)(g * 2 + 1) * n + r] = mk(rTMx, rTMy);
      }
      consumer_barrier(OB_PAIR_CONSUMERS);
      for(int o = tid; o < n2; o += OB_PAIR_CONSUMERS) {
        const int out_h = o / n, out_r = o - out_h * n;
        cplx sum = buf[(size_t)out_h * n + out_r];
        for(int gg = 1; gg < G; ++gg)
          sum = cadd(sum, buf[(size_t)(gg * 2 + out_h) * n + out_r]);
        a.rowpart[(size_t)sg * n2 + o] = sum;
      }
      ++ev;
    }
  }
}

// XP[p][c] = (x_TE[c], x_TM[c]);  XS[p][c] = s_c (x_TE[c], x_TM[c])
__global__ void k_pairs_prepare_x(const cplx *__restrict__ x, int nobj, int n, cplx *__restrict__ XP,
                                  cplx *__restrict__ XS) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(e >= (size_t)nobj * n)
    return;
  const int p = (int)(e / n), c = (int)(e - (size_t)p * n);
  const cplx te = x[(size_t)p * 2 * n + c], tm = x[(size_t)p * 2 * n + n + c];
  const double sg = (degree_of(c) & 1) ? -1.0 : 1.0;
  XP[2 * e] = te;
  XP[2 * e + 1] = tm;
  XS[2 * e] = mk(sg * te.x, sg * te.y);
  XS[2 * e + 1] = mk(sg * tm.x, sg * tm.y);
}

// acc_p = sum of the row-side partials of the segments of row p (segment order) + the column-side partials of the
// local pairs (i, p), i ascending.  finalize != 0: y_p = x_p - T_p .* acc_p written directly (single rank).
#define OB_REDUCE_THREADS 1024
__global__ void __launch_bounds__(OB_REDUCE_THREADS)
k_pairs_reduce(const cplx *__restrict__ rowpart, const cplx *__restrict__ colpart, const int *__restrict__ row_seg,
               const int2 *__restrict__ col_range, long p0, int nobj, int n,
               const cplx *__restrict__ x, const cplx *__restrict__ Tdiag, cplx *__restrict__ out, int finalize) {
  // One CTA per particle p.  The up-to-(nobj-1) column-side partials of p are split over `parts` thread groups,
  // each thread keeps 8 independent running sums (the loads are latency-bound: a serial chain of 199 L2 reads cost
  // 36 us per apply before); everything is combined in a fixed order (slots pairwise, then part 0..parts-1).
  __shared__ cplx sh[OB_REDUCE_THREADS];
  const int p = blockIdx.x;
  const int n2 = 2 * n;
  const int parts = max(1, (int)blockDim.x / n2);
  const int part = threadIdx.x / n2, e0 = threadIdx.x - part * n2;
  const int2 cr = col_range[p]; // rows i in [cr.x, cr.y) have a local pair (i, p)
  const int s0 = row_seg[p], s1 = row_seg[p + 1];
  const int nterms = (s1 - s0) + (cr.y - cr.x);
  for(int eb = 0; eb < n2; eb += (int)blockDim.x) { // n2 <= blockDim.x in practice: one trip
    const int e = n2 <= (int)blockDim.x ? e0 : eb + (int)threadIdx.x;
    const bool live = e < n2 && (n2 > (int)blockDim.x || part < parts);
    cplx a0 = mk(0, 0), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
    if(live) {
      const int stride = n2 <= (int)blockDim.x ? parts : 1, first = n2 <= (int)blockDim.x ? part : 0;
      int k = first;
      auto term = [&](int t) -> cplx { // t-th term in the fixed global order: row segments, then i ascending
        if(t < s1 - s0)
          return rowpart[(size_t)(s0 + t) * n2 + e];
        const int i = cr.x + (t - (s1 - s0));
        // local index of pair (i, p): its global index i nobj - i (i + 1) / 2 + (p - i - 1) minus the rank's p0
        const long q = (long)i * nobj - (long)i * (i + 1) / 2 + (p - i - 1) - p0;
        return colpart[(size_t)q * n2 + e];
      };
      for(; k + 7 * stride < nterms; k += 8 * stride) {
        const cplx v0 = term(k), v1 = term(k + stride), v2 = term(k + 2 * stride), v3 = term(k + 3 * stride);
        const cplx v4 = term(k + 4 * stride), v5 = term(k + 5 * stride), v6 = term(k + 6 * stride),
                   v7 = term(k + 7 * stride);
        a0 = cadd(a0, v0);
        a1 = cadd(a1, v1);
        a2 = cadd(a2, v2);
        a3 = cadd(a3, v3);
        a4 = cadd(a4, v4);
        a5 = cadd(a5, v5);
        a6 = cadd(a6, v6);
        a7 = cadd(a7, v7);
      }
      for(; k < nterms; k += stride)
        a0 = cadd(a0, term(k));
    }
    cplx s = cadd(cadd(cadd(a0, a1), cadd(a2, a3)), cadd(cadd(a4, a5), cadd(a6, a7)));
    if(n2 <= (int)blockDim.x) {
      sh[threadIdx.x] = s;
      __syncthreads();
      if(part == 0 && e < n2)
        for(int q = 1; q < parts; ++q)
          s = cadd(s, sh[q * n2 + e]);
      __syncthreads();
      if(part != 0)
        continue;
    }
    if(e < n2) {
      const size_t o = (size_t)p * n2 + e;
      if(finalize)
        out[o] = csub(x[o], cmul(Tdiag[o], s));
      else
        out[o] = s;
    }
  }
}

// y = x - T .* acc  (after the cross-rank sum of acc)
__global__ void k_pairs_finalize(const cplx *__restrict__ x, const cplx *__restrict__ Tdiag, const cplx *__restrict__ acc,
                                 size_t N, cplx *__restrict__ y) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(e < N)
    y[e] = csub(x[e], cmul(Tdiag[e], acc[e]));
}

// dense reference block (i, j), 2n x 2n column-major, rebuilt from the pair storage (tests / ob_fetch_block)
__global__ void k_pairs_expand_block(const cplx *__restrict__ ABq, int n, int swapped, int diagonal,
                                     const cplx *__restrict__ Ti, cplx *__restrict__ out) {
  const int n2 = 2 * n;
  for(int e = blockIdx.x * blockDim.x + threadIdx.x; e < n2 * n2; e += gridDim.x * blockDim.x) {
    const int c = e / n2, r = e - c * n2;
    if(diagonal) {
      out[e] = mk(r == c ? 1.0 : 0.0, 0.0);
      continue;
    }
    const int rh = r / n, rr = r - rh * n, ch = c / n, cc = c - ch * n;
    const bool isB = rh != ch;
    cplx v = ABq[(size_t)(isB ? n * n : 0) + (size_t)cc * n + rr];
    if(swapped) { // block (j, i) from pair (i, j): s_r s_c A, -s_r s_c B
      const int par = (degree_of(rr) + degree_of(cc) + (isB ? 1 : 0)) & 1;
      if(par)
        v = cneg(v);
    }
    out[e] = cneg(cmul(Ti[r], v));
  }
}

// ---------------------------------------------------------------------------------------------
// host: plan
// ---------------------------------------------------------------------------------------------
static long pair_index(long nobj, long i, long j) { return i * nobj - i * (i + 1) / 2 + (j - i - 1); }

template <class T> static T *upload_vec(std::vector<T> const &v) {
  T *d = nullptr;
  if(v.empty())
    return d;
  OB_CUDA(cudaMalloc(&d, v.size() * sizeof(T)));
  OB_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

void pair_plan_release(PairPlan &p) {
  void *ptrs[] = {p.pair_ij, p.segs, p.cta_seg, p.row_seg, p.col_range, p.col_first, p.rowpart, p.colpart, p.XP, p.XS,
                  p.acc};
  for(void *q : ptrs)
    if(q)
      cudaFree(q);
  p = PairPlan();
}

static int g_pairs_kb = 0, g_pairs_groups = 0; // tuning overrides (ob_set_option "pairs_kb" / "pairs_groups"); 0 = auto
void pair_plan_tuning(int kb, int groups) {
  g_pairs_kb = kb;
  g_pairs_groups = groups;
}

void pair_plan_build(PairPlan &p, int nobj, int n, int world, int rank, int sm_count) {
  pair_plan_release(p);
  p.nobj = nobj;
  p.n = n;
  const long P = (long)nobj * (nobj - 1) / 2;
  p.p0 = P * rank / world;
  p.p1 = P * (rank + 1) / world;
  const long Ploc = p.p1 - p.p0;
  p.npairs = Ploc;
  // pipeline geometry: KB columns per stage, as many stages as fit next to the reduction scratch
  // column groups G: as many (row, group) lanes as the consumer threads allow; KB = 4 columns per thread and stage
  // (measured on C4, n = 80: KB 4/8/12/16 -> 4.58/6.14/6.70/6.85 TB/s: short stages are pipeline-sync bound),
  // shrunk until at least 4 stages fit in shared memory
  p.G = std::max(1, std::min(OB_PAIR_CONSUMERS / n, 8));
  if(g_pairs_groups > 0)
    p.G = std::max(1, std::min(g_pairs_groups, p.G));
  p.KB = 4 * p.G;
  if(g_pairs_kb > 0)
    p.KB = g_pairs_kb;
  p.KB = std::min(p.KB, n);
  p.G = std::min(p.G, p.KB);
  while(p.KB > p.G && 4 * ((size_t)2 * p.KB * n + 4 * p.KB) * sizeof(cplx) + (size_t)2 * p.G * 2 * n * sizeof(cplx) >
                          (size_t)215 * 1024)
    p.KB -= p.G;
  const size_t stage = ((size_t)2 * p.KB * n + 4 * p.KB) * sizeof(cplx);
  const size_t scratch = (size_t)2 * p.G * 2 * n * sizeof(cplx);
  const size_t budget = 220 * 1024;
  int ns = (int)((budget - scratch - 256) / stage);
  p.NS = std::max(2, std::min(ns, 12));
  p.smem = p.NS * stage + scratch + 2 * p.NS * sizeof(uint64_t) + 16;
  if(n > OB_PAIR_CONSUMERS)
    throw Error("pair operator: n exceeds the consumer count");
  // local pair list (i, j), ordered by i then j
  std::vector<int2> ij((size_t)Ploc);
  std::vector<long> col_first(nobj, 0); // local index of pair (i, i+1) if it were local (may be negative / out of range)
  {
    long q = 0;
    for(long i = 0; i < nobj - 1; ++i) {
      const long g0 = pair_index(nobj, i, i + 1);
      col_first[i] = g0 - p.p0;
      for(long j = i + 1; j < nobj; ++j) {
        const long gidx = g0 + (j - i - 1);
        if(gidx >= p.p0 && gidx < p.p1)
          ij[(size_t)(q++)] = make_int2((int)i, (int)j);
      }
    }
  }
  // CTA ranges (equal pair counts) split into row segments
  p.grid = (int)std::max<long>(1, std::min<long>(sm_count, Ploc));
  std::vector<int4> segs;
  std::vector<int> cta_seg(p.grid + 1, 0);
  for(int b = 0; b < p.grid; ++b) {
    const long s0 = Ploc * b / p.grid, s1 = Ploc * (b + 1) / p.grid;
    cta_seg[b] = (int)segs.size();
    long q = s0;
    while(q < s1) {
      const int i = ij[(size_t)q].x;
      long e = q;
      while(e < s1 && ij[(size_t)e].x == i)
        ++e;
      segs.push_back(make_int4(i, (int)q, (int)e, 0));
      q = e;
    }
  }
  cta_seg[p.grid] = (int)segs.size();
  p.nseg = (int)segs.size();
  std::vector<int> row_seg(nobj + 1, 0);
  {
    size_t k = 0;
    for(int i = 0; i < nobj; ++i) {
      row_seg[i] = (int)k;
      while(k < segs.size() && segs[k].x == i)
        ++k;
    }
    row_seg[nobj] = (int)segs.size();
  }
  std::vector<int2> col_range(nobj);
  for(long pcol = 0; pcol < nobj; ++pcol) {
    // pairs (i, pcol), i < pcol: global index increases with i -> the local ones form a contiguous i range
    int lo = (int)pcol, hi = (int)pcol;
    bool any = false;
    for(long i = 0; i < pcol; ++i) {
      const long gidx = pair_index(nobj, i, pcol);
      if(gidx >= p.p0 && gidx < p.p1) {
        if(!any) {
          lo = (int)i;
          any = true;
        }
        hi = (int)i + 1;
      }
    }
    col_range[pcol] = any ? make_int2(lo, hi) : make_int2(0, 0);
  }
  p.pair_ij = upload_vec(ij);
  p.segs = upload_vec(segs);
  p.cta_seg = upload_vec(cta_seg);
  p.row_seg = upload_vec(row_seg);
  p.col_range = upload_vec(col_range);
  p.col_first = upload_vec(col_first);
  const size_t n2 = 2 * (size_t)n;
  OB_CUDA(cudaMalloc(&p.rowpart, std::max<size_t>(1, segs.size()) * n2 * sizeof(cplx)));
  OB_CUDA(cudaMalloc(&p.colpart, std::max<size_t>(1, (size_t)Ploc) * n2 * sizeof(cplx)));
  OB_CUDA(cudaMalloc(&p.XP, (size_t)nobj * n2 * sizeof(cplx)));
  OB_CUDA(cudaMalloc(&p.XS, (size_t)nobj * n2 * sizeof(cplx)));
  OB_CUDA(cudaMalloc(&p.acc, (size_t)nobj * n2 * sizeof(cplx)));
}

size_t pair_storage_elems(PairPlan const &p) { return (size_t)p.npairs * 2 * p.n * p.n; }

// streams the local pairs once; afterwards `acc_or_y` holds either y (finalize) or this rank's partial sums
void launch_matvec_pairs(PairPlan const &p, const cplx *AB, const cplx *x, const cplx *Tdiag, cplx *acc_or_y,
                         int finalize, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, bool x_staged) {
  const int n = p.n;
  if(!x_staged) { // XP / XS already written by the fused Arnoldi step that produced x (ob_vec.cu)
    const size_t tot = (size_t)p.nobj * n;
    k_pairs_prepare_x<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, p.nobj, n, p.XP, p.XS);
    OB_CUDA(cudaGetLastError());
  }
  if(e0)
    cudaEventRecord(e0, st);
  if(p.npairs > 0) {
    static size_t attr_smem = 0;
    if(p.smem > attr_smem) {
      OB_CUDA(cudaFuncSetAttribute((const void *)k_matvec_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)p.smem));
      attr_smem = p.smem;
    }
    PairKernelArgs a;
    a.AB = AB;
    a.XP = p.XP;
    a.XS = p.XS;
    a.pair_ij = p.pair_ij;
    a.segs = p.segs;
    a.cta_seg = p.cta_seg;
    a.rowpart = p.rowpart;
    a.colpart = p.colpart;
    a.n = n;
    a.KB = p.KB;
    a.NS = p.NS;
    a.G = p.G;
    k_matvec_pairs<<<p.grid, OB_PAIR_THREADS, p.smem, st>>>(a);
    OB_CUDA(cudaGetLastError());
  }
  if(e1)
    cudaEventRecord(e1, st);
  launch_pairs_reduce(p, x, Tdiag, acc_or_y, finalize, st);
}

// fixed-order sum of rowpart / colpart (shared with the rotated-axial form, ob_rot.cu)
void launch_pairs_reduce(PairPlan const &p, const cplx *x, const cplx *Tdiag, cplx *acc_or_y, int finalize,
                         cudaStream_t st) {
  k_pairs_reduce<<<p.nobj, OB_REDUCE_THREADS, 0, st>>>(p.rowpart, p.colpart, p.row_seg, p.col_range, p.p0, p.nobj, p.n, x,
                                                        Tdiag, acc_or_y, finalize);
  OB_CUDA(cudaGetLastError());
}

void launch_pairs_finalize(const cplx *x, const cplx *Tdiag, const cplx *acc, size_t N, cplx *y, cudaStream_t st) {
  k_pairs_finalize<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(x, Tdiag, acc, N, y);
  OB_CUDA(cudaGetLastError());
}

void launch_pairs_expand_block(PairPlan const &p, const cplx *AB, int i, int j, const cplx *Tdiag, cplx *out,
                               cudaStream_t st) {
  const int n = p.n;
  const cplx *Ti = Tdiag + (size_t)i * 2 * n;
  if(i == j) {
    k_pairs_expand_block<<<64, 256, 0, st>>>(nullptr, n, 0, 1, Ti, out);
  } else {
    const long a = std::min(i, j), b = std::max(i, j);
    const long gidx = pair_index(p.nobj, a, b);
    if(gidx < p.p0 || gidx >= p.p1)
      throw Error("block is not local to this rank");
    k_pairs_expand_block<<<64, 256, 0, st>>>(AB + (size_t)(gidx - p.p0) * 2 * n * n, n, i > j ? 1 : 0, 0, Ti, out);
  }
  OB_CUDA(cudaGetLastError());
}

} // namespace ob
