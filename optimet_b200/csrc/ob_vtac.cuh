// ob_vtac.cuh -- vector translation-addition coefficients (VTAC) for one displacement, one CTA.
//
// Replaces, for one particle pair, the reference chain
//   TranslationAdditionCoefficients (srcAna/TranslationAdditionCoefficients.cpp:71-135; memoised
//   std::map recursion, one AMOS + one boost Y_nm call per seed)  ->
//   Coupling::coefficients_A / coefficients_B (srcAna/Coupling.cpp:30-51).
//
// B200-first formulation (not the reference's):
//   * the scalar coefficients are factored as beta(n,m,l,k) = exp(i(m-k)phi) * R(n,m,l,k); R obeys the
//     same real-coefficient recurrences (Stout 2002 App. C) with phi = 0 seeds
//         R(0,0,l,k) = sqrt(4pi) * s * N_l^|k|(cos theta) * z_l(kr),  s = (-1)^l (k>=0) or (-1)^(l+k) (k<0)
//     and the reference's second ("negative") recursion instance collapses to the symmetry
//         beta(n,-m,l,k) = (-1)^(m+k) exp(i(-m-k)phi) R(n,m,l,-k)
//     (checked against the oracle's literal two-instance restatement to 1e-15);
//   * the recursion is marched level by level in n for all chains m <= n at once (thousands of
//     independent entries per step), ping-ponging two shared-memory buffers with in-place update of
//     the n-2 level; A/B columns of level n are emitted while level n+1 is computed (one barrier per level);
//   * one radial sequence z_l(kr), l = 0..2 nMax, and one Legendre table per CTA, in shared memory.
#pragma once
#include "ob_special.cuh"

namespace ob {

struct VtacTables {
  int NM, L, T;                  // nMax, 2 nMax, complex entries per level buffer
  const double *ap, *am, *bp, *bm;     // a+,a-,b+,b- (TranslationAdditionCoefficients.cpp:35-61), [(L+2)^2] at l(l+1)+k
  const double *inv_ap_nm, *am_nm;     // 1/a+(n-1,m), a-(n-1,m) at n*(NM+1)+m
  const double *inv_bp_n;              // 1/b+(n-1,n-1) at n
  const unsigned char *lamOf;          // l of triangular index idx = l(l+1)+k, idx < (L+1)^2
  const int *off;                      // chain offsets, [NM+2]
  const double *rowc;                  // per row r=(l,k): fA, fB, t1, t2, u0, u1, u2, pad   (Coupling.cpp:33-37, 43-48)
  const double *colc;                  // per column p=(n,mu): gn, s1, s2, pad
};

__host__ __device__ inline int vtac_buffer_entries(int NM) {
  int T = 0;
  for(int m = 0; m <= NM; ++m)
    T += (2 * NM - m + 1) * (2 * NM - m + 1);
  return T;
}
// dynamic shared memory needed by vtac_block (bytes)
__host__ __device__ inline size_t vtac_smem_bytes(int NM) {
  int L = 2 * NM;
  size_t b = 2 * (size_t)vtac_buffer_entries(NM) * sizeof(cplx); // two level buffers
  b += (size_t)(L + 1) * sizeof(cplx);                          // radial z_l
  b += (size_t)(4 * NM + 1) * sizeof(cplx);                     // phase table
  b += (size_t)((L + 1) * (L + 2) / 2) * sizeof(double);        // Legendre
  return b;
}

struct VtacSmem {
  cplx *buf[2];
  cplx *zl;
  cplx *ph;
  double *nlm;
  __device__ VtacSmem(unsigned char *base, int NM) {
    int T = vtac_buffer_entries(NM), L = 2 * NM;
    buf[0] = (cplx *)base;
    buf[1] = buf[0] + T;
    zl = buf[1] + T;
    ph = zl + (L + 1);
    nlm = (double *)(ph + (4 * NM + 1));
  }
};

// phase-free coefficient R(n, mu, lam, kap) of the current level (buffer G), incl. the mu<0 symmetry
__device__ __forceinline__ cplx vtac_beta(const cplx *G, const int *off, int n, int mu, int lam, int kap) {
  int ak = kap < 0 ? -kap : kap;
  int am = mu < 0 ? -mu : mu;
  if(lam < 0 || ak > lam || am > n)
    return mk(0, 0);
  if(mu >= 0)
    return G[off[mu] + lam * (lam + 1) + kap];
  cplx v = G[off[am] + lam * (lam + 1) - kap];
  return ((mu + kap) & 1) ? cneg(v) : v;
}

// Runs the whole VTAC computation for displacement (r, theta, phi) and wavenumber k on one CTA and
// calls emit(p, r, A, B) for every column p = flat(n, mu) and row r = flat(l, k):
//   A = Coupling.diagonal(p, r), B = Coupling.offdiagonal(p, r)      (Coupling.cpp:53-76).
// Thread t owns row (t % gs) for the whole call (gs = rows rounded up to a warp multiple) and the
// column group t / gs; emit.finish() is called once at the end by every thread.
// `regular` = true selects j_l (incident/origin translations), false selects h1_l (particle coupling).
template <class Emit>
__device__ void vtac_block(VtacTables const &tb, unsigned char *smem_raw, double r, double the, double phi, cplx k,
                           bool regular, Emit &emit) {
  const int NM = tb.NM, L = tb.L;
  const int tid = threadIdx.x, nthr = blockDim.x;
  VtacSmem sm(smem_raw, NM);
  const int nrows = flat_max(NM);

  // ---- seeds: radial sequence (one thread), Legendre (one thread per order m), phases ----
  if(tid == 0) {
    cplx z = cscale(k, r);
    if(regular)
      sph_bessel_j(z, L, sm.zl);
    else
      sph_hankel1(z, L, sm.zl);
  }
  if(tid >= 32 && tid < 32 + L + 1) {
    int m = tid - 32;
    double x = cos(the), s = sin(the);
    if(s < 0)
      s = -s;
    // out[l] for l=m..L lives at nlm[l(l+1)/2 + m]: strided writes through a small adaptor
    double pmm = 1.0; // sqrt(4 pi) folded in: sqrt(4pi) * sqrt(1/(4pi))
    for(int i = 1; i <= m; ++i)
      pmm *= -sqrt((double)(2 * i + 1) / (double)(2 * i)) * s;
    sm.nlm[m * (m + 1) / 2 + m] = pmm;
    if(L > m) {
      double pmmp1 = x * sqrt((double)(2 * m + 3)) * pmm;
      sm.nlm[(m + 1) * (m + 2) / 2 + m] = pmmp1;
      for(int l = m + 2; l <= L; ++l) {
        double a = sqrt((double)(4 * l * l - 1) / (double)(l * l - m * m));
        double b = sqrt((double)((l - 1) * (l - 1) - m * m) / (double)(4 * (l - 1) * (l - 1) - 1));
        double pll = a * (x * pmmp1 - b * pmm);
        pmm = pmmp1;
        pmmp1 = pll;
        sm.nlm[l * (l + 1) / 2 + m] = pll;
      }
    }
  }
  if(tid >= 96 && tid < 96 + 4 * NM + 1) {
    int d = tid - 96 - 2 * NM;
    double s, c;
    sincos((double)d * phi, &s, &c);
    sm.ph[tid - 96] = mk(c, s);
  }
  __syncthreads();
  {
    cplx *G = sm.buf[0];
    const int sz = (L + 1) * (L + 1);
    for(int idx = tid; idx < sz; idx += nthr) {
      int lam = tb.lamOf[idx];
      int kap = idx - lam * (lam + 1);
      int ak = kap < 0 ? -kap : kap;
      int sg = kap >= 0 ? lam : lam + kap;
      double v = sm.nlm[lam * (lam + 1) / 2 + ak];
      if(sg & 1)
        v = -v;
      G[idx] = cscale(sm.zl[lam], v);
    }
  }
  __syncthreads();

  // ---- per-thread row constants for the emission ----
  const int gs = (nrows + 31) & ~31;
  const int ngroups = nthr / gs;
  const int row = tid % gs;
  const int grp = tid / gs;
  const bool row_active = row < nrows && grp < ngroups;
  int rl = 0, rk = 0;
  double fA = 0, fB = 0, t1 = 0, t2 = 0, u0 = 0, u1 = 0, u2 = 0;
  if(row_active) {
    unflatten(row, rl, rk);
    const double *rc = tb.rowc + 8 * row;
    fA = rc[0];
    fB = rc[1];
    t1 = rc[2];
    t2 = rc[3];
    u0 = rc[4];
    u1 = rc[5];
    u2 = rc[6];
  }

  auto compute_level = [&](int n) {
    const int Ln = L - n;
    const int sz = (Ln + 1) * (Ln + 1);
    const int items = (n + 1) * sz;
    cplx *dstb = sm.buf[n & 1];
    const cplx *src = sm.buf[(n - 1) & 1];
    for(int it = tid; it < items; it += nthr) {
      int m = it / sz;
      int idx = it - m * sz;
      int lam = tb.lamOf[idx];
      int kap = idx - lam * (lam + 1);
      cplx v = mk(0, 0);
      int up = (lam + 1) * (lam + 2) + kap, dn = (lam - 1) * lam + kap;
      if(m == n) { // sectorial step (TranslationAdditionCoefficients.cpp:113-117)
        const cplx *s = src + tb.off[n - 1];
        int k1 = kap - 1;
        if(lam >= 1 && (k1 < 0 ? -k1 : k1) <= lam - 1)
          v = cscale(s[dn - 1], __ldg(tb.bp + dn - 1));
        cplx w = s[up - 1];
        double c = __ldg(tb.bm + up - 1);
        v.x = fma(w.x, c, v.x);
        v.y = fma(w.y, c, v.y);
        v = cscale(v, __ldg(tb.inv_bp_n + n));
      } else { // general step (:119-124)
        const cplx *s = src + tb.off[m];
        if((kap < 0 ? -kap : kap) <= lam - 1)
          v = cscale(s[dn], __ldg(tb.ap + dn));
        cplx w = s[up];
        double c = __ldg(tb.am + up);
        v.x = fma(w.x, c, v.x);
        v.y = fma(w.y, c, v.y);
        if(n - 2 >= m) {
          cplx o = dstb[tb.off[m] + idx];
          double a = __ldg(tb.am_nm + n * (NM + 1) + m);
          v.x = fma(-o.x, a, v.x);
          v.y = fma(-o.y, a, v.y);
        }
        v = cscale(v, __ldg(tb.inv_ap_nm + n * (NM + 1) + m));
      }
      dstb[tb.off[m] + idx] = v;
    }
  };

  auto emit_level = [&](int n) {
    if(!row_active)
      return;
    const cplx *G = sm.buf[n & 1];
    const int l = rl, kk = rk;
    for(int mu = n - grp; mu >= -n; mu -= ngroups) {
      const int p = flat_index(n, mu);
      const double *cc = tb.colc + 4 * p;
      const double gn = __ldg(cc), s1 = __ldg(cc + 1), s2 = __ldg(cc + 2);
      // A (Coupling.cpp:30-38)
      cplx b0 = vtac_beta(G, tb.off, n, mu, l, kk);
      cplx bpv = vtac_beta(G, tb.off, n, mu + 1, l, kk + 1);
      cplx bmv = vtac_beta(G, tb.off, n, mu - 1, l, kk - 1);
      double c0 = (double)(2 * kk * mu), c1 = s1 * t1, c2 = s2 * t2;
      cplx a;
      a.x = c0 * b0.x + c1 * bpv.x + c2 * bmv.x;
      a.y = c0 * b0.y + c1 * bpv.y + c2 * bmv.y;
      a = cscale(a, fA * gn);
      // B (Coupling.cpp:40-51): factor -i/2 sqrt(...)
      cplx g0 = vtac_beta(G, tb.off, n, mu, l - 1, kk);
      cplx gp = vtac_beta(G, tb.off, n, mu + 1, l - 1, kk + 1);
      cplx gm = vtac_beta(G, tb.off, n, mu - 1, l - 1, kk - 1);
      double d0 = (double)(2 * mu) * u0, d1 = s1 * u1, d2 = s2 * u2;
      cplx b;
      b.x = d0 * g0.x + d1 * gp.x - d2 * gm.x;
      b.y = d0 * g0.y + d1 * gp.y - d2 * gm.y;
      double fb = fB * gn;
      b = mk(b.y * fb, -b.x * fb); // times -i
      cplx phs = sm.ph[mu - kk + 2 * NM];
      emit.item(p, row, cmul(a, phs), cmul(b, phs));
    }
  };

  // ---- level march ----
  compute_level(1);
  __syncthreads();
  for(int n = 1; n <= NM; ++n) {
    if(n < NM)
      compute_level(n + 1);
    emit_level(n);
    __syncthreads();
  }
}

} // namespace ob
