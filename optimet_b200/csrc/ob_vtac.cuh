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
  const int *off;                      // chain offsets, [NM+2]
  const double *rowc;                  // per row r=(l,k): fA, fB, t1, t2, u0, u1, u2, pad   (Coupling.cpp:33-37, 43-48)
  const double *colc;                  // per column p=(n,mu): gn, s1, s2, pad
  const double *legA, *legB;           // Legendre recurrence coefficients at l(l+1)/2+m
  const double *legD, *legF;           // sqrt(2m+3) and -sqrt((2m+1)/(2m)) at m
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
  b += (size_t)(L + 2) * sizeof(cplx);                          // radial z_l + one zero entry (target of masked reads)
  b += (size_t)(4 * NM + 1) * sizeof(cplx);                     // phase table
  b += (size_t)((L + 1) * (L + 2) / 2) * sizeof(double);        // Legendre
  b += (size_t)((NM + 2 + 2 + 1) & ~1) * sizeof(int);            // chain offsets (even count: doubles follow)
  b += (size_t)2 * (L + 2) * (L + 2) * sizeof(double);          // a+, a- tables of the general step
  b += (size_t)(((L + 1) * (L + 1) + 15) & ~15);                // lam of a triangular index (bytes)
  return b;
}

struct VtacSmem {
  cplx *buf[2];
  cplx *zl;
  cplx *ph;
  double *nlm;
  int *offs;
  double *aps, *ams;
  unsigned char *lamS;
  __device__ VtacSmem(unsigned char *base, int NM) {
    int T = vtac_buffer_entries(NM), L = 2 * NM;
    buf[0] = (cplx *)base;
    buf[1] = buf[0] + T;
    zl = buf[1] + T;
    ph = zl + (L + 2);
    nlm = (double *)(ph + (4 * NM + 1));
    offs = (int *)(nlm + (L + 1) * (L + 2) / 2);
    aps = (double *)(offs + ((NM + 2 + 2 + 1) & ~1));
    ams = aps + (L + 2) * (L + 2);
    lamS = (unsigned char *)(ams + (L + 2) * (L + 2));
  }
};

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

  // ---- seeds: radial sequence (one thread), Legendre (one thread per order m), phases, chain offsets ----
  if(tid == 0) {
    sm.zl[L + 1] = mk(0, 0);
    cplx z = cscale(k, r);
    if(regular)
      sph_bessel_j(z, L, sm.zl);
    else
      sph_hankel1(z, L, sm.zl);
  }
  if(tid >= 32 && tid < 32 + L + 1) {
    // sqrt(4 pi) N_l^m(cos theta), l = m..L, at nlm[l(l+1)/2 + m]; the square-root coefficients come from host tables
    // (same IEEE operations as computing them here, without the FP64 sqrt/div latency chain)
    const int m = tid - 32;
    double x = cos(the), s = sin(the);
    if(s < 0)
      s = -s;
    double pmm = 1.0; // sqrt(4 pi) folded in: sqrt(4pi) * sqrt(1/(4pi))
    for(int i = 1; i <= m; ++i)
      pmm *= __ldg(tb.legF + i) * s;
    sm.nlm[m * (m + 1) / 2 + m] = pmm;
    if(L > m) {
      double pmmp1 = x * __ldg(tb.legD + m) * pmm;
      sm.nlm[(m + 1) * (m + 2) / 2 + m] = pmmp1;
      for(int l = m + 2; l <= L; ++l) {
        const double a = __ldg(tb.legA + l * (l + 1) / 2 + m), b = __ldg(tb.legB + l * (l + 1) / 2 + m);
        const double pll = a * (x * pmmp1 - b * pmm);
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
  if(tid >= 192 && tid < 192 + NM + 2)
    sm.offs[tid - 192] = tb.off[tid - 192];
  for(int e = tid; e < (L + 2) * (L + 2); e += nthr) { // general-step coefficients and the lam lookup, CTA-local
    sm.aps[e] = __ldg(tb.ap + e);
    sm.ams[e] = __ldg(tb.am + e);
    if(e < (L + 1) * (L + 1))
      sm.lamS[e] = (unsigned char)__fsqrt_rn((float)e); // e in [lam^2, (lam+1)^2)
  }
  __syncthreads();
  {
    cplx *G = sm.buf[0];
    const int sz = (L + 1) * (L + 1);
    for(int idx = tid; idx < sz; idx += nthr) {
      const int lam = sm.lamS[idx];
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
  int rl = 1, rk = 0;
  double rA0 = 0, rA1 = 0, rA2 = 0, rB0 = 0, rB1 = 0, rB2 = 0;
  if(row_active) {
    unflatten(row, rl, rk);
    const double *rc = tb.rowc + 8 * row;
    const double fA = rc[0], fB = rc[1];
    rA0 = fA * (double)(2 * rk); // with beta(n,mu,l,k) (times mu)
    rA1 = fA * rc[2];            // with beta(n,mu+1,l,k+1)
    rA2 = fA * rc[3];            // with beta(n,mu-1,l,k-1)
    rB0 = fB * 2.0 * rc[4];      // with beta(n,mu,l-1,k) (times mu)
    rB1 = fB * rc[5];            // with beta(n,mu+1,l-1,k+1)
    rB2 = fB * rc[6];            // with beta(n,mu-1,l-1,k-1)
  }
  // index bases of this row inside one chain: lam = l and lam = l-1; the mirrored (-kappa) ones serve mu' < 0
  const int b0 = rl * (rl + 1) + rk, b1 = rl * (rl - 1) + rk;
  const int b0n = b0 - 2 * rk, b1n = b1 - 2 * rk;
  // validity of kappa' = k + d at lam = l (always for d = 0) and lam = l - 1; an invalid entry always meets an
  // exactly-zero coefficient (the square roots in Coupling.cpp:33-37, 43-48 vanish there), it only must not be read
  const bool vP0 = rk < rl, vM0 = rk > -rl, vZ1 = (rk < rl) && (rk > -rl), vP1 = rk <= rl - 2, vM1 = rk >= 2 - rl;

  // one shared-memory base with arithmetic offsets (a pointer array indexed by n & 1 makes the compiler fall back to
  // generic loads with 64-bit address arithmetic)
  cplx *const buf0 = sm.buf[0];
  const int Tbuf = tb.T;

  auto compute_level = [&](int n) {
    const int Ln = L - n;
    const int sz = (Ln + 1) * (Ln + 1);
    const int items = (n + 1) * sz;
    const int dbase = (n & 1) * Tbuf, sbase = ((n - 1) & 1) * Tbuf;
    const double inv_bp = __ldg(tb.inv_bp_n + n);
    // (m, idx) of item `it`, advanced incrementally: it += nthr  <=>  m += qm, idx += rm (+ carry)
    const int qm = nthr / sz, rm = nthr - qm * sz;
    int m = tid / sz, idx = tid - m * sz;
    for(int it = tid; it < items; it += nthr) {
      const int lam = sm.lamS[idx];
      const int kap = idx - lam * (lam + 1);
      const int ak = kap < 0 ? -kap : kap;
      cplx v = mk(0, 0);
      const int up = idx + 2 * lam + 2, dn = idx - 2 * lam; // (lam+1, kap) and (lam-1, kap)
      if(m == n) { // sectorial step (TranslationAdditionCoefficients.cpp:113-117)
        const int so = sbase + sm.offs[n - 1];
        const int k1 = kap - 1;
        if(lam >= 1 && (k1 < 0 ? -k1 : k1) <= lam - 1)
          v = cscale(buf0[so + dn - 1], __ldg(tb.bp + dn - 1));
        const cplx w = buf0[so + up - 1];
        const double c = __ldg(tb.bm + up - 1);
        v.x = fma(w.x, c, v.x);
        v.y = fma(w.y, c, v.y);
        v = cscale(v, inv_bp);
      } else { // general step (:119-124)
        const int om = sm.offs[m];
        const int so = sbase + om;
        if(ak <= lam - 1)
          v = cscale(buf0[so + dn], sm.aps[dn]);
        const cplx w = buf0[so + up];
        const double c = sm.ams[up];
        v.x = fma(w.x, c, v.x);
        v.y = fma(w.y, c, v.y);
        if(n - 2 >= m) {
          const cplx o = buf0[dbase + om + idx];
          const double a = __ldg(tb.am_nm + n * (NM + 1) + m);
          v.x = fma(-o.x, a, v.x);
          v.y = fma(-o.y, a, v.y);
        }
        v = cscale(v, __ldg(tb.inv_ap_nm + n * (NM + 1) + m));
      }
      buf0[dbase + sm.offs[m] + idx] = v;
      m += qm;
      idx += rm;
      if(idx >= sz) {
        idx -= sz;
        ++m;
      }
    }
  };

  // One (column mu, row) item: six shared-memory reads (masked ones hit the zero slot), A and B, phase, store.
  auto emit_item = [&](const cplx *G, int p, int mu, int i_z0, int i_y0, int i_zp, int i_yp, int i_zm, int i_ym,
                       double f0, double fp, double fm) {
    const double *cc = tb.colc + 4 * p;
    const double gn = __ldg(cc), s1 = __ldg(cc + 1), s2 = __ldg(cc + 2);
    const cplx z0 = G[i_z0], y0 = G[i_y0], zp = G[i_zp], yp = G[i_yp], zm = G[i_zm], ym = G[i_ym];
    const double k0 = gn * (double)mu * f0, k1 = gn * s1 * fp, k2 = gn * s2 * fm;
    // A (Coupling.cpp:30-38) and B (Coupling.cpp:40-51; factor -i/2 sqrt(...))
    const double ca0 = rA0 * k0, ca1 = rA1 * k1, ca2 = rA2 * k2;
    const double cb0 = rB0 * k0, cb1 = rB1 * k1, cb2 = -(rB2 * k2);
    cplx a, b;
    a.x = ca0 * z0.x + ca1 * zp.x + ca2 * zm.x;
    a.y = ca0 * z0.y + ca1 * zp.y + ca2 * zm.y;
    b.x = cb0 * y0.x + cb1 * yp.x + cb2 * ym.x;
    b.y = cb0 * y0.y + cb1 * yp.y + cb2 * ym.y;
    b = mk(b.y, -b.x); // times -i
    const cplx phs = sm.ph[mu - rk + 2 * NM];
    emit.item(p, row, cmul(a, phs), cmul(b, phs));
  };

  auto emit_level = [&](int n) {
    if(!row_active)
      return;
    const cplx *G = buf0 + (n & 1) * Tbuf;
    const int ZI = (int)(sm.zl + (L + 1) - G); // index of the zero entry seen from this level buffer
    const int pbase = n * (n + 1) - 1;         // flat(n, mu) = pbase - mu
    // Masked entries (outside the chain, or with an exactly vanishing coefficient: the square roots in
    // Coupling.cpp:33-37, 43-48 vanish there) read the zero slot.  Entries with mu' < 0 come from the mirrored entry of
    // chain |mu'| with sign (-1)^(mu'+kappa'); the parity of mu' + kappa' = 2 mu' + k - mu is the same for
    // mu' = mu-1, mu, mu+1.  All threads of a warp share mu (a warp is 32 consecutive rows of one column group), so
    // the three cases below do not diverge.
#pragma unroll 1
    for(int mu = n - grp; mu >= -n; mu -= ngroups) {
      const int p = pbase - mu;
      if(mu >= 1 && mu < n) { // mu - 1, mu, mu + 1 all in [0, n]: direct entries, no sign
        const int o0 = sm.offs[mu], op = sm.offs[mu + 1], om = sm.offs[mu - 1];
        emit_item(G, p, mu, o0 + b0, vZ1 ? o0 + b1 : ZI, vP0 ? op + b0 + 1 : ZI, vP1 ? op + b1 + 1 : ZI,
                  vM0 ? om + b0 - 1 : ZI, vM1 ? om + b1 - 1 : ZI, 1.0, 1.0, 1.0);
      } else if(mu <= -2 && mu > -n) { // all three mirrored
        const int a0 = -mu;
        const int o0 = sm.offs[a0], op = sm.offs[a0 - 1], om = sm.offs[a0 + 1];
        const double flip = ((rk - mu) & 1) ? -1.0 : 1.0;
        emit_item(G, p, mu, o0 + b0n, vZ1 ? o0 + b1n : ZI, vP0 ? op + b0n - 1 : ZI, vP1 ? op + b1n - 1 : ZI,
                  vM0 ? om + b0n + 1 : ZI, vM1 ? om + b1n + 1 : ZI, flip, flip, flip);
      } else { // mu in {n, 0, -1, -n}: mixed signs and the chain ends
        const double flip = ((rk - mu) & 1) ? -1.0 : 1.0;
        const int mp = mu + 1, mm = mu - 1;
        const int a0 = mu < 0 ? -mu : mu, ap = mp < 0 ? -mp : mp, am = mm < 0 ? -mm : mm;
        const int o0 = sm.offs[a0], op = sm.offs[ap], om = sm.offs[am]; // offs has NM + 2 entries: |mu'| = n + 1 is readable
        const bool hp = mu < n, hm = mu > -n;
        emit_item(G, p, mu, o0 + (mu >= 0 ? b0 : b0n), vZ1 ? o0 + (mu >= 0 ? b1 : b1n) : ZI,
                  (hp && vP0) ? op + (mp >= 0 ? b0 + 1 : b0n - 1) : ZI, (hp && vP1) ? op + (mp >= 0 ? b1 + 1 : b1n - 1) : ZI,
                  (hm && vM0) ? om + (mm >= 0 ? b0 - 1 : b0n + 1) : ZI, (hm && vM1) ? om + (mm >= 0 ? b1 - 1 : b1n + 1) : ZI,
                  mu < 0 ? flip : 1.0, mp < 0 ? flip : 1.0, mm < 0 ? flip : 1.0);
      }
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
