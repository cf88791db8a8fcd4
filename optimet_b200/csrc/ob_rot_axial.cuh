// ob_rot_axial.cuh -- axial-only translation recursion for one pair, written once for the device (one warp per pair,
// lanes stride the items of a level) and for the host (lane 0 of 1: the CPU check of this very source in
// tests/test_rot_axial_host.py, which compiles tests/rot_axial_host.cpp around this header).
//
// With theta = 0 the scalar coefficients beta(n, m, l, k) vanish unless k = m and the reference's recursion
// (srcAna/TranslationAdditionCoefficients.cpp:102-124) closes on those entries: O(nMax^3) per pair instead of the
// O(nMax^4) of the full block.  Transliterated from tests/rot_axial_model.py.
#pragma once
#include "ob_common.cuh"
#include "ob_special.cuh"
#include <vector>

#if defined(__CUDA_ARCH__)
#define OB_SYNCWARP() __syncwarp()
#else
#define OB_SYNCWARP() ((void)0)
#endif

namespace ob {

__host__ __device__ constexpr inline int rot_n0(int mu) { return mu > 1 ? mu : 1; }
__host__ __device__ constexpr inline int rot_offX(int NM, int mu) { // sum_{u<mu} (NM - max(u,1) + 1)^2
  // mu >= 1: NM^2 + sum_{s = NM - mu + 2}^{NM} s^2
  return mu <= 0 ? 0
                 : NM * NM + (NM * (NM + 1) * (2 * NM + 1) / 6 - (NM - mu + 1) * (NM - mu + 2) * (2 * (NM - mu + 1) + 1) / 6);
}
// FRAGMENT ORDER of the record's matrices.  The apply kernel (k_matvec_rot) reads every matrix as A operands of the FP64
// tensor-core instruction DMMA.8x8x4: lane (fr, fc) = (lane >> 2, lane & 3) holds element (row 8 t + fr, K index 4 s + fc)
// of row tile t and K step s.  Stored row- or column-major with the odd leading dimensions of these matrices, the 16
// lanes of a half-warp hit every 8-byte bank up to twice (4 shared-memory wavefronts per fragment where 2 is the
// minimum: ncu, profiles/r3c_matvec_rot_full.txt).  So the assembly kernels write each fragment as ONE CONTIGUOUS RUN in
// lane order, compacted to its valid rows (Rt <= 8) and valid K entries (Kv <= 4): element (row, kcol) of a rows x K
// matrix sits at  8 t K + Rt 4 s + (row - 8 t) Kv + (kcol - 4 s).  Same bytes as the plain layout (no padding), and a
// half-warp reads <= 16 consecutive doubles: conflict-free.  Lanes past Rt or Kv read neighbouring finite entries
// (rows of a product are independent; K-tail entries are zeroed in registers).
__host__ __device__ constexpr inline int rot_frag_index(int rows, int K, int row, int kcol) {
  const int t = row >> 3, Rt = rows - 8 * t < 8 ? rows - 8 * t : 8, s = kcol >> 2, Kv = K - 4 * s < 4 ? K - 4 * s : 4;
  return 8 * t * K + Rt * 4 * s + (row - 8 * t) * Kv + (kcol - 4 * s);
}
// the a class of a small-d matrix: rows a = 1 .. n keep the lane of row a of the s class (fr = a - 8 t), so tile 0 holds
// the seven rows a = 1 .. 7 behind a phantom row 0; K index kcol = a' - 1
__host__ __device__ constexpr inline int rot_frag_index_a(int n, int a, int kcol) {
  const int t = a >> 3, first = t ? 8 * t : 1, last = 8 * t + 7 < n ? 8 * t + 7 : n, Rt = last - first + 1;
  const int s = kcol >> 2, Kv = n - 4 * s < 4 ? n - 4 * s : 4;
  return (first - 1) * n + Rt * 4 * s + (a - first) * Kv + (kcol - 4 * s);
}
// record index of the axial coefficient (mu; column degree n, row degree l): the apply contracts over n (K index) and
// produces l (row); the plain layout was rot_offX(mu) + (n - n0) w + (l - n0)
__host__ __device__ constexpr inline int rot_cidx(int NM, int mu, int n, int l) {
  const int n0 = rot_n0(mu), w = NM - n0 + 1;
  return rot_offX(NM, mu) + rot_frag_index(w, w, l - n0, n - n0);
}
__host__ __device__ inline double ta_a_plus(int n, int m) {
  return -sqrt((double)((n + m + 1) * (n - m + 1)) / (double)((2 * n + 1) * (2 * n + 3)));
}
__host__ __device__ inline double ta_a_minus(int n, int m) {
  return sqrt((double)((n + m) * (n - m)) / (double)((2 * n + 1) * (2 * n - 1)));
}
__host__ __device__ inline double ta_b_plus(int n, int m) {
  return sqrt((double)((n + m + 2) * (n + m + 1)) / (double)((2 * n + 1) * (2 * n + 3)));
}
__host__ __device__ inline double ta_b_minus(int n, int m) {
  return sqrt((double)((n - m) * (n - m - 1)) / (double)((2 * n + 1) * (2 * n - 1)));
}
// entries of one warp's level buffers: [3][NM + 2][2 NM + 3] complex (level n % 3, chain m, degree l); entries with
// l < m are never written and must be zero on entry (and stay zero)
__host__ __device__ inline int rot_axial_buf_entries(int NM) { return 3 * (NM + 2) * (2 * NM + 3); }

// Index-only coefficient tables of the recursion and of the A / B emission for one nMax (the square roots of the
// reference's a+-, b+- and of the Coupling prefactors depend on (n, m, l) only; evaluating them per pair was most of the
// assembly time).  Built on the host by rot_axial_tables_build from the very expressions of the table-free path below,
// which stays for the host check and as the specification.
//   rec  [c][offR(n) + t], c = 0..3: c0, c1, c2, 1 / denominator     of item t = m * span + (l - m) of level n
//   emit [c][offE(n) + t], c = 0..7: fa, a0, a1, a2, fb, b0, b1, b2   of item t = mu * NM + (l - 1) of level n
//   (structure of arrays: the lanes of a warp read consecutive items, i.e. consecutive doubles of one component)
//   ridx = m | l << 8 | flags << 16: 1 live, 2 beta(n-1, m, l-1) exists, 4 beta(n-2, m, l) exists, 8 sectorial step
//   eidx = mu | l << 8 | flags << 16: 1 live, then one bit per term of the A / B sums that exists (2, 4, 8: beta(mu, l),
//          beta(mu + 1, l), beta(|mu - 1|, l); 16, 32, 64: the same at l - 1); eout = index of the entry in the record
struct RotAxTab {
  const double *rec, *emit;
  const int *ridx, *eidx, *eout;
  int nrec, nem; // entries per component
};
__host__ __device__ inline int rot_axial_offR(int NM, int n) { // sum_{j=1}^{n-1} (j + 1)(2 NM - j + 1)
  int o = 0;
  for(int j = 1; j < n; ++j)
    o += (j + 1) * (2 * NM - j + 1);
  return o;
}
__host__ __device__ inline int rot_axial_offE(int NM, int n) { return NM * ((n - 1) * (n + 2) / 2); } // NM sum_{j<n} (j + 1)
inline void rot_axial_tables_build(int NM, std::vector<double> &rec, std::vector<double> &emit, std::vector<int> &ridx,
                                   std::vector<int> &eidx, std::vector<int> &eout) {
  const int LL = 2 * NM;
  rec.assign((size_t)rot_axial_offR(NM, NM + 1) * 4, 0.0);
  ridx.assign((size_t)rot_axial_offR(NM, NM + 1), 0);
  emit.assign((size_t)rot_axial_offE(NM, NM + 1) * 8, 0.0);
  eidx.assign((size_t)rot_axial_offE(NM, NM + 1), 0);
  eout.assign((size_t)rot_axial_offE(NM, NM + 1), 0);
  for(int n = 1; n <= NM; ++n) {
    const int span = LL - n + 1;
    for(int t = 0; t < (n + 1) * span; ++t) {
      const int m = t / span, l = m + (t - m * span);
      const size_t NR = (size_t)rot_axial_offR(NM, NM + 1), ir = (size_t)rot_axial_offR(NM, n) + t;
      double c[4] = {0, 0, 0, 0};
      if(l > LL - n) {
        ridx[(size_t)rot_axial_offR(NM, n) + t] = m | (l << 8);
        continue;
      }
      {
        const int mr = m == n ? n - 1 : m;
        const int fl = 1 | (l - 1 >= mr ? 2 : 0) | ((m != n && n - 2 >= m) ? 4 : 0) | (m == n ? 8 : 0);
        ridx[(size_t)rot_axial_offR(NM, n) + t] = m | (l << 8) | (fl << 16);
      }
      if(m == n) {
        c[0] = l - 1 >= n - 1 ? ta_b_plus(l - 1, n - 1) : 0.0;
        c[1] = ta_b_minus(l + 1, n - 1);
        c[2] = 0.0;
        c[3] = 1.0 / ta_b_plus(n - 1, n - 1);
      } else {
        c[0] = l - 1 >= m ? ta_a_plus(l - 1, m) : 0.0;
        c[1] = ta_a_minus(l + 1, m);
        c[2] = n - 2 >= m ? ta_a_minus(n - 1, m) : 0.0;
        c[3] = 1.0 / ta_a_plus(n - 1, m);
      }
      for(int q = 0; q < 4; ++q)
        rec[q * NR + ir] = c[q];
    }
    for(int t = 0; t < (n + 1) * NM; ++t) {
      const int mu = t / NM, l = 1 + (t - mu * NM);
      const int n0 = rot_n0(mu);
      const bool live = !(n < n0 || l < n0);
      if(!live) {
        eidx[(size_t)rot_axial_offE(NM, n) + t] = mu | (l << 8);
        continue;
      }
      {
        const int mp1 = mu + 1, mm1 = mu > 0 ? mu - 1 : 1, lm = l - 1;
        const int fl = 1 | ((mu <= n && mu <= l) ? 2 : 0) | ((mp1 <= n && mp1 <= l) ? 4 : 0) | ((mm1 <= n && mm1 <= l) ? 8 : 0) |
                       ((mu <= n && mu <= lm) ? 16 : 0) | ((mp1 <= n && mp1 <= lm) ? 32 : 0) | ((mm1 <= n && mm1 <= lm) ? 64 : 0);
        eidx[(size_t)rot_axial_offE(NM, n) + t] = mu | (l << 8) | (fl << 16);
        eout[(size_t)rot_axial_offE(NM, n) + t] = rot_cidx(NM, mu, l, n); // written through the transpose symmetry, see below
      }
      const size_t NE = (size_t)rot_axial_offE(NM, NM + 1), ie = (size_t)rot_axial_offE(NM, n) + t;
      double c[8];
      c[0] = 0.5 / sqrt((double)(l * (l + 1) * n * (n + 1)));
      c[1] = 2.0 * mu * mu;
      c[2] = sqrt((double)((n - mu) * (n + mu + 1) * (l - mu) * (l + mu + 1)));
      c[3] = sqrt((double)((n + mu) * (n - mu + 1) * (l + mu) * (l - mu + 1)));
      c[4] = -0.5 * sqrt((2.0 * l + 1.0) / ((double)(2 * l - 1) * (double)(l * (l + 1)) * (double)(n * (n + 1))));
      c[5] = 2.0 * mu * sqrt((double)((l - mu) * (l + mu)));
      c[6] = sqrt((double)((n - mu) * (n + mu + 1) * (l - mu) * (l - mu - 1)));
      c[7] = sqrt((double)((n + mu) * (n - mu + 1) * (l + mu) * (l + mu - 1)));
      // Level n produces (n, l) for every l, i.e. one K index of every ROW of the fragment order: 8-byte stores 32 bytes
      // apart.  A(n, l) = (-1)^(n + l) A(l, n) and the same for B (reciprocity of the axial translation), so the level
      // writes sg (n, l) into the place of (l, n) instead: the lanes of a level fill runs of consecutive K entries.  The
      // sign rides on the two prefactors (exact)
      if((n + l) & 1) {
        c[0] = -c[0];
        c[4] = -c[4];
      }
      for(int q = 0; q < 8; ++q)
        emit[q * NE + ie] = c[q];
    }
  }
}

// Axial A[(n,mu),(l,mu)], B[...] for translation r along z with wavenumber k into Aout / Bout (plain compact layout
// e = rot_offX(mu) + (n - n0)(NM - n0 + 1) + (l - n0)).  `lane` of `nlanes` cooperating threads; buf as above.
// combine = 1: Aout[e] = A + B for every mu and Bout[e - NM^2] = A - B for mu >= 1 (interleaved complex).
// combine = 2 (record layout of ob_rot.cu): planes of doubles in FRAGMENT ORDER, entry (l, n) written as (-1)^(n + l) (n, l),
// Aout -> [Re(A+B)[X] | Im(A+B)[X]], Bout -> [Re(A-B)[X - NM^2] | Im(A-B)[X - NM^2]], X = rot_offX(NM, NM + 1).
__host__ __device__ inline void rot_axial_pair(int NM, cplx k, double r, cplx *buf, cplx *Aout, cplx *Bout, int lane,
                                               int nlanes, int combine = 0) {
  const int LL = 2 * NM, W = LL + 3, CH = NM + 2;
  // seeds (n = m = 0): sqrt(4 pi) (-1)^l Y_l0(0) h_l = (-1)^l sqrt(2l + 1) h_l(k r); every lane runs the short upward
  // Hankel recurrence and keeps the orders it owns
  {
    cplx h[2 * 13 + 2];
    sph_hankel1(cscale(k, r), LL + 1, h);
    for(int l = lane; l <= LL; l += nlanes) {
      const double f = ((l & 1) ? -1.0 : 1.0) * sqrt(2.0 * l + 1.0);
      buf[(0 * CH + 0) * W + l] = cscale(h[l], f);
    }
  }
  OB_SYNCWARP();
  for(int n = 1; n <= NM; ++n) {
    cplx *cur = buf + (size_t)(n % 3) * CH * W;
    const cplx *p1 = buf + (size_t)((n - 1) % 3) * CH * W, *p2 = buf + (size_t)((n + 1) % 3) * CH * W; // n-1, n-2
    const int span = LL - n + 1; // l in [m, LL - n]: index t = m * span + (l - m) over a (n + 1) x span rectangle
    for(int t = lane; t < (n + 1) * span; t += nlanes) {
      const int m = t / span, l = m + (t - m * span);
      if(l > LL - n)
        continue;
      cplx v;
      if(m == n) { // sectorial step (:113-117)
        const cplx lo = l - 1 >= n - 1 ? p1[(n - 1) * W + (l - 1)] : mk(0, 0);
        const cplx up = p1[(n - 1) * W + (l + 1)];
        const double c0 = l - 1 >= n - 1 ? ta_b_plus(l - 1, n - 1) : 0.0, c1 = ta_b_minus(l + 1, n - 1);
        const double inv = 1.0 / ta_b_plus(n - 1, n - 1);
        v = mk((lo.x * c0 + up.x * c1) * inv, (lo.y * c0 + up.y * c1) * inv);
      } else { // general step (:119-124)
        const cplx lo = l - 1 >= m ? p1[m * W + (l - 1)] : mk(0, 0);
        const cplx up = p1[m * W + (l + 1)];
        const cplx o = n - 2 >= m ? p2[m * W + l] : mk(0, 0);
        const double c0 = l - 1 >= m ? ta_a_plus(l - 1, m) : 0.0, c1 = ta_a_minus(l + 1, m);
        const double c2 = n - 2 >= m ? ta_a_minus(n - 1, m) : 0.0, inv = 1.0 / ta_a_plus(n - 1, m);
        v = mk((lo.x * c0 + up.x * c1 - o.x * c2) * inv, (lo.y * c0 + up.y * c1 - o.y * c2) * inv);
      }
      cur[m * W + l] = v;
    }
    OB_SYNCWARP();
    // A, B of column degree n for every mu <= n and row degree l (Coupling.cpp:30-51 with k = m = mu)
    for(int t = lane; t < (n + 1) * NM; t += nlanes) {
      const int mu = t / NM, l = 1 + (t - mu * NM);
      if(n < rot_n0(mu) || l < rot_n0(mu))
        continue;
      const double fa = 0.5 / sqrt((double)(l * (l + 1) * n * (n + 1)));
      const double a0 = 2.0 * mu * mu;
      const double a1 = sqrt((double)((n - mu) * (n + mu + 1) * (l - mu) * (l + mu + 1)));
      const double a2 = sqrt((double)((n + mu) * (n - mu + 1) * (l + mu) * (l - mu + 1)));
      const double fb = -0.5 * sqrt((2.0 * l + 1.0) / ((double)(2 * l - 1) * (double)(l * (l + 1)) * (double)(n * (n + 1))));
      const double b0 = 2.0 * mu * sqrt((double)((l - mu) * (l + mu)));
      const double b1 = sqrt((double)((n - mu) * (n + mu + 1) * (l - mu) * (l - mu - 1)));
      const double b2 = sqrt((double)((n + mu) * (n - mu + 1) * (l + mu) * (l + mu - 1)));
      const int n0 = rot_n0(mu);
      const int mp1 = mu + 1, mm1 = mu > 0 ? mu - 1 : 1; // |mu - 1|: beta(n,-m,l,-m) = beta(n,m,l,m)
      // beta(n, m', l', m') at this level; zero outside 0 <= m' <= min(n, l')
      const cplx t0 = (mu <= n && mu <= l) ? cur[mu * W + l] : mk(0, 0);
      const cplx tp = (mp1 <= n && mp1 <= l) ? cur[mp1 * W + l] : mk(0, 0);
      const cplx tm = (mm1 <= n && mm1 <= l) ? cur[mm1 * W + l] : mk(0, 0);
      const cplx Av = mk(fa * (a0 * t0.x + a1 * tp.x + a2 * tm.x), fa * (a0 * t0.y + a1 * tp.y + a2 * tm.y));
      const int lm = l - 1;
      const cplx u0 = (lm >= 0 && mu <= n && mu <= lm) ? cur[mu * W + lm] : mk(0, 0);
      const cplx up = (lm >= 0 && mp1 <= n && mp1 <= lm) ? cur[mp1 * W + lm] : mk(0, 0);
      const cplx um = (lm >= 0 && mm1 <= n && mm1 <= lm) ? cur[mm1 * W + lm] : mk(0, 0);
      const cplx sB = mk(b0 * u0.x + b1 * up.x - b2 * um.x, b0 * u0.y + b1 * up.y - b2 * um.y);
      const cplx Bv = mk(-fb * sB.y, fb * sB.x); // times i fb (factor = (0, fb))
      const int w = NM - n0 + 1, e = rot_offX(NM, mu) + (n - n0) * w + (l - n0);
      if(combine == 2) {
        // sg (n, l) stored as entry (l, n): (n, l) = (-1)^(n + l) (l, n), see rot_axial_tables_build
        const int X = rot_offX(NM, NM + 1), XM = X - NM * NM, f = rot_cidx(NM, mu, l, n);
        const double sg = ((n + l) & 1) ? -1.0 : 1.0;
        double *P = (double *)Aout, *M = (double *)Bout;
        P[f] = sg * (Av.x + Bv.x);
        P[X + f] = sg * (Av.y + Bv.y);
        if(mu >= 1) {
          M[f - NM * NM] = sg * (Av.x - Bv.x);
          M[XM + f - NM * NM] = sg * (Av.y - Bv.y);
        }
      } else if(combine) {
        Aout[e] = cadd(Av, Bv);
        if(mu >= 1)
          Bout[e - NM * NM] = csub(Av, Bv);
      } else {
        Aout[e] = Av;
        Bout[e] = Bv;
      }
    }
    OB_SYNCWARP();
  }
}

// The path the kernel runs: the same recursion and emission with every index-only quantity read from RotAxTab, planar
// record output (combine = 2 of rot_axial_pair), and TWO level buffers instead of three: level n overwrites level n - 2 in
// place (an item reads the n - 2 value of its own (m, l) only, then writes there; everything else it reads is level
// n - 1).  buf: [2][NM + 2][2 NM + 3] complex per warp; stale entries are never read (the flags say which terms exist).
// tests/test_rot_axial_host.py holds it bit-identical to rot_axial_pair on the host.
__host__ __device__ inline int rot_axial_fast_entries(int NM) { return 2 * (NM + 2) * (2 * NM + 3); }
// seeds (n = m = 0): sqrt(4 pi) (-1)^l Y_l0(0) h_l = (-1)^l sqrt(2l + 1) h_l(k r), order l kept by lane l mod nlanes
struct RotAxialSeed {
  cplx *buf;
  int LL, lane, nlanes;
  __host__ __device__ void operator()(int l, cplx h) {
    if(l <= LL && l % nlanes == lane) {
      const double f = ((l & 1) ? -1.0 : 1.0) * sqrt(2.0 * l + 1.0);
      buf[l] = cscale(h, f);
    }
  }
};
__host__ __device__ inline void rot_axial_pair_fast(int NM, cplx k, double r, cplx *buf, double *Cp, double *Cm, int lane,
                                                    int nlanes, RotAxTab const &tab) {
  const int LL = 2 * NM, W = LL + 3, CH = NM + 2;
  const int X = rot_offX(NM, NM + 1), XM = X - NM * NM;
  { // every lane runs the short upward Hankel recurrence in registers and stores the orders it owns (level 0, chain 0)
    RotAxialSeed seed = {buf, LL, lane, nlanes};
    sph_hankel1_each(cscale(k, r), LL + 1, seed);
  }
  OB_SYNCWARP();
  int offR = 0, offE = 0;
  for(int n = 1; n <= NM; ++n) {
    cplx *cur = buf + (size_t)(n & 1) * CH * W;            // holds level n - 2, becomes level n
    const cplx *p1 = buf + (size_t)((n & 1) ^ 1) * CH * W; // level n - 1
    const int nrec = (n + 1) * (LL - n + 1), nem = (n + 1) * NM;
    const double *rc = tab.rec + offR;
    const int *ri = tab.ridx + offR;
    const int NR = tab.nrec, NE = tab.nem;
    for(int t = lane; t < nrec; t += nlanes) {
      const int ix = ri[t], fl = ix >> 16;
      if(!(fl & 1))
        continue;
      const int m = ix & 0xff, l = (ix >> 8) & 0xff;
      const double c0 = rc[t], c1 = rc[NR + t], c2 = rc[2 * NR + t], inv = rc[3 * NR + t];
      const cplx *q = p1 + ((fl & 8) ? n - 1 : m) * W + l; // the sectorial step reads chain n - 1
      const cplx lo = (fl & 2) ? q[-1] : mk(0, 0), up = q[1];
      const cplx o = (fl & 4) ? cur[m * W + l] : mk(0, 0);
      cur[m * W + l] = mk((lo.x * c0 + up.x * c1 - o.x * c2) * inv, (lo.y * c0 + up.y * c1 - o.y * c2) * inv);
    }
    OB_SYNCWARP();
    const double *ec = tab.emit + offE;
    const int *ei = tab.eidx + offE, *eo = tab.eout + offE;
    for(int t = lane; t < nem; t += nlanes) {
      const int ix = ei[t], fl = ix >> 16;
      if(!(fl & 1))
        continue;
      const int mu = ix & 0xff, l = (ix >> 8) & 0xff, mm1 = mu > 0 ? mu - 1 : 1;
      const double *c = ec + t;
      const cplx *b0p = cur + mu * W + l, *bpp = b0p + W, *bmp = cur + mm1 * W + l;
      const cplx t0 = (fl & 2) ? b0p[0] : mk(0, 0), tp = (fl & 4) ? bpp[0] : mk(0, 0), tm = (fl & 8) ? bmp[0] : mk(0, 0);
      const cplx u0 = (fl & 16) ? b0p[-1] : mk(0, 0), up = (fl & 32) ? bpp[-1] : mk(0, 0), um = (fl & 64) ? bmp[-1] : mk(0, 0);
      const double fa = c[0], a0 = c[NE], a1 = c[2 * NE], a2 = c[3 * NE], fb = c[4 * NE], b0 = c[5 * NE], b1 = c[6 * NE],
                   b2 = c[7 * NE];
      const cplx Av = mk(fa * (a0 * t0.x + a1 * tp.x + a2 * tm.x), fa * (a0 * t0.y + a1 * tp.y + a2 * tm.y));
      const cplx sB = mk(b0 * u0.x + b1 * up.x - b2 * um.x, b0 * u0.y + b1 * up.y - b2 * um.y);
      const cplx Bv = mk(-fb * sB.y, fb * sB.x); // times i fb
      const int e = eo[t];
      Cp[e] = Av.x + Bv.x;
      Cp[X + e] = Av.y + Bv.y;
      if(mu >= 1) {
        Cm[e - NM * NM] = Av.x - Bv.x;
        Cm[XM + e - NM * NM] = Av.y - Bv.y;
      }
    }
    OB_SYNCWARP();
    offR += nrec;
    offE += nem;
  }
}

} // namespace ob
