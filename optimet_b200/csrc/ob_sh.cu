// ob_sh.cu -- second-harmonic pieces: K5 (Clebsch-Gordan / Gaunt tables), K4 (SH source vectors)
// and K6b (SH absorption cross section).
//   reference: srcAna/Symbol.cpp:28-48, 143-150, 1036-1446 (tables; GSL gsl_sf_coupling_3j/6j/9j),
//              srcAna/Symbol.cpp:52-78, 155-353 + srcAna/Geometry.cpp:250-310 (v', u', u''),
//              srcAna/PreconditionedMatrix.cpp:1347-1436 (K, K1ana),
//              srcAna/Symbol.cpp:358-477 + srcAna/Geometry.cpp:428-455 + srcAna/Result.cpp:763-794 (abs SH).
//
// B200-first formulation: every table entry factors as  CG(J M | J1 M1 J2 M2) x red_t(J, J1, J2)
// (all nine closed forms carry that Clebsch-Gordan coefficient), so (i) the expensive 9j/6j sums are
// evaluated once per (J, J1, J2) triple, (ii) entries with M1 + M2 != M are exactly zero and the
// bilinear contractions skip them: per (particle, SH harmonic) ~ n * nMax pairs instead of n^2.
// The Bessel factors the reference re-evaluates inside the innermost loop (Symbol.cpp:200-204) depend
// on the order and the particle only: computed once per CTA into shared memory.
#include "ob_internal.h"
#include "ob_special.cuh"

namespace ob {

__constant__ double c_fact[171];

static void upload_factorials() {
  static bool done[64] = {false};
  if(!first_use_on_device(done))
    return;
  double f[171];
  f[0] = 1.0;
  for(int i = 1; i <= 170; ++i)
    f[i] = f[i - 1] * (double)i;
  OB_CUDA(cudaMemcpyToSymbol(c_fact, f, sizeof(f)));
}

__device__ __forceinline__ bool tri_bad(int a, int b, int c) { return c < abs(a - b) || c > a + b; }
__device__ __forceinline__ double tri_delta(int a, int b, int c) {
  return c_fact[a + b - c] * c_fact[a - b + c] * c_fact[-a + b + c] / c_fact[a + b + c + 1];
}
// gsl_sf_coupling_3j restated (integer j), Racah sum
__device__ double wigner3j(int j1, int j2, int j3, int m1, int m2, int m3) {
  if(j1 < 0 || j2 < 0 || j3 < 0)
    return 0;
  if(tri_bad(j1, j2, j3) || m1 + m2 + m3 != 0 || abs(m1) > j1 || abs(m2) > j2 || abs(m3) > j3)
    return 0;
  int kmin = max(0, max(j2 - j3 - m1, j1 - j3 + m2));
  int kmax = min(j1 + j2 - j3, min(j1 - m1, j2 + m2));
  double sum = 0;
  for(int k = kmin; k <= kmax; ++k) {
    double t = 1.0 / (c_fact[k] * c_fact[j1 + j2 - j3 - k] * c_fact[j1 - m1 - k] * c_fact[j2 + m2 - k] *
                      c_fact[j3 - j2 + m1 + k] * c_fact[j3 - j1 - m2 + k]);
    sum += (k & 1) ? -t : t;
  }
  double nrm = sqrt(tri_delta(j1, j2, j3) * c_fact[j1 + m1] * c_fact[j1 - m1] * c_fact[j2 + m2] * c_fact[j2 - m2] *
                    c_fact[j3 + m3] * c_fact[j3 - m3]);
  double r = nrm * sum;
  return (abs(j1 - j2 - m3) & 1) ? -r : r;
}
__device__ double wigner6j(int j1, int j2, int j3, int j4, int j5, int j6) {
  if(j1 < 0 || j2 < 0 || j3 < 0 || j4 < 0 || j5 < 0 || j6 < 0)
    return 0;
  if(tri_bad(j1, j2, j3) || tri_bad(j1, j5, j6) || tri_bad(j4, j2, j6) || tri_bad(j4, j5, j3))
    return 0;
  int a1 = j1 + j2 + j3, a2 = j1 + j5 + j6, a3 = j4 + j2 + j6, a4 = j4 + j5 + j3;
  int b1 = j1 + j2 + j4 + j5, b2 = j2 + j3 + j5 + j6, b3 = j3 + j1 + j6 + j4;
  int kmin = max(max(a1, a2), max(a3, a4));
  int kmax = min(b1, min(b2, b3));
  double sum = 0;
  for(int k = kmin; k <= kmax; ++k) {
    double t = c_fact[k + 1] / (c_fact[k - a1] * c_fact[k - a2] * c_fact[k - a3] * c_fact[k - a4] * c_fact[b1 - k] *
                                c_fact[b2 - k] * c_fact[b3 - k]);
    sum += (k & 1) ? -t : t;
  }
  return sqrt(tri_delta(j1, j2, j3) * tri_delta(j1, j5, j6) * tri_delta(j4, j2, j6) * tri_delta(j4, j5, j3)) * sum;
}
__device__ double wigner9j(int j11, int j12, int j13, int j21, int j22, int j23, int j31, int j32, int j33) {
  if(j11 < 0 || j12 < 0 || j13 < 0 || j21 < 0 || j22 < 0 || j23 < 0 || j31 < 0 || j32 < 0 || j33 < 0)
    return 0;
  if(tri_bad(j11, j12, j13) || tri_bad(j21, j22, j23) || tri_bad(j31, j32, j33) || tri_bad(j11, j21, j31) ||
     tri_bad(j12, j22, j32) || tri_bad(j13, j23, j33))
    return 0;
  int kmin = max(abs(j11 - j33), max(abs(j32 - j21), abs(j23 - j12)));
  int kmax = min(j11 + j33, min(j32 + j21, j23 + j12));
  double sum = 0;
  for(int k = kmin; k <= kmax; ++k)
    sum += (double)(2 * k + 1) * wigner6j(j11, j21, j31, j32, j33, k) * wigner6j(j12, j22, j32, j21, k, j23) *
           wigner6j(j13, j23, j33, k, j11, j12);
  return sum;
}
// Symbol.cpp:44-48
__device__ double clegor(int j, int m, int j1, int m1, int j2, int m2) {
  double s = ((m + j1 - j2) & 1) ? -1.0 : 1.0;
  return s * sqrt(2.0 * j + 1.0) * wigner3j(j1, j2, j, m1, m2, -m);
}
// m-independent part of Symbol.cpp:143-150
__device__ double wred(int L1, int J1, int L2, int J2, int L) {
  if(L1 < 0 || L2 < 0)
    return 0;
  double s = ((J2 + L1 + L) & 1) ? -1.0 : 1.0;
  return s *
         sqrt((2.0 * J1 + 1.0) * (2.0 * J2 + 1.0) * (2.0 * L1 + 1.0) * (2.0 * L2 + 1.0) /
              (4.0 * 3.14159265358979323846 * (2.0 * L + 1.0))) *
         wigner6j(L1, L2, L, J2, J1, 1) * clegor(L, 0, L1, 0, L2, 0);
}

// red[t][J][J1][J2], J in 1..nMaxS, J1,J2 in 1..nMax (index (J-1)*nMax*nMax + (J1-1)*nMax + (J2-1))
__global__ void k_cg_reduced(int nMax, int nMaxS, double *__restrict__ red) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = nMaxS * nMax * nMax;
  if(idx >= total)
    return;
  const int J = idx / (nMax * nMax) + 1;
  const int J1 = (idx / nMax) % nMax + 1;
  const int J2 = idx % nMax + 1;
  const double s32 = sqrt(3.0 / 2.0 / 3.14159265358979323846);
  const double dJ = J, dJ1 = J1, dJ2 = J2;
  const double sJp = sqrt(dJ / (2.0 * dJ + 1.0)), sJm = sqrt((dJ + 1.0) / (2.0 * dJ + 1.0));
  double r[9];
  // C_10m1 (Symbol.cpp:1058-1075)
  r[0] = s32 * (2.0 * dJ1 + 1.0) *
         (sqrt(dJ2 * (2.0 * dJ2 - 1.0)) * wigner9j(J1, J1, 1, J2, J2 - 1, 1, J, J + 1, 1) *
              clegor(J + 1, 0, J1, 0, J2 - 1, 0) * sJp -
          sqrt((dJ2 + 1.0) * (2.0 * dJ2 + 3.0)) * wigner9j(J1, J1, 1, J2, J2 + 1, 1, J, J + 1, 1) *
              clegor(J + 1, 0, J1, 0, J2 + 1, 0) * sJp +
          sqrt(dJ2 * (2.0 * dJ2 - 1.0)) * wigner9j(J1, J1, 1, J2, J2 - 1, 1, J, J - 1, 1) *
              clegor(J - 1, 0, J1, 0, J2 - 1, 0) * sJm -
          sqrt((dJ2 + 1.0) * (2.0 * dJ2 + 3.0)) * wigner9j(J1, J1, 1, J2, J2 + 1, 1, J, J - 1, 1) *
              clegor(J - 1, 0, J1, 0, J2 + 1, 0) * sJm);
  // C_11m1 (Symbol.cpp:1108-1142)
  {
    const double q1 = sqrt((dJ1 + 1.0) * dJ2 * (2.0 * dJ1 - 1.0) * (2.0 * dJ2 - 1.0));
    const double q2 = sqrt((dJ1 + 1.0) * (dJ2 + 1.0) * (2.0 * dJ1 - 1.0) * (2.0 * dJ2 + 3.0));
    const double q3 = sqrt(dJ1 * dJ2 * (2.0 * dJ1 + 3.0) * (2.0 * dJ2 - 1.0));
    const double q4 = sqrt(dJ1 * (dJ2 + 1.0) * (2.0 * dJ1 + 3.0) * (2.0 * dJ2 + 3.0));
    double acc = 0;
    for(int s = 0; s < 2; ++s) {
      const int Jx = s == 0 ? J + 1 : J - 1;
      const double sj = s == 0 ? sJp : sJm;
      acc += (q1 * wigner9j(J1, J1 - 1, 1, J2, J2 - 1, 1, J, Jx, 1) * clegor(Jx, 0, J1 - 1, 0, J2 - 1, 0) -
              q2 * wigner9j(J1, J1 - 1, 1, J2, J2 + 1, 1, J, Jx, 1) * clegor(Jx, 0, J1 - 1, 0, J2 + 1, 0) +
              q3 * wigner9j(J1, J1 + 1, 1, J2, J2 - 1, 1, J, Jx, 1) * clegor(Jx, 0, J1 + 1, 0, J2 - 1, 0) -
              q4 * wigner9j(J1, J1 + 1, 1, J2, J2 + 1, 1, J, Jx, 1) * clegor(Jx, 0, J1 + 1, 0, J2 + 1, 0)) *
             sj;
    }
    r[1] = s32 * acc;
    // C_01m1 (Symbol.cpp:1213-1226)
    r[3] = s32 * (q1 * wigner9j(J1, J1 - 1, 1, J2, J2 - 1, 1, J, J, 1) * clegor(J, 0, J1 - 1, 0, J2 - 1, 0) -
                  q2 * wigner9j(J1, J1 - 1, 1, J2, J2 + 1, 1, J, J, 1) * clegor(J, 0, J1 - 1, 0, J2 + 1, 0) +
                  q3 * wigner9j(J1, J1 + 1, 1, J2, J2 - 1, 1, J, J, 1) * clegor(J, 0, J1 + 1, 0, J2 - 1, 0) -
                  q4 * wigner9j(J1, J1 + 1, 1, J2, J2 + 1, 1, J, J, 1) * clegor(J, 0, J1 + 1, 0, J2 + 1, 0));
  }
  // C_00m1 (Symbol.cpp:1173-1180)
  r[2] = s32 * (2.0 * dJ1 + 1.0) *
         (sqrt(dJ2 * (2.0 * dJ2 - 1.0)) * wigner9j(J1, J1, 1, J2, J2 - 1, 1, J, J, 1) * clegor(J, 0, J1, 0, J2 - 1, 0) -
          sqrt((dJ2 + 1.0) * (2.0 * dJ2 + 3.0)) * wigner9j(J1, J1, 1, J2, J2 + 1, 1, J, J, 1) *
              clegor(J, 0, J1, 0, J2 + 1, 0));
  // W tables (Symbol.cpp:1261-1276, 1312-1327, 1361, 1396-1401, 1433-1438)
  const double a1 = sqrt(dJ1 / (2.0 * dJ1 + 1.0)), b1 = sqrt((dJ1 + 1.0) / (2.0 * dJ1 + 1.0));
  const double a2 = sqrt(dJ2 / (2.0 * dJ2 + 1.0)), b2 = sqrt((dJ2 + 1.0) / (2.0 * dJ2 + 1.0));
  const double wmm = wred(J1 - 1, J1, J2 - 1, J2, J), wpp = wred(J1 + 1, J1, J2 + 1, J2, J);
  const double wmp = wred(J1 - 1, J1, J2 + 1, J2, J), wpm = wred(J1 + 1, J1, J2 - 1, J2, J);
  r[4] = a1 * a2 * wmm + b1 * b2 * wpp - a1 * b2 * wmp - b1 * a2 * wpm;
  r[5] = b1 * b2 * wmm + a1 * a2 * wpp + b1 * a2 * wmp + a1 * b2 * wpm;
  r[6] = wred(J1, J1, J2, J2, J);
  r[7] = b1 * wred(J1 - 1, J1, J2, J2, J) + a1 * wred(J1 + 1, J1, J2, J2, J);
  r[8] = b2 * wred(J1, J1, J2 - 1, J2, J) + a2 * wred(J1, J1, J2 + 1, J2, J);
  for(int t = 0; t < 9; ++t)
    red[(size_t)t * total + idx] = r[t];
}

// dense tables in the reference layout [k n^2 + p n + q] = CG(J M|J1 M1 J2 M2) * red_t(J,J1,J2)
__global__ void k_cg_fill(int nMax, int nMaxS, const double *__restrict__ red, double *t0, double *t1, double *t2,
                          double *t3, double *t4, double *t5, double *t6, double *t7, double *t8) {
  const int n = flat_max(nMax), ns = flat_max(nMaxS);
  const size_t total = (size_t)ns * n * n;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= total)
    return;
  const int q = (int)(idx % n), p = (int)((idx / n) % n), k = (int)(idx / ((size_t)n * n));
  int J, M, J1, M1, J2, M2;
  unflatten(k, J, M);
  unflatten(p, J1, M1);
  unflatten(q, J2, M2);
  double *T[9] = {t0, t1, t2, t3, t4, t5, t6, t7, t8};
  if(M1 + M2 != M) {
    for(int t = 0; t < 9; ++t)
      T[t][idx] = 0.0;
    return;
  }
  const double cg = clegor(J, M, J1, M1, J2, M2);
  const int rtotal = nMaxS * nMax * nMax;
  const int ridx = (J - 1) * nMax * nMax + (J1 - 1) * nMax + (J2 - 1);
  for(int t = 0; t < 9; ++t)
    T[t][idx] = cg * red[(size_t)t * rtotal + ridx];
}

void launch_cg_tables(int nMax, int nMaxS, double *const T[9], cudaStream_t st) {
  upload_factorials();
  const int rtotal = nMaxS * nMax * nMax;
  double *red = nullptr;
  OB_CUDA(cudaMalloc(&red, (size_t)9 * rtotal * sizeof(double)));
  k_cg_reduced<<<(rtotal + 63) / 64, 64, 0, st>>>(nMax, nMaxS, red);
  OB_CUDA(cudaGetLastError());
  const size_t total = (size_t)flat_max(nMaxS) * flat_max(nMax) * flat_max(nMax);
  k_cg_fill<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(nMax, nMaxS, red, T[0], T[1], T[2], T[3], T[4], T[5],
                                                            T[6], T[7], T[8]);
  OB_CUDA(cudaGetLastError());
  OB_CUDA(cudaStreamSynchronize(st));
  cudaFree(red);
}

// ---------------------------------------------------------------------------------------------
// K4: SH source vectors.  One CTA per particle; a warp owns one SH harmonic kk at a time, lanes run
// over p, each lane visits the <= nMax q's with M2 = M - M1.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ cplx warp_sum(cplx v) {
  for(int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_down_sync(0xffffffffu, v.x, o);
    v.y += __shfl_down_sync(0xffffffffu, v.y, o);
  }
  return v;
}

// SH_P particles per CTA: every table entry (nine scattered 8-byte reads per (k, p, q), the L2-bound part of the kernel)
// is loaded once and applied to all of them
#define SH_P 4
__global__ void __launch_bounds__(512)
k_sh_source(ShInputs in, int j0, int count, const cplx *__restrict__ Xint_conj, const cplx *__restrict__ TSH1o,
            const cplx *__restrict__ TSH2o, const cplx *__restrict__ IauxSH2, cplx *__restrict__ K,
            cplx *__restrict__ K1ana) {
  __shared__ cplx A0[SH_P][OB_MAX_FLAT], A1[SH_P][OB_MAX_FLAT], Am1[SH_P][OB_MAX_FLAT];
  __shared__ cplx fa0[SH_P][OB_MAX_NMAX + 2], fa1[SH_P][OB_MAX_NMAX + 2], fam1[SH_P][OB_MAX_NMAX + 2];
  const int jb = j0 + blockIdx.x * SH_P;           // first particle of this CTA
  const int np = min(SH_P, j0 + count - jb);       // particles of this CTA (the last CTA may hold fewer)
  const int nMax = in.nMax, n = flat_max(nMax), ns = flat_max(in.nMaxS);
  if(threadIdx.x < SH_P) {
    const int j = jb + min((int)threadIdx.x, np - 1), pp = threadIdx.x;
    const double R = in.radius[j];
    const cplx waveK_j1 = cscale(csqrt_(cmul(in.eps[j], in.mu[j])), in.omega);
    // per-order prefactors of A_0, A_1, A_m1 (Symbol.cpp:52-78)
    cplx d[OB_MAX_NMAX + 2], dd[OB_MAX_NMAX + 2];
    cplx z = cscale(waveK_j1, R);
    sph_bessel_j(z, nMax + 1, d);
    cplx iz = cdiv(mk(1, 0), z);
    for(int i = 0; i <= nMax; ++i)
      dd[i] = csub(cscale(cmul(iz, d[i]), (double)i), d[i + 1]);
    cplx ik = cdiv(mk(1, 0), waveK_j1);
    for(int i = 0; i <= nMax; ++i) {
      fa0[pp][i] = d[i];
      fa1[pp][i] = cmuli(cmul(ik, cadd(cmul(waveK_j1, dd[i]), cscale(d[i], 1.0 / R))));
      fam1[pp][i] = cmuli(cscale(cmul(cscale(ik, 1.0 / R), d[i]), sqrt((double)i * (i + 1.0))));
    }
  }
  __syncthreads();
  for(int e = threadIdx.x; e < SH_P * n; e += blockDim.x) {
    const int pp = e / n, p = e - pp * n, j = jb + min(pp, np - 1);
    int l, m;
    unflatten(p, l, m);
    cplx c = Xint_conj[(size_t)j * 2 * n + p], d = Xint_conj[(size_t)j * 2 * n + n + p];
    A0[pp][p] = cmul(fa0[pp][l], c);
    A1[pp][p] = cmul(fa1[pp][l], d);
    Am1[pp][p] = cmul(fam1[pp][l], d);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const double mu0 = 4.0 * 3.14159265358979323846 * 1e-7;
  const double eps0 = 1.0 / (mu0 * 299792458.0 * 299792458.0);
  const cplx eta_ratio = cdiv(csqrt_(cdiv(in.mu_b, in.eps_b)), csqrt_(mk(mu0 / eps0, 0))); // sqrt(mu_b/eps_b)/sqrt(mu_0/eps_0)
  const cplx waveK_01 = mk(in.omega * sqrt(eps0 * mu0), 0);
  for(int kk = warp; kk < ns; kk += nwarps) {
    int J, M;
    unflatten(kk, J, M);
    const size_t kbase = (size_t)kk * n * n;
    cplx sum_v[SH_P], sum_u[SH_P], gmn[SH_P], fmn[SH_P];
#pragma unroll
    for(int pp = 0; pp < SH_P; ++pp)
      sum_v[pp] = sum_u[pp] = gmn[pp] = fmn[pp] = mk(0, 0);
    for(int p = lane; p < n; p += 32) {
      int J1, M1;
      unflatten(p, J1, M1);
      const int M2 = M - M1;
      const int aM2 = M2 < 0 ? -M2 : M2;
      for(int J2 = max(1, aM2); J2 <= nMax; ++J2) {
        const int q = flat_index(J2, M2);
        const size_t t = kbase + (size_t)p * n + q;
        // v' (Symbol.cpp:256-260), u' (:199-203), u'' (Symbol.cpp:322-336)
        const double c00 = __ldg(in.tab[2] + t), c01 = __ldg(in.tab[3] + t);
        const double c10 = __ldg(in.tab[0] + t), c11 = __ldg(in.tab[1] + t);
        const double wm = __ldg(in.tab[4] + t);
        const double w11 = __ldg(in.tab[5] + t), w00 = __ldg(in.tab[6] + t), w10 = __ldg(in.tab[7] + t),
                     w01 = __ldg(in.tab[8] + t);
#pragma unroll
        for(int pp = 0; pp < SH_P; ++pp) {
          const cplx a0p = A0[pp][p], a1p = A1[pp][p], am1p = Am1[pp][p];
          const cplx a0q = A0[pp][q], a1q = A1[pp][q], am1q = Am1[pp][q];
          const cplx a1p_am1q = cmul(a1p, am1q), a0p_am1q = cmul(a0p, am1q);
          sum_v[pp].x += a1p_am1q.x * c00 + a0p_am1q.x * c01;
          sum_v[pp].y += a1p_am1q.y * c00 + a0p_am1q.y * c01;
          sum_u[pp].x += a1p_am1q.x * c10 + a0p_am1q.x * c11;
          sum_u[pp].y += a1p_am1q.y * c10 + a0p_am1q.y * c11;
          const cplx mm = cmul(am1p, am1q);
          gmn[pp].x += mm.x * wm;
          gmn[pp].y += mm.y * wm;
          const cplx t11 = cmul(a1p, a1q), t00 = cmul(a0p, a0q), t10 = cmul(a1p, a0q), t01 = cmul(a0p, a1q);
          fmn[pp].x += t11.x * w11 + t00.x * w00 + t10.x * w10 + t01.x * w01;
          fmn[pp].y += t11.y * w11 + t00.y * w00 + t10.y * w10 + t01.y * w01;
        }
      }
    }
#pragma unroll
    for(int pp = 0; pp < SH_P; ++pp) {
      const cplx sv = warp_sum(sum_v[pp]), su = warp_sum(sum_u[pp]), gm = warp_sum(gmn[pp]), fm = warp_sum(fmn[pp]);
      if(lane == 0 && pp < np) {
        const int j = jb + pp;
        const double R = in.radius[j];
        const cplx ksiparppar = in.ksiparppar[j], ksippp = in.ksippp[j], gamma = in.gamma[j];
        // Symbol.cpp:266-267 / :210-211
        cplx vp = cmul(cmul(cscale(sv, 2.0), ksiparppar), eta_ratio);
        cplx up = cmul(cmul(cmuli(cscale(su, 2.0)), ksiparppar), eta_ratio);
        // Symbol.cpp:347-351
        const double sq = sqrt((double)(J * (J + 1)));
        cplx inv = cdiv(mk(1.0, 0), cscale(waveK_01, R));
        cplx term1 = cmul(cmuli(cscale(cmul(ksippp, gm), sq)), inv);
        cplx ge = cmul(gamma, cdiv(mk(eps0, 0), in.eps_SH[j]));
        cplx term2 = cmul(cmuli(cscale(cmul(ge, cadd(gm, fm)), sq)), inv);
        cplx upp = cadd(term1, term2);
        // PreconditionedMatrix.cpp:1381-1383 and :1424-1426 (v'' == 0, Geometry.cpp:296)
        const size_t o = (size_t)j * 2 * ns;
        K[o + kk] = cmul(TSH1o[o + kk], vp);
        K[o + ns + kk] = cadd(cmul(TSH1o[o + ns + kk], up), cmul(TSH2o[o + ns + kk], upp));
        K1ana[o + kk] = mk(0, 0);
        K1ana[o + ns + kk] = cmul(IauxSH2[o + ns + kk], upp);
      }
    }
  }
}

void launch_sh_source(ShInputs const &in, int j0, int count, const cplx *Xint_conj, const cplx *TSH1o,
                      const cplx *TSH2o, const cplx *IauxSH2, cplx *K, cplx *K1ana, cudaStream_t st) {
  if(count <= 0)
    return;
  k_sh_source<<<(count + SH_P - 1) / SH_P, 512, 0, st>>>(in, j0, count, Xint_conj, TSH1o, TSH2o, IauxSH2, K, K1ana);
  OB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// K6b: SH absorption cross section, per particle  sum_kk ACSshcoeff  (Symbol.cpp:358-477)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
k_abs_sh(ShInputs in, int j0, const cplx *__restrict__ Xint, const cplx *__restrict__ Xint_SH,
         cplx *__restrict__ out) {
  // per Gauss radius and order: j, j', g = k j' + j/r, h = k^2 j'' - j/r^2 + k j'/r
  __shared__ cplx fj[4][OB_MAX_NMAX + 2], fd[4][OB_MAX_NMAX + 2], fg[4][OB_MAX_NMAX + 2], fh[4][OB_MAX_NMAX + 2];
  __shared__ cplx sj[4][OB_MAX_NMAX + 2], sd[4][OB_MAX_NMAX + 2];
  __shared__ cplx cs[OB_MAX_FLAT], ds[OB_MAX_FLAT];
  __shared__ double wre[16], wim[16];
  const int j = j0 + blockIdx.x;
  const int nMax = in.nMax, nMaxS = in.nMaxS, n = flat_max(nMax), ns = flat_max(nMaxS);
  const double R = in.radius[j];
  const double mu0 = 4.0 * 3.14159265358979323846 * 1e-7;
  const double eps0 = 1.0 / (mu0 * 299792458.0 * 299792458.0);
  const cplx waveK_j1 = cscale(csqrt_(cmul(in.eps[j], in.mu[j])), in.omega);
  const cplx waveK_SH = cscale(csqrt_(cmul(in.eps_SH[j], in.mu_SH[j])), 2.0 * in.omega);
  const cplx waveK_01 = mk(in.omega * sqrt(eps0 * mu0), 0);
  const double xi[4] = {-0.3399810435848563, 0.3399810435848563, -0.8611363115940526, 0.8611363115940526};
  const double wi[4] = {0.6521451548625461, 0.6521451548625461, 0.3478548451374538, 0.3478548451374538};
  if(threadIdx.x < 4) {
    const int ii = threadIdx.x;
    const double r = (R / 2.0) * xi[ii] + R / 2.0;
    cplx d[OB_MAX_NMAX + 3], dd[OB_MAX_NMAX + 2];
    cplx z = cscale(waveK_j1, r);
    sph_bessel_j(z, nMax + 1, d);
    cplx iz = cdiv(mk(1, 0), z);
    for(int i = 0; i <= nMax; ++i)
      dd[i] = csub(cscale(cmul(iz, d[i]), (double)i), d[i + 1]);
    const cplx k2 = cmul(waveK_j1, waveK_j1);
    for(int i = 0; i <= nMax; ++i) {
      fj[ii][i] = d[i];
      fd[ii][i] = dd[i];
      fg[ii][i] = cadd(cmul(waveK_j1, dd[i]), cscale(d[i], 1.0 / r));
      // Bessel.h:227-228: j''_i = -(i/z) j'_i + (i/z^2) j_i + j'_{i-1}
      cplx ddd = mk(0, 0);
      if(i >= 1)
        ddd = cadd(cadd(cneg(cscale(cmul(iz, dd[i]), (double)i)), cscale(cmul(cmul(iz, iz), d[i]), (double)i)), dd[i - 1]);
      fh[ii][i] = cadd(csub(cmul(k2, ddd), cscale(d[i], 1.0 / (r * r))), cscale(cmul(waveK_j1, dd[i]), 1.0 / r));
    }
    cplx zs = cscale(waveK_SH, r);
    sph_bessel_j(zs, nMaxS + 1, d);
    cplx izs = cdiv(mk(1, 0), zs);
    for(int i = 0; i <= nMaxS; ++i) {
      sj[ii][i] = d[i];
      sd[ii][i] = csub(cscale(cmul(izs, d[i]), (double)i), d[i + 1]);
    }
  }
  for(int p = threadIdx.x; p < n; p += blockDim.x) {
    cs[p] = Xint[(size_t)j * 2 * n + p];
    ds[p] = Xint[(size_t)j * 2 * n + n + p];
  }
  __syncthreads();
  // radial products of Symbol.cpp:80-141 depend on (Gauss point, n1, n2) only: tabulated once per particle
  // (the reference -- and the first version of this kernel -- re-evaluates them, with their divisions and the
  // sqrt(n1 n2 (n1+1)(n2+1)), inside the (k, p, q) loops).  Layout [ii][n1][n2][6], orders 1..nMax.
  extern __shared__ __align__(16) unsigned char ftab_raw[];
  cplx *ftab = (cplx *)ftab_raw;
  for(int e = threadIdx.x; e < 4 * nMax * nMax; e += blockDim.x) {
    const int ii = e / (nMax * nMax), rem = e - ii * nMax * nMax, n1 = rem / nMax + 1, n2 = rem - (n1 - 1) * nMax + 1;
    const double r = (R / 2.0) * xi[ii] + R / 2.0;
    const cplx j1 = fj[ii][n1], j2 = fj[ii][n2], e1 = fd[ii][n1], e2 = fd[ii][n2];
    const cplx g1 = fg[ii][n1], g2 = fg[ii][n2], h1 = fh[ii][n1], h2 = fh[ii][n2];
    const double sq12 = sqrt((double)(n1 * n2 * (n1 + 1) * (n2 + 1)));
    const cplx F_00 = cmul(j1, e2);
    const cplx F_11 = cmul(g2, g1);
    const cplx j1j2 = cmul(j1, j2);
    const cplx F_m1m1 = cscale(j1j2, 1.0 / (r * r));
    const cplx sym = cadd(cmul(j1, e2), cmul(e1, j2));
    const cplx F_d00 = cmul(waveK_j1, sym);
    const cplx F_d11 = cadd(cmul(h1, g2), cmul(h2, g1));
    const cplx F_dm1m1 = csub(cscale(cmul(waveK_j1, sym), 1.0 / (r * r)), cscale(j1j2, 2.0 / (r * r * r)));
    cplx *o = ftab + (size_t)e * 6;
    o[0] = F_d00;
    o[1] = F_d11;
    o[2] = cscale(F_dm1m1, sq12);
    o[3] = F_00;
    o[4] = F_11;
    o[5] = cscale(F_m1m1, sq12);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const cplx pref = cmul(cdiv(mk(-eps0, 0), in.eps_SH[j]), in.gamma[j]); // (-eps_0/eps_j2) gamma
  const cplx ik2 = cdiv(mk(1, 0), cmul(waveK_j1, waveK_j1));
  cplx total = mk(0, 0);
  for(int kk = warp; kk < ns; kk += nwarps) {
    int J, M;
    unflatten(kk, J, M);
    const size_t kbase = (size_t)kk * n * n;
    const double sqJ = sqrt((double)(J * (J + 1)));
    cplx Xm1[4], Xp1[4];
    for(int ii = 0; ii < 4; ++ii)
      Xm1[ii] = Xp1[ii] = mk(0, 0);
    for(int p = lane; p < n; p += 32) {
      int n1, M1;
      unflatten(p, n1, M1);
      const int M2 = M - M1, aM2 = M2 < 0 ? -M2 : M2;
      const cplx c1 = cs[p], d1 = ds[p];
      for(int n2 = max(1, aM2); n2 <= nMax; ++n2) {
        const int q = flat_index(n2, M2);
        const size_t t = kbase + (size_t)p * n + q;
        const double Wm1m1 = __ldg(in.tab[4] + t), W11 = __ldg(in.tab[5] + t), W00 = __ldg(in.tab[6] + t);
        const cplx cc = cscale(cmul(c1, cs[q]), W00);
        const cplx ddk = cmul(cmul(d1, ds[q]), ik2);
#pragma unroll
        for(int ii = 0; ii < 4; ++ii) {
          const cplx *f = ftab + ((size_t)(ii * nMax + (n1 - 1)) * nMax + (n2 - 1)) * 6;
          // Symbol.cpp:440-451 (the common factor sqrt(J(J+1))/r of the +1 component is applied after the sum)
          cfma(Xm1[ii], cc, f[0]);
          cfma(Xm1[ii], ddk, cadd(cscale(f[1], W11), cscale(f[2], Wm1m1)));
          cfma(Xp1[ii], cc, f[3]);
          cfma(Xp1[ii], ddk, cadd(cscale(f[4], W11), cscale(f[5], Wm1m1)));
        }
      }
    }
    for(int ii = 0; ii < 4; ++ii) {
      const double r = (R / 2.0) * xi[ii] + R / 2.0;
      Xp1[ii] = cscale(Xp1[ii], sqJ / r);
    }
    cplx integral = mk(0, 0);
    const cplx cmnSH = Xint_SH[(size_t)j * 2 * ns + kk], dmnSH = Xint_SH[(size_t)j * 2 * ns + ns + kk];
    for(int ii = 0; ii < 4; ++ii) {
      cplx xm = cmul(pref, warp_sum(Xm1[ii]));
      cplx xp = cmul(pref, warp_sum(Xp1[ii]));
      if(lane == 0) {
        const double r = (R / 2.0) * xi[ii] + R / 2.0;
        // Symbol.cpp:460-464
        const cplx kd = cmul(cmul(waveK_01, dmnSH), cdiv(mk(1, 0), waveK_SH));
        const cplx Xm1SH = cscale(cmul(kd, sj[ii][J]), sqJ / r);
        const cplx X0 = cneg(cmul(cmul(waveK_01, cmnSH), sj[ii][J]));
        const cplx Xp1SH = cmul(kd, cadd(cscale(sj[ii][J], 1.0 / r), sd[ii][J]));
        const cplx s1 = cadd(xm, Xm1SH), s2 = cadd(xp, Xp1SH);
        // Symbol.cpp:467-469 (products z * conj(z) are real)
        integral.x += wi[ii] * r * r * (cnorm(s1) + cnorm(X0) + cnorm(s2));
      }
    }
    if(lane == 0)
      total = cadd(total, cscale(integral, R / 2.0));
  }
  if(lane == 0) {
    wre[warp] = total.x;
    wim[warp] = total.y;
  }
  __syncthreads();
  if(threadIdx.x == 0) {
    cplx s = mk(0, 0);
    for(int w = 0; w < nwarps; ++w)
      s = cadd(s, mk(wre[w], wim[w]));
    out[blockIdx.x] = s;
  }
}

void launch_abs_sh(ShInputs const &in, int j0, int count, const cplx *Xint, const cplx *Xint_SH, cplx *out,
                   cudaStream_t st) {
  if(count <= 0)
    return;
  const size_t ftab_bytes = (size_t)4 * in.nMax * in.nMax * 6 * sizeof(cplx);
  // static (~12 KB) + dynamic shared memory crosses the 48 KB default already at nMax = 10: always opt in
  OB_CUDA(cudaFuncSetAttribute((const void *)k_abs_sh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ftab_bytes));
  k_abs_sh<<<count, 512, ftab_bytes, st>>>(in, j0, Xint, Xint_SH, out);
  OB_CUDA(cudaGetLastError());
}

} // namespace ob
