// ob_mie.cu -- K0: per-sphere Mie factors on device (complex-argument spherical Bessel/Hankel).
//   reference: srcAna/Scatterer.cpp:39-99 (getTLocal), :101-161 (getTLocalSH), :163-271
//   (getTLocalSH1_outer / getTLocalSH2_outer), :274-412 (getIaux, getIauxSH1, getIauxSH2).
// One thread per (particle, harmonic); every value is replicated over the 2n+1 m's of its order,
// layout [TE(n) ; TM(n)] with n = nMax(nMax+2) per half.
#include "ob_internal.h"
#include "ob_special.cuh"

namespace ob {

struct RB {
  cplx psi, dpsi, ksi, dksi, psirho, dpsirho;
};

// spherical functions and derivatives z'_i = -z_{i+1} + (i/z) z_i  (Bessel.h:125-130)
__device__ static void bessel_with_derivative(bool hankel, cplx z, int nmax, cplx *d, cplx *dd) {
  // d, dd: nmax+2 entries of scratch (orders 0..nmax+1 for d)
  if(cabs_(z) <= 1e-10) { // Bessel.h:71-75
    for(int i = 0; i <= nmax + 1; ++i)
      d[i] = dd[i] = mk(0, 0);
    if(!hankel)
      d[0] = mk(1, 0);
    return;
  }
  if(hankel)
    sph_hankel1(z, nmax + 1, d);
  else
    sph_bessel_j(z, nmax + 1, d);
  cplx iz = cdiv(mk(1, 0), z);
  for(int i = 0; i <= nmax; ++i)
    dd[i] = csub(cscale(cmul(iz, d[i]), (double)i), d[i + 1]);
}

__global__ void k_mie(MieInputs in, cplx *o0, cplx *o1, cplx *o2, cplx *o3, cplx *o4, cplx *o5, cplx *o6) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= 2 * in.nobj)
    return;
  const int j = t >> 1, harmonic = (t & 1) + 1;
  const int nm = harmonic == 1 ? in.nMax : in.nMaxS;
  const int N = flat_max(nm);
  const double radius = in.radius[j];
  cplx eps = harmonic == 1 ? in.eps[j] : in.eps_SH[j];
  cplx mu = harmonic == 1 ? in.mu[j] : in.mu_SH[j];
  const double w = harmonic == 1 ? in.omega : 2.0 * in.omega;
  cplx k_s = cscale(csqrt_(cmul(eps, mu)), w);
  cplx k_b = cscale(csqrt_(cmul(in.eps_b, in.mu_b)), w);
  cplx rho = cdiv(k_s, k_b);
  cplx r_0 = cscale(k_b, radius);
  cplx mu_sob = cdiv(mu, in.mu_b);
  cplx J[OB_MAX_NMAX + 2], dJ[OB_MAX_NMAX + 2], Jr[OB_MAX_NMAX + 2], dJr[OB_MAX_NMAX + 2], H[OB_MAX_NMAX + 2],
      dH[OB_MAX_NMAX + 2];
  bessel_with_derivative(false, r_0, nm, J, dJ);
  bessel_with_derivative(false, cmul(rho, r_0), nm, Jr, dJr);
  bessel_with_derivative(true, r_0, nm, H, dH);
  cplx x_b2 = r_0;                // k_b_SH * radius
  cplx x_i2 = cscale(k_s, radius); // k_s_SH * radius
  cplx zeta_boj2 = cdiv(csqrt_(cdiv(in.mu_b, in.eps_b)), csqrt_(cdiv(mu, eps)));
  cplx I = mk(0, 1);
  for(int n = 1, cur = 0; n <= nm; cur += 2 * n + 1, ++n) {
    RB b;
    b.psi = cmul(r_0, J[n]);
    b.dpsi = cadd(cmul(r_0, dJ[n]), J[n]);
    b.ksi = cmul(r_0, H[n]);
    b.dksi = cadd(cmul(r_0, dH[n]), H[n]);
    cplx rr = cmul(r_0, rho);
    b.psirho = cmul(rr, Jr[n]);
    b.dpsirho = cadd(cmul(rr, dJr[n]), Jr[n]);
    cplx v[5][2]; // up to five output families for this harmonic
    int nout = 0;
    cplx *outs[5];
    // T (Scatterer.cpp:83-90 / :139-146)
    {
      cplx q1 = cdiv(b.dpsi, b.psi), q2 = cdiv(b.dpsirho, b.psirho), q3 = cdiv(b.dksi, b.ksi);
      cplx pk = cdiv(b.psi, b.ksi);
      cplx TE = cdiv(cmul(pk, csub(cmul(mu_sob, q1), cmul(rho, q2))), csub(cmul(rho, q2), cmul(mu_sob, q3)));
      cplx TM = cdiv(cmul(pk, csub(cmul(mu_sob, q2), cmul(rho, q1))), csub(cmul(rho, q3), cmul(mu_sob, q2)));
      v[nout][0] = TE;
      v[nout][1] = TM;
      outs[nout++] = harmonic == 1 ? o0 : o1;
    }
    if(harmonic == 1) {
      // Iaux (Scatterer.cpp:305-310)
      cplx mu_j = mu, mu_0 = in.mu_b;
      cplx num = cmul(mu_j, rho);
      cplx dTE = csub(cmul(cmul(cmul(mu_0, rho), b.dpsirho), b.psi), cmul(cmul(mu_j, b.psirho), b.dpsi));
      cplx dTM = csub(cmul(cmul(mu_j, b.psi), b.dpsirho), cmul(cmul(cmul(mu_0, rho), b.psirho), b.dpsi));
      v[nout][0] = cmul(cdiv(num, dTE), I);
      v[nout][1] = cmul(cdiv(num, dTM), I);
      outs[nout++] = o4;
    } else {
      cplx den1 = csub(cmul(cmul(zeta_boj2, b.ksi), b.dpsirho), cmul(b.psirho, b.dksi)); // zeta ksi dpsirho - psirho dksi
      cplx den2 = csub(cmul(cmul(zeta_boj2, b.psirho), b.dksi), cmul(b.ksi, b.dpsirho)); // zeta psirho dksi - ksi dpsirho
      // TSH1_outer (Scatterer.cpp:203, 208)
      v[nout][0] = cdiv(cneg(cmul(x_b2, b.psirho)), den1);
      v[nout][1] = cdiv(cneg(cmul(x_b2, b.dpsirho)), den2);
      outs[nout++] = o2;
      // TSH2_outer (Scatterer.cpp:259, 265)
      cplx bnpp = cdiv(cmul(cmul(zeta_boj2, x_b2), b.dpsirho), den1);
      cplx anpp = cdiv(cmul(cmul(zeta_boj2, x_b2), b.psirho), den2);
      v[nout][0] = bnpp;
      v[nout][1] = anpp;
      outs[nout++] = o3;
      // IauxSH1 (Scatterer.cpp:352-354)
      cplx e1 = cdiv(cneg(cmul(x_i2, b.ksi)), cmul(x_b2, b.psirho));
      cplx e2 = cdiv(cneg(cmul(x_i2, b.dksi)), cmul(x_b2, b.dpsirho));
      v[nout][0] = e1;
      v[nout][1] = e2;
      outs[nout++] = o5;
      // IauxSH2 (Scatterer.cpp:399-405)
      cplx f1 = cdiv(cmul(x_i2, b.dksi), cmul(cmul(zeta_boj2, x_b2), b.dpsirho));
      cplx f2 = cdiv(cmul(x_i2, b.ksi), cmul(cmul(zeta_boj2, x_b2), b.psirho));
      v[nout][0] = cmul(bnpp, cadd(e1, f1));
      v[nout][1] = cmul(anpp, cadd(e2, f2));
      outs[nout++] = o6;
    }
    for(int f = 0; f < nout; ++f) {
      cplx *dst = outs[f] + (size_t)j * 2 * N;
      for(int i = 0; i < 2 * n + 1; ++i) {
        dst[cur + i] = v[f][0];
        dst[cur + N + i] = v[f][1];
      }
    }
  }
}

void launch_mie(MieInputs const &in, cplx *const out[7], cudaStream_t st) {
  int threads = 2 * in.nobj;
  k_mie<<<(threads + 63) / 64, 64, 0, st>>>(in, out[0], out[1], out[2], out[3], out[4], out[5], out[6]);
  OB_CUDA(cudaGetLastError());
}

} // namespace ob
