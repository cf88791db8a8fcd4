// ob_internal.h -- internal declarations shared by the .cu translation units (not part of the C ABI).
#pragma once
#include "ob_common.cuh"
#include <cmath>
#include <complex>
#include <string>
#include <vector>

#define OB_VTAC_THREADS 384
#define OB_MAX_NMAX 13           // shared-memory limit of the VTAC level buffers (2 x T(nMax) x 16 B <= 227 KB)
#define OB_MAX_FLAT (13 * 15)    // nMax (nMax + 2) at OB_MAX_NMAX

#include "ob_vtac.cuh"

namespace ob {

// index-only VTAC tables for one nMax, resident on one device
struct VtacTableSet {
  int nMax = -1;
  size_t smem = 0;
  VtacTables tb;
  std::vector<void *> allocs;
  void build(int NM);
  void release();
};

// ---- ob_vtac.cu ----
void launch_assemble(VtacTableSet const &ts, const double *xyz, const cplx *Tdiag, cplx k, int nobj, int row0,
                     int nrows, cplx *S, size_t ld, cudaStream_t st);
void launch_vtac_single(VtacTableSet const &ts, double r, double the, double phi, cplx k, int regular, cplx *A, cplx *B,
                        cudaStream_t st);
void launch_translate_apply(VtacTableSet const &ts, const double *xyz, cplx k, int j0, int count, const cplx *x,
                            int x_stride, const cplx *scale, cplx *out, cudaStream_t st);
void launch_sca_sum(VtacTableSet const &ts, const double *xyz, cplx k, int j0, int count, const cplx *x, double *out,
                    cudaStream_t st);

// ---- ob_mie.cu ----
struct MieInputs {
  int nobj, nMax, nMaxS;
  double omega;
  cplx eps_b, mu_b;
  const double *radius;                       // device, nobj
  const cplx *eps, *mu, *eps_SH, *mu_SH;      // device, nobj each (absolute values)
};
// out[which]: 0 T_FF, 1 T_SH, 2 TSH1_outer, 3 TSH2_outer, 4 Iaux, 5 IauxSH1, 6 IauxSH2; each nobj x 2n
void launch_mie(MieInputs const &in, cplx *const out[7], cudaStream_t st);

// ---- ob_sh.cu ----
struct ShInputs {
  int nobj, nMax, nMaxS;
  double omega;
  cplx eps_b, mu_b;
  const double *radius;
  const cplx *eps, *mu, *eps_SH, *mu_SH, *ksippp, *ksiparppar, *gamma;
  const double *tab[9]; // C_10m1, C_11m1, C_00m1, C_01m1, W_m1m1, W_11, W_00, W_10, W_01
};
void launch_cg_tables(int nMax, int nMaxS, double *const T[9], cudaStream_t st);
void launch_sh_source(ShInputs const &in, int j0, int count, const cplx *Xint_conj, const cplx *TSH1o,
                      const cplx *TSH2o, const cplx *IauxSH2, cplx *K, cplx *K1ana, cudaStream_t st);
void launch_abs_sh(ShInputs const &in, int j0, int count, const cplx *Xint, const cplx *Xint_SH, cplx *out,
                   cudaStream_t st);

// ---- ob_matvec.cu ----
struct MatvecPlan {
  int M = 0, N = 0;       // local rows, columns
  size_t ld = 0;
  int tiles = 0, chunks = 0, cols_per_chunk = 0, grid = 0;
  cplx *partial = nullptr; // [chunks][M] when chunks > 1
  int variant = 0;
};
void matvec_plan(MatvecPlan &p, int M, int N, size_t ld, int sm_count, int variant);
void matvec_plan_release(MatvecPlan &p);
void launch_matvec(MatvecPlan const &p, const cplx *S, const cplx *x, cplx *y, cudaStream_t st,
                   cudaEvent_t e0 = nullptr, cudaEvent_t e1 = nullptr);
size_t matvec_launches_per_apply(MatvecPlan const &p);

// ---- ob_pairs.cu (compact pair operator: unscaled A^T, B^T of the pairs i < j) ----
struct PairPlan {
  int nobj = 0, n = 0, grid = 0, nseg = 0;
  long p0 = 0, p1 = 0, npairs = 0; // global pair range [p0, p1) owned by this rank
  int KB = 0, NS = 0, G = 0;
  size_t smem = 0;
  int2 *pair_ij = nullptr;
  int4 *segs = nullptr;
  int *cta_seg = nullptr, *row_seg = nullptr;
  int2 *col_range = nullptr;
  long *col_first = nullptr;
  cplx *rowpart = nullptr, *colpart = nullptr, *XP = nullptr, *XS = nullptr, *acc = nullptr;
};
void pair_plan_build(PairPlan &p, int nobj, int n, int world, int rank, int sm_count);
void pair_plan_release(PairPlan &p);
void pair_plan_tuning(int kb, int groups);
size_t pair_storage_elems(PairPlan const &p);
void launch_matvec_pairs(PairPlan const &p, const cplx *AB, const cplx *x, const cplx *Tdiag, cplx *acc_or_y,
                         int finalize, cudaStream_t st, cudaEvent_t e0 = nullptr, cudaEvent_t e1 = nullptr,
                         bool x_staged = false);
void launch_pairs_reduce(PairPlan const &p, const cplx *x, const cplx *Tdiag, cplx *acc_or_y, int finalize,
                         cudaStream_t st);
void launch_pairs_finalize(const cplx *x, const cplx *Tdiag, const cplx *acc, size_t N, cplx *y, cudaStream_t st);
void launch_pairs_expand_block(PairPlan const &p, const cplx *AB, int i, int j, const cplx *Tdiag, cplx *out,
                               cudaStream_t st);
// ob_vtac.cu: assembly of the pair storage (one CTA per local pair)
void assemble_pairs_tuning(int minb); // 0 auto, 2 or 3 resident CTAs per SM
void launch_assemble_pairs(VtacTableSet const &ts, const double *xyz, cplx k, const int2 *pair_ij, long npairs,
                           cplx *AB, cudaStream_t st);

// ---- ob_rot.cu (rotated-axial operator: per pair phases, axial A+-B, flip-basis Wigner small-d; see the file header) ----
#define OB_ROT_MAX_THREADS 224 // ceil32(nMax (nMax + 2)) at OB_MAX_NMAX
struct RotLayout {
  int NM = 0, n = 0, X = 0, nDs = 0, nDa = 0, LF = 0, nh = 0; // nMax, harmonics, axial entries, small-d reals, F / channel lengths
  size_t offCp = 0, offCm = 0, offDs = 0, offDa = 0, rec_bytes = 0; // byte offsets inside one pair record (phases at 0)
};
struct RotPlan {
  int nobj = 0, n = 0, I = 0, grid = 0, nseg = 0, nblocks = 0, threads = 0, ctas_per_sm = 0;
  long npairs = 0, nstrips = 0, strip0 = 0; // local pairs / strips, global index of the first local strip
  size_t smem = 0;
  int2 *pair_ij = nullptr; // (i, j) of every local record (assembly)
  int4 *pinfo = nullptr;   // (i, j, local strip, flags) of every local record (apply)
  int *cta_pair = nullptr, *cta_seg = nullptr, *blk_seg = nullptr;
  long *blk_strip = nullptr;
  cplx *rowpart = nullptr, *colpart = nullptr, *acc = nullptr;
};
RotLayout rot_layout(int NM);
void rot_plan_build(RotPlan &p, int nobj, int NM, int world, int rank, int sm_count);
void rot_plan_release(RotPlan &p);
// -1 keeps a setting; assembly: 1 = axial-only recursion (default), 0 = vtac_block-based; rows per block; CTAs per SM
void rot_tuning(int assembly, int rows, int ctas_per_sm);
void launch_assemble_rot(VtacTableSet const &ts, const double *xyz, cplx k, const int2 *pair_ij, long npairs,
                         unsigned char *recs, RotLayout const &L, int sm_count, cudaStream_t st, bool geometry = true);
// geo: records holding the geometry sections (phases, small-d) of the same pairs; == recs when the harmonic owns them
void launch_matvec_rot(RotPlan const &p, RotLayout const &L, const unsigned char *recs, const unsigned char *geo, const cplx *x, const cplx *Tdiag,
                       cplx *acc_or_y, int finalize, cudaStream_t st, cudaEvent_t e0 = nullptr, cudaEvent_t e1 = nullptr);

// ---- ob_aca.cu (ACA-compressed operator: U V^T-style low-rank far blocks, dense near blocks) ----
struct AcaDesc {
  cplx *U = nullptr; // low rank: dim x rank column-major; dense: dim x dim column-major block
  cplx *V = nullptr; // low rank: rank rows of dim entries
  int rank = 0;      // > 0 low rank, -1 dense near block, 0 identity diagonal
  int pad = 0;
};
struct AcaScratch { // assembly scratch of one batch of block-rows: dense slab + full-rank U, V; kept by the context
  cplx *slab = nullptr, *scrU = nullptr, *scrV = nullptr;
  int2 *d_jobs = nullptr, *d_jobs_dense = nullptr;
  int *d_rank = nullptr;
  size_t elems = 0, blocks = 0;
  void release();
};
struct AcaOperator {
  int nobj = 0, dim = 0, first = 0, count = 0, nch = 1;
  bool built = false;
  std::vector<cplx *> chunks; // one allocation per assembly batch, sized from the ranks, reused across builds
  std::vector<size_t> chunk_cap;
  size_t partial_elems = 0, desc_blocks = 0;
  AcaDesc *desc = nullptr;    // device, [count][nobj]
  std::vector<AcaDesc> h_desc;
  int *piv = nullptr;         // device, [count][nobj][2][dim]: pivot rows I then pivot columns J
  cplx *partial = nullptr;    // [nch][count dim]
  double stored_elems = 0, rank_sum = 0;
  long n_lowrank = 0, n_dense = 0;
  int rank_max = 0;
  void release();
};
void aca_build(AcaOperator &op, AcaScratch &scr, VtacTableSet const &ts, const double *d_xyz, const cplx *Tdiag, cplx k, int nobj,
               int first, int count, const double *h_xyz, const double *h_radius, double eps, size_t budget_bytes,
               int sm_count, cudaStream_t st, long &launches);
void launch_matvec_aca(AcaOperator const &op, const cplx *x, cplx *y_slice, cudaStream_t st, cudaEvent_t e0 = nullptr,
                       cudaEvent_t e1 = nullptr);
void aca_compress_single(const cplx *C_dev, int dim, double eps, cplx *U_dev, cplx *V_dev, int *rank_dev, int *piv_dev,
                         cudaStream_t st);

// ---- ob_fields.cu (near-field maps: Result::getEHFields / setFields) ----
struct FieldInputs {
  int nobj, nMax, nMaxS, do_sh;
  double omega;
  cplx waveK, eps_b, mu_b;
  const double *xyz, *radius;                      // device
  const cplx *eps, *mu, *eps_SH, *mu_SH, *gamma;   // device, nobj each (absolute values)
  const cplx *ainc;                                // [a ; b] incident coefficients at the origin, 2n
  const cplx *Xsca, *Xint, *XscaSH, *XintSH;       // solution vectors, full length
  const double *tab[9];                            // CG tables (W_m1m1, W_11, W_00 are used)
};
// pts: npts x (r, theta, phi); out: npts x 4 x 3 complex (E_FF, H_FF, E_SH, H_SH; Cartesian); inner: npts
void launch_fields(FieldInputs const &in, long npts, const double *pts_dev, cplx *out_dev, int *inner_dev,
                   cudaStream_t st);

// ---- ob_lu.cu (device direct solve: blocked LU with partial pivoting, zgesv-style) ----
struct LuWork {
  int cap = 0;
  int *ipiv = nullptr, *pivrow = nullptr, *info = nullptr, *part_i = nullptr;
  unsigned *sync = nullptr;
  double *part_v = nullptr;
  cplx *Linv = nullptr, *v1 = nullptr, *v2 = nullptr;
  void alloc(int N, int sm_count);
  void release();
};
int lu_solve(cplx *A, int N, size_t lda, LuWork &w, const cplx *b, cplx *x, int sm_count, cudaStream_t st,
             long &launches);

// ---- ob_vec.cu (Krylov vector kernels) ----
// h[t] = v_t^H w for t < j (V is ldv-strided), deterministic two-stage reduction
void launch_multi_dot(const cplx *V, size_t ldv, int j, const cplx *w, int N, cplx *h_dev, cplx *scratch,
                      cudaStream_t st);
// w -= sum_t h[t] v_t
void launch_multi_axpy(const cplx *V, size_t ldv, int j, const cplx *h_dev, cplx *w, int N, cudaStream_t st);
void launch_norm2(const cplx *w, int N, double *out_dev, double *scratch, cudaStream_t st);
void launch_scale_to(const cplx *w, double inv, cplx *v, int N, cudaStream_t st); // v = w * inv
void launch_axpby(cplx a, const cplx *x, cplx b, const cplx *y, cplx *z, int N, cudaStream_t st); // z = a x + b y
void launch_combine(const cplx *V, size_t ldv, int j, const cplx *coef_dev, cplx *x, int N, cudaStream_t st); // x += V c
void launch_hadamard(const cplx *a, const cplx *b, const cplx *c, cplx *out, int N, int conj_out, cudaStream_t st);
size_t vec_scratch_elems(int N, int jmax);
// fused Arnoldi step (one cooperative launch): mode 0 = classical GS + DGKS, 1 = modified GS; h_out[0..j] = h,
// h_out[j+1] = ||w||^2 before, h_out[j+2] = ||w||^2 after; vnext = w / ||w||; optional XP/XS staging (ob_pairs.cu)
double measure_fp64_peak(int sm_count, cudaStream_t st); // TFLOP/s, DFMA micro-benchmark
bool arnoldi_fused_supported(int N, int sm_count);
size_t arnoldi_scratch_elems(int N, int jmax, int sm_count);
void launch_arnoldi_step(const cplx *V, size_t ldv, int j, cplx *w, int N, int mode, cplx *h_out, cplx *partial,
                         unsigned *sync, cplx *vnext, cplx *XP, cplx *XS, int n_harm, int sm_count, cudaStream_t st);

} // namespace ob
