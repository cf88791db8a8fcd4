/* optimet_b200.h -- C ABI of the B200-native multiple-scattering hot path.
 *
 * Drop-in boundary for OPTIMET-3D's solver plugin: everything
 * optimet::solver::AbstractSolver::{update, solve} (srcAna/Solver.h:62-144) and the
 * Result cross-section reductions (srcAna/Result.cpp:557-794) compute is reachable through
 * these entry points with plain pointers and sizes.  INTEGRATION.md shows the C++11 adaptor
 * (`solver::B200Matrix : AbstractSolver`) a maintainer adds to srcAna/Solver.cpp:30-54.
 *
 * Conventions
 *   - complex numbers are interleaved (re, im) doubles == std::complex<double> == Eigen t_complex;
 *   - all pointers are HOST pointers; every call returns 0 on success, non-zero on error
 *     (message via ob_last_error), mirroring the reference's std::runtime_error sites;
 *   - harmonic: 1 = fundamental (FF), 2 = second harmonic (SH);
 *   - n = nMax (nMax + 2) harmonics per polarisation, particle block = 2n, flat index
 *     p = l(l+1) - m - 1 (srcAna/CompoundIterator.h:24-31), vector layout per particle
 *     [TE(n) ; TM(n)], N = 2 n N_obj;
 *   - one context per process and GPU; with world > 1 the context owns the particle block-rows
 *     ob_partition() assigns to its rank and Krylov vectors are replicated on every rank.
 */
#ifndef OPTIMET_B200_H
#define OPTIMET_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ob_ctx ob_ctx;

/* GMRES flavours.  ZCOMP = the in-tree Gmres_Zcomp (srcAna/PreconditionedMatrix.cpp:892-985:
 * x0 = 0, modified Gram-Schmidt, |g|/||b|| stopping rule, max_iters per cycle, max_restarts cycles).
 * BELOS = Belos "GMRES" as driven by srcAna/scalapack/LinearSystemSolver.hpp:94-142 (x0 = b, DGKS,
 * ||r||/||r0|| rule, restart = "Num Blocks", max_iters = "Maximum Iterations" total).
 * DIRECT = device dense solve (blocked LU, partial pivoting): the counterpart of the serial
 * S.colPivHouseholderQr().solve(Q) (srcAna/PreconditionedMatrixSolver.h:58,75, taken when ACA is off) and of pzgesv_
 * (srcAna/ScalapackSolver.cpp:54-128; Solver = "eigen" | "scalapack", srcAna/Solver.cpp:40-43).  One GPU, needs
 * 16 N^2 bytes; tol / max_iters / restart are ignored, iters = 0 is returned. */
enum { OB_GMRES_ZCOMP = 1, OB_GMRES_BELOS = 2, OB_SOLVE_DIRECT = 3 };

typedef struct ob_gmres_opts {
  int flavour;      /* OB_GMRES_ZCOMP | OB_GMRES_BELOS | OB_SOLVE_DIRECT */
  double tol;       /* "Convergence Tolerance" / Gmres_Zcomp tol */
  int max_iters;    /* ZCOMP: maxit per cycle; BELOS: "Maximum Iterations" */
  int restart;      /* BELOS: "Num Blocks" (ignored by ZCOMP) */
  int max_restarts; /* ZCOMP: no_rest; BELOS: "Maximum Restarts" */
} ob_gmres_opts;

/* ---- context ---- */
int ob_create(int device, ob_ctx **out);
void ob_destroy(ob_ctx *ctx);
const char *ob_last_error(ob_ctx *ctx); /* ctx may be NULL: last error of a failed ob_create */
int ob_device_info(ob_ctx *ctx, int *sm_count, size_t *free_bytes, size_t *total_bytes);

/* ---- multi-GPU row sharding (replaces the BLACS/MPI layer, srcAna/scalapack, srcAna/mpi) ----
 * ob_partition: contiguous particle block-rows with the remainder rule of
 * srcAna/PreconditionedMatrix.cpp:418-424 (pure host arithmetic, no GPU needed). */
int ob_partition(int nobj, int world, int rank, int *first, int *count);
int ob_comm_unique_id(char out[128]);                                  /* rank 0, then broadcast by the host */
int ob_comm_init(ob_ctx *ctx, const char uid[128], int rank, int world); /* NCCL communicator over NVLink */
/* plan and assemble as `rank` of `world` WITHOUT a communicator: ob_matvec_partial then returns this shard's partial sums
 * acc_p (pair and rotated-axial forms: the local pairs applied to x, before the cross-rank sum);
 * y = x - T .* sum_ranks acc.  Lets one device execute every shard in turn (tests/test_gpu_shards.py checks the
 * sharding on a one-GPU box); the sharded solvers themselves need ob_comm_init or ob_create_multi. */
int ob_set_shard(ob_ctx *ctx, int rank, int world);
int ob_matvec_partial(ob_ctx *ctx, int harmonic, const double *x, double *acc);

/* page-lock / release caller memory that ob_run copies results into (cudaHostRegister / cudaHostUnregister): the host
 * layer does this once for its persistent coefficient vectors so that the device -> host copies run at PCIe speed */
int ob_host_register(ob_ctx *ctx, void *ptr, size_t bytes);
int ob_host_unregister(ob_ctx *ctx, void *ptr);

/* ---- one process, several GPUs.  The reference's solver::factory returns the serial solver unconditionally in
 * non-MPI builds (srcAna/Solver.cpp:30-34): a serial Optimet3D gets all the GPUs of the box through this group.  One
 * context per device, NCCL communicators created inside the process, one host worker thread per GPU for every call;
 * rank r owns its shard exactly as with one process per GPU.  Inputs are replicated to every context, outputs are
 * rank 0's (the replicated vectors are bit-identical on every rank), cross sections are summed over the ranks. */
typedef struct ob_multi ob_multi;
int ob_create_multi(int ngpu, const int *devices, ob_multi **out);
void ob_destroy_multi(ob_multi *m);
int ob_multi_size(const ob_multi *m);
ob_ctx *ob_multi_ctx(ob_multi *m, int rank);
const char *ob_multi_last_error(ob_multi *m);
int ob_multi_set_cluster(ob_multi *m, int nobj, const double *xyz_m, const double *radius_m, int nMax, int nMaxS);
int ob_multi_set_frequency(ob_multi *m, double omega, const double waveK[2], const double eps_b[2], const double mu_b[2],
                           const double *eps, const double *mu, const double *eps_SH, const double *mu_SH,
                           const double *ksippp, const double *ksiparppar, const double *gamma);
int ob_multi_set_incident(ob_multi *m, const double *a, const double *b);
int ob_multi_set_option(ob_multi *m, const char *name, double value);
int ob_multi_run(ob_multi *m, const ob_gmres_opts *opts, int do_sh, double *X_sca, double *X_int, double *X_sca_SH,
                 double *X_int_SH, double cs[5], int stats[2]);

/* ---- problem definition (what solver->update(run) reads from Geometry / Excitation) ---- */
/* positions: Cartesian metres (Tools::toCartesian of Scatterer::vR), radius in metres */
int ob_set_cluster(ob_ctx *ctx, int nobj, const double *xyz_m, const double *radius_m, int nMax, int nMaxS);
/* omega = c k0; waveK = k0 sqrt(eps_b,r mu_b,r) (Excitation::waveK); eps/mu absolute (ElectroMagnetic::epsilon, mu);
 * per-particle arrays hold N_obj complex values each; ksippp/ksiparppar/gamma = SH tensor coefficients */
int ob_set_frequency(ob_ctx *ctx, double omega, const double waveK[2], const double eps_b[2], const double mu_b[2],
                     const double *eps, const double *mu, const double *eps_SH, const double *mu_SH,
                     const double *ksippp, const double *ksiparppar, const double *gamma);
/* plane-wave coefficients at the origin, n complex each (Excitation::dataIncAp / dataIncBp) */
int ob_set_incident(ob_ctx *ctx, const double *a_origin, const double *b_origin);

/* ---- unit-level surface (parity with the reference's free functions / classes) ---- */
/* Coupling(relR, k, nMax, regular_flag): A = diagonal, B = offdiagonal, n x n column-major
 * (srcAna/Coupling.h:29-41; regular_flag has the ctor's meaning: true = particle coupling/Hankel) */
int ob_vtac(ob_ctx *ctx, const double relR_sph[3], const double k[2], int regular_flag, int nMax, double *A,
            double *B);
/* which: 0 T_FF, 1 T_SH, 2 T_SH1_outer, 3 T_SH2_outer, 4 Iaux, 5 IauxSH1, 6 IauxSH2 (srcAna/Scatterer.cpp:39-412);
 * out: N_obj x 2n complex */
int ob_particle_factors(ob_ctx *ctx, int which, double *out);
/* local incident coefficients of every particle, N complex (Excitation::getIncLocal) */
int ob_inc_local(ob_ctx *ctx, double *out);

/* ---- matrix (preconditioned_scattering_matrix[SH], srcAna/PreconditionedMatrix.cpp:350-400, 555-610) ---- */
int ob_assemble(ob_ctx *ctx, int harmonic);
int ob_release_matrix(ob_ctx *ctx, int harmonic);
int ob_fetch_block(ob_ctx *ctx, int harmonic, int i, int j, double *out); /* 2n x 2n column-major; i must be local */
int ob_fetch_matrix(ob_ctx *ctx, int harmonic, double *out);              /* local slab, (2n count) x N column-major */
int ob_matvec(ob_ctx *ctx, int harmonic, const double *x, double *y);     /* y = S x, full length N on every rank */

/* ---- ACA-compressed operator (<ACA compression="yes">; ob_set_option("operator", 2) before ob_assemble) ----
 * Scattering_matrix_ACA_FF / _SH (srcAna/PreconditionedMatrix.cpp:489-551, 699-759): blocks with
 * distance >= 2 (r_i + r_j) are stored as U (2n x r) V (r x 2n) from ACA_compression (:760-859, eps 1e-3, pivots by
 * getMaxInd :861-889), near blocks dense, the diagonal as the identity; ob_matvec / ob_solve / ob_run then apply
 * matvec (:1058-1085).  ob_fetch_block / ob_fetch_matrix are not available in this form.
 * ob_aca_block: rank > 0: U = 2n x rank column-major, V = rank rows of 2n entries, I / J = pivot rows / columns in the
 * order taken (may be NULL); rank = -1: dense near block in U (2n x 2n column-major); rank = 0: identity diagonal.
 * Buffers must hold 2n x 2n complex (U, V) and 2n ints (I, J).  i must be local to this rank. */
int ob_aca_block(ob_ctx *ctx, int harmonic, int i, int j, int *rank, double *U, double *V, int *I, int *J);
/* unit surface of ACA_compression(U, V, CoupMat): caller-supplied dim x dim column-major block, same outputs */
int ob_aca_compress(ob_ctx *ctx, int dim, const double *C, int *rank, double *U, double *V, int *I, int *J);
/* out: stored bytes, dense-equivalent bytes of the local slab, low-rank blocks, dense near blocks, mean rank, max rank */
int ob_aca_stats(ob_ctx *ctx, int harmonic, double out[6]);

/* ---- sources (source_vector, source_vectorSH, source_vectorSH_K1ana; PreconditionedMatrix.cpp:1327-1436) ---- */
int ob_source_ff(ob_ctx *ctx, double *Q);
int ob_set_cg_tables(ob_ctx *ctx, const double *const tables[9]); /* order of Simulation.cpp:616 */
int ob_build_cg_tables(ob_ctx *ctx);                              /* device build (symbol::*coeff, Symbol.cpp:1036-1446) */
int ob_fetch_cg_table(ob_ctx *ctx, int t, double *out);
int ob_source_sh(ob_ctx *ctx, const double *Xint_conj, double *K, double *K1ana);

/* ---- solve (Solver::solve pieces) ---- */
/* rhs == NULL: use the resident source of that harmonic (Q or K) */
int ob_solve(ob_ctx *ctx, int harmonic, const double *rhs, double *x, const ob_gmres_opts *opts, int *iters,
             double *relres);
/* unit-level surface of OB_SOLVE_DIRECT: x = A^-1 b for a caller-supplied N x N column-major complex matrix (the
 * arithmetic behind Eigen's colPivHouseholderQr().solve, PreconditionedMatrixSolver.h:58, and pzgesv_,
 * scalapack/LinearSystemSolver.hpp:84-91).  A is not modified; fails with "singular" on an exactly zero pivot. */
int ob_dense_solve(ob_ctx *ctx, int N, const double *A, const double *b, double *x);
int ob_unprecondition_ff(ob_ctx *ctx, const double *X_sca, double *X_int);                          /* Solver.cpp:57-77 */
int ob_unprecondition_sh(ob_ctx *ctx, const double *X_sca_SH, const double *K1ana, double *X_int_SH); /* Solver.cpp:95-116 */

/* ---- whole step, device resident: update() + solve() + Result cross sections for the current
 * frequency (Simulation.cpp:648-667).  Any of the output vector pointers may be NULL.
 * cs: ext_FF, sca_FF, abs_FF (= ext - sca, Simulation.cpp:659), sca_SH, abs_SH.  stats: iters_FF, iters_SH.
 * With world > 1 (ob_comm_init / ob_set_shard) cs holds THIS RANK's partial sums over its own particles, as the
 * reference's ranks hold theirs before MPI_Gather (Simulation.cpp:510-572): the caller adds them over the ranks
 * (ob_multi_run and optimet_b200/sharding.py do).  The coefficient vectors are complete on every rank. */
int ob_run(ob_ctx *ctx, const ob_gmres_opts *opts, int do_sh, double *X_sca, double *X_int, double *X_sca_SH,
           double *X_int_SH, double cs[5], int stats[2]);

/* ---- reductions (Result::get*CrossSection*, srcAna/Result.cpp:557-794); per-rank partial sums when world > 1 ---- */
int ob_cross_sections(ob_ctx *ctx, const double *X_sca, const double *X_int, const double *X_sca_SH,
                      const double *X_int_SH, int do_sh, double cs[5]);

/* ---- near-field maps (Result::setFields / getEHFields with projection = false, srcAna/Result.cpp:74-300, 896-934;
 * AuxCoefficients M, N, X-1, X+1, srcAna/AuxCoefficients.cpp:108-343; Geometry::checkInner / COEFFpartSH,
 * srcAna/Geometry.cpp:147-163, 458-495; symbol::CXm1 / CXp1, srcAna/Symbol.cpp:482-635) ----
 * pts_sph: npts x (r, theta, phi) in metres / radians as OutputGrid::getPoint returns them (OutputGrid.cpp:132-157).
 * X_*: the four solution vectors of solve(); NULL = the device-resident ones of the last ob_run.
 * out: npts x 4 x 3 complex = E_FF, H_FF, E_SH, H_SH with Cartesian components x, y, z (the reference stores them in
 * the rrr / the / phi slots of SphericalP); E_FF, H_FF include the incident field outside the spheres; E_SH includes
 * the particular solution inside.  inner (may be NULL): index of the sphere containing the point, -1 outside. */
int ob_fields(ob_ctx *ctx, long npts, const double *pts_sph, const double *X_sca, const double *X_int,
              const double *X_sca_SH, const double *X_int_SH, int do_sh, double *out, int *inner);

/* ---- instrumentation ---- */
/* device-event timings (ms) of the last ob_run: 0 factors+source, 1 assemble FF, 2 solve FF, 3 SH source,
 * 4 assemble SH, 5 solve SH, 6 cross sections, 7 matvec total (streaming kernel only), 8 matvec count,
 * 9 kernel launches, 10 algorithmic bytes of one operator apply in the form streamed, 11-13 iteration trace */
int ob_timings(ob_ctx *ctx, double out[16]);
/* CUDA-event stopwatch on the library's stream: op 0 = start, op 1 = stop (+ synchronise) -> elapsed ms */
int ob_timer(ob_ctx *ctx, int op, double *ms);
/* measured FP64 FMA peak of the device in TFLOP/s (DFMA micro-benchmark): the denominator north_star asks for when
 * the assembly is reported "as achieved FP64 FLOP/s against B200 FP64 peak" */
int ob_measure_fp64_peak(ob_ctx *ctx, double *tflops);
/* options: "operator" (0 dense slab | 1 pair form, default | 2 ACA-compressed | 3 rotated-axial form: exact, 12-15x
 * fewer bytes than the pair form, see csrc/ob_rot.cu), "eps_aca" (1e-3), "aca_budget_mb", "assemble_minb", "rot_assembly" (1, default: axial-only recursion | 0: cross-check path through the full translation block), "rot_share" (1, default: the harmonic assembled second reads phases / small-d matrices from the first one's records), "keep_matrices", "fused_arnoldi", "matvec_variant",
 * "pairs_kb", "pairs_groups" (tuning), "trace_iterations", "reset_timings" */
int ob_set_option(ob_ctx *ctx, const char *name, double value);

#ifdef __cplusplus
}
#endif
#endif
