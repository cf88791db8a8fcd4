// ob_api.cu -- context, GMRES drivers, row-sharded collectives and the C ABI (include/optimet_b200.h).
#include "../../include/optimet_b200.h"
#include "ob_internal.h"
#include <algorithm>
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>

namespace ob {

typedef std::complex<double> hcd;

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time (the torch-bundled libnccl.so.2 when the host process imported torch)
// ---------------------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if(handle)
      return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for(const char *nm : names) {
      handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if(handle)
        break;
    }
    if(!handle)
      return false;
#define OB_SYM(field, name) *(void **)(&field) = dlsym(handle, name)
    OB_SYM(GetUniqueId, "ncclGetUniqueId");
    OB_SYM(CommInitRank, "ncclCommInitRank");
    OB_SYM(CommDestroy, "ncclCommDestroy");
    OB_SYM(AllGather, "ncclAllGather");
    OB_SYM(AllReduce, "ncclAllReduce");
    OB_SYM(Broadcast, "ncclBroadcast");
    OB_SYM(GroupStart, "ncclGroupStart");
    OB_SYM(GroupEnd, "ncclGroupEnd");
    OB_SYM(GetErrorString, "ncclGetErrorString");
#undef OB_SYM
    return GetUniqueId && CommInitRank && AllGather && AllReduce && Broadcast && GroupStart && GroupEnd;
  }
};
static NcclApi g_nccl;
#define OB_NCCL(call)                                                                                                  \
  do {                                                                                                                 \
    ncclResult_t r__ = (call);                                                                                         \
    if(r__ != ncclSuccess)                                                                                             \
      throw ob::Error(std::string("NCCL error: ") +                                                                    \
                      (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "unknown"));                               \
  } while(0)

template <class T> struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    if(count <= n && p)
      return;
    release();
    if(count == 0)
      return;
    OB_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void release() {
    if(p)
      cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

static void partition(int nobj, int world, int rank, int &first, int &count) {
  // remainder rule of srcAna/PreconditionedMatrix.cpp:418-424
  if(rank < nobj % world) {
    first = rank * (nobj / world + 1);
    count = nobj / world + 1;
  } else {
    first = rank * (nobj / world) + nobj % world;
    count = nobj / world;
  }
}

struct HarmonicState {
  int nMax = 0, n = 0;
  cplx k = mk(0, 0);
  DevBuf<cplx> S; // dense form: local slab, column-major, ld = M_loc
  bool assembled = false;
  MatvecPlan plan;
  // pair form (ob_pairs.cu): unscaled A^T, B^T of the local pairs i < j
  int mode = 0; // operator form this harmonic was assembled in (0 dense, 1 pairs, 2 ACA-compressed)
  AcaOperator aca;
  // rotated-axial form (ob_rot.cu): one record per local pair i < j
  DevBuf<unsigned char> rot;
  // records the apply takes the geometry sections (phases, small-d matrices) from: the harmonic's own, or those of the
  // other harmonic when that one was assembled first for the same pairs (the sections do not depend on k)
  const unsigned char *rot_geo = nullptr;
  RotLayout rl;
  RotPlan rplan;
  int rplan_world = -1, rplan_rank = -1;
  DevBuf<cplx> AB;
  PairPlan pplan;
  int pplan_world = -1, pplan_rank = -1;
};

} // namespace ob

using namespace ob;

struct ob_ctx {
  int device = 0, sm_count = 148;
  cudaStream_t st = nullptr;
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  std::string err;

  // cluster
  int nobj = 0, nMax = 0, nMaxS = 0, first = 0, count = 0; // local particle rows [first, first+count)
  DevBuf<double> xyz, radius;
  std::vector<double> h_xyz, h_radius; // host copies: ACA admissibility test (PreconditionedMatrix.cpp:526)
  double eps_aca = 1e-3;               // PreconditionedMatrix.cpp:772
  size_t aca_budget = (size_t)4 << 30; // scratch bytes of one ACA assembly batch
  AcaScratch aca_scratch;
  DevBuf<cplx> spare_AB; // parked pair storage (keep_matrices = 0)
  // frequency / materials
  bool have_freq = false, have_inc = false;
  double omega = 0;
  cplx waveK = mk(0, 0), eps_b = mk(0, 0), mu_b = mk(0, 0);
  DevBuf<cplx> mat[7]; // eps, mu, eps_SH, mu_SH, ksippp, ksiparppar, gamma
  std::vector<hcd> h_epsr_SH; // eps_SH / eps0 per particle (sigma in Result.cpp:784)
  DevBuf<cplx> fac[7]; // Mie factors
  bool fac_valid = false;
  DevBuf<cplx> ainc;   // [a ; b] 2n
  VtacTableSet tabs[2];
  HarmonicState hs[2];
  // CG tables
  DevBuf<double> cg[9];
  int cg_nmax = -1;
  // resident vectors (full length N, replicated)
  DevBuf<cplx> Q, Ksrc, K1ana, Xsca, Xint, XscaSH, XintSH, tmpA, tmpB;
  // GMRES workspace
  DevBuf<cplx> V, w, h_dev, dot_scratch, ycoef, arn_partial;
  // direct solve (ob_lu.cu): dense N x N work matrix, released after each solve
  DevBuf<cplx> lu_mat;
  LuWork lu;
  DevBuf<unsigned> arn_sync;
  bool fused_arnoldi = true;
  DevBuf<double> red_d;
  DevBuf<cplx> red_c;
  // instrumentation
  double tim[16] = {0};
  long launches = 0;
  int matvec_variant = 0;
  int operator_mode = 1; // 0 = dense slab (reference layout), 1 = compact pair form (default), 2 = ACA-compressed
  bool keep_matrices = true;
  bool rot_share = true; // rotated-axial form: the second harmonic reads the geometry sections of the first one's records
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm0 = nullptr, evm1 = nullptr, evt0 = nullptr, evt1 = nullptr;
  // matvec timing without a host sync per apply: a pool of event pairs, read back lazily (flush_matvec_timing)
  std::vector<cudaEvent_t> mv_ev; // 4 per apply: kernel start, kernel end, [trace: after reduce, after Arnoldi step]
  int mv_ev_used = 0;
  bool trace = false;             // "trace_iterations": GPU-timeline breakdown of a GMRES iteration into tim[11..13]
  std::vector<char> mv_ev_arn;    // whether slot 3 of an entry was recorded
  hcd *h_pinned = nullptr; // pinned staging for the per-iteration Hessenberg column read-back
  size_t h_pinned_cap = 0;
  cudaEvent_t evh = nullptr; // read-back of the Hessenberg column complete
  bool speculate = true;     // enqueue the next operator apply ahead of the host's convergence test (gmres_belos)

  int N(int h) const { return 2 * hs[h - 1].n * nobj; }
  int Mloc(int h) const { return 2 * hs[h - 1].n * count; }
};

namespace ob {

static std::string g_create_error;

static void check_harmonic(int harmonic) {
  if(harmonic != 1 && harmonic != 2)
    throw Error("harmonic must be 1 (FF) or 2 (SH)");
}
static void need(bool cond, const char *msg) {
  if(!cond)
    throw Error(msg);
}

static VtacTableSet &tables_for(ob_ctx *c, int nMax) {
  for(int i = 0; i < 2; ++i)
    if(c->tabs[i].nMax == nMax)
      return c->tabs[i];
  for(int i = 0; i < 2; ++i)
    if(c->tabs[i].nMax < 0) {
      c->tabs[i].build(nMax);
      return c->tabs[i];
    }
  c->tabs[1].build(nMax);
  return c->tabs[1];
}

static void ensure_factors(ob_ctx *c) {
  if(c->fac_valid)
    return;
  need(c->have_freq, "ob_set_frequency has not been called");
  MieInputs in;
  in.nobj = c->nobj;
  in.nMax = c->nMax;
  in.nMaxS = c->nMaxS;
  in.omega = c->omega;
  in.eps_b = c->eps_b;
  in.mu_b = c->mu_b;
  in.radius = c->radius.p;
  in.eps = c->mat[0].p;
  in.mu = c->mat[1].p;
  in.eps_SH = c->mat[2].p;
  in.mu_SH = c->mat[3].p;
  cplx *out[7];
  for(int f = 0; f < 7; ++f) {
    int nm = (f == 0 || f == 4) ? c->nMax : c->nMaxS;
    c->fac[f].alloc((size_t)c->nobj * 2 * flat_max(nm));
    out[f] = c->fac[f].p;
  }
  launch_mie(in, out, c->st);
  c->launches += 1;
  c->fac_valid = true;
}

// all ranks hold `vec` (full length); each rank produced its own slice [first*blk, (first+count)*blk)
static void allgather_slices(ob_ctx *c, cplx *vec, int blk) {
  if(c->world == 1)
    return;
  need(c->comm != nullptr, "world > 1 but ob_comm_init has not been called");
  if(c->nobj % c->world == 0) {
    size_t cnt = (size_t)c->count * blk * 2; // doubles
    OB_NCCL(g_nccl.AllGather((const void *)(vec + (size_t)c->first * blk), (void *)vec, cnt, ncclDouble, c->comm,
                             c->st));
  } else {
    OB_NCCL(g_nccl.GroupStart());
    for(int r = 0; r < c->world; ++r) {
      int f, n;
      partition(c->nobj, c->world, r, f, n);
      cplx *ptr = vec + (size_t)f * blk;
      OB_NCCL(g_nccl.Broadcast((const void *)ptr, (void *)ptr, (size_t)n * blk * 2, ncclDouble, r, c->comm, c->st));
    }
    OB_NCCL(g_nccl.GroupEnd());
  }
}

// The rotated-axial records of harmonic h (0-based) are about to be freed or moved: a harmonic that reads its geometry
// sections from them must be assembled again.
static void rot_records_drop(ob_ctx *c, int h) {
  HarmonicState &H = c->hs[h], &O = c->hs[1 - h];
  if(H.rot.p && O.rot_geo == H.rot.p) {
    O.assembled = false;
    O.rot_geo = nullptr;
  }
  H.rot.release();
  H.rot_geo = nullptr;
}

static void assemble(ob_ctx *c, int harmonic) {
  check_harmonic(harmonic);
  need(c->nobj > 0, "ob_set_cluster has not been called");
  ensure_factors(c);
  HarmonicState &H = c->hs[harmonic - 1];
  const size_t M = (size_t)c->Mloc(harmonic), N = (size_t)c->N(harmonic);
  H.k = harmonic == 1 ? c->waveK : cscale(c->waveK, 2.0);
  VtacTableSet &ts = tables_for(c, H.nMax);
  if(c->operator_mode == 1) {
    if(H.pplan.nobj != c->nobj || H.pplan.n != H.n || H.pplan_world != c->world || H.pplan_rank != c->rank) {
      pair_plan_build(H.pplan, c->nobj, H.n, c->world, c->rank, c->sm_count);
      H.pplan_world = c->world;
      H.pplan_rank = c->rank;
    }
    H.S.release();
    H.aca.release();
    rot_records_drop(c, harmonic - 1);
    c->aca_scratch.release();
    if(!H.AB.p && c->spare_AB.p && c->spare_AB.n >= std::max<size_t>(1, pair_storage_elems(H.pplan))) {
      std::swap(H.AB.p, c->spare_AB.p);
      std::swap(H.AB.n, c->spare_AB.n);
    }
    H.AB.alloc(std::max<size_t>(1, pair_storage_elems(H.pplan)));
    launch_assemble_pairs(ts, c->xyz.p, H.k, H.pplan.pair_ij, H.pplan.npairs, H.AB.p, c->st);
    c->launches += 1;
    H.mode = 1;
    H.assembled = true;
    return;
  }
  H.AB.release();
  c->spare_AB.release();
  if(c->operator_mode == 3) { // rotated-axial form: phases, axial A/B and Wigner small-d per pair (ob_rot.cu)
    H.S.release();
    H.aca.release();
    if(H.rplan.nobj != c->nobj || H.rplan.n != H.n || H.rplan_world != c->world || H.rplan_rank != c->rank) {
      rot_plan_build(H.rplan, c->nobj, H.nMax, c->world, c->rank, c->sm_count);
      H.rplan_world = c->world;
      H.rplan_rank = c->rank;
    }
    H.rl = rot_layout(H.nMax);
    const size_t rot_bytes = std::max<size_t>(16, (size_t)H.rplan.npairs * H.rl.rec_bytes);
    if(rot_bytes > H.rot.n)
      rot_records_drop(c, harmonic - 1);
    H.rot.alloc(rot_bytes);
    // phases and small-d matrices depend on the geometry only: when the other harmonic holds them for the same pairs
    // (assembled for the current cluster, both resident), this one stores its axial coefficients alone and k_rot_tables
    // is skipped; the apply fetches the geometry sections from the other harmonic's records
    HarmonicState &O = c->hs[2 - harmonic];
    const bool share = c->rot_share && c->keep_matrices && O.assembled && O.mode == 3 && O.rot.p && O.rot_geo == O.rot.p &&
                       O.nMax == H.nMax && O.rplan.nobj == H.rplan.nobj && O.rplan.I == H.rplan.I &&
                       O.rplan.npairs == H.rplan.npairs && O.rplan.strip0 == H.rplan.strip0 &&
                       O.rplan.nstrips == H.rplan.nstrips && O.rplan_world == H.rplan_world && O.rplan_rank == H.rplan_rank;
    launch_assemble_rot(ts, c->xyz.p, H.k, H.rplan.pair_ij, H.rplan.npairs, H.rot.p, H.rl, c->sm_count, c->st, !share);
    H.rot_geo = share ? O.rot.p : H.rot.p;
    c->launches += share ? 1 : 2;
    H.mode = 3;
    H.assembled = true;
    return;
  }
  rot_records_drop(c, harmonic - 1);
  if(c->operator_mode == 2) { // Scattering_matrix_ACA_FF / _SH (PreconditionedMatrix.cpp:489-551, 699-759)
    H.S.release();
    aca_build(H.aca, c->aca_scratch, ts, c->xyz.p, c->fac[harmonic == 1 ? 0 : 1].p, H.k, c->nobj, c->first, c->count, c->h_xyz.data(),
              c->h_radius.data(), c->eps_aca, c->aca_budget, c->sm_count, c->st, c->launches);
    H.mode = 2;
    H.assembled = true;
    return;
  }
  H.aca.release();
  c->aca_scratch.release();
  H.mode = 0;
  H.S.alloc(M * N);
  launch_assemble(ts, c->xyz.p, c->fac[harmonic == 1 ? 0 : 1].p, H.k, c->nobj, c->first, c->count, H.S.p, M, c->st);
  c->launches += 1;
  if(H.plan.M != (int)M || H.plan.N != (int)N || H.plan.variant != c->matvec_variant)
    matvec_plan(H.plan, (int)M, (int)N, M, c->sm_count, c->matvec_variant);
  H.assembled = true;
}

static void flush_matvec_timing(ob_ctx *c) {
  for(int i = 0; i < c->mv_ev_used; i += 4) {
    cudaEventSynchronize(c->mv_ev[i + 1]);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->mv_ev[i], c->mv_ev[i + 1]);
    c->tim[7] += ms;
    c->tim[8] += 1;
    if(c->trace) {
      cudaEventSynchronize(c->mv_ev[i + 2]);
      cudaEventElapsedTime(&ms, c->mv_ev[i + 1], c->mv_ev[i + 2]);
      c->tim[11] += ms; // streaming kernel end -> operator result complete (reduce / all-reduce / finalize)
      if(c->mv_ev_arn[i / 4]) {
        cudaEventSynchronize(c->mv_ev[i + 3]);
        cudaEventElapsedTime(&ms, c->mv_ev[i + 2], c->mv_ev[i + 3]);
        c->tim[12] += ms; // -> Arnoldi step done
        if(i + 4 < c->mv_ev_used) {
          cudaEventElapsedTime(&ms, c->mv_ev[i + 3], c->mv_ev[i + 4]);
          c->tim[13] += ms; // -> next streaming kernel starts (read-back, host sync, Givens, launch latency)
        }
      }
    }
  }
  c->mv_ev_used = 0;
}

// y (full length, replicated) = S x
static void matvec(ob_ctx *c, int harmonic, const cplx *x, cplx *y, bool x_staged = false) {
  HarmonicState &H = c->hs[harmonic - 1];
  need(H.assembled, "matrix not assembled (call ob_assemble)");
  if(c->mv_ev_used + 4 > (int)c->mv_ev.size())
    flush_matvec_timing(c);
  cudaEvent_t evm0 = c->mv_ev[c->mv_ev_used], evm1 = c->mv_ev[c->mv_ev_used + 1];
  cudaEvent_t ev_done = c->mv_ev[c->mv_ev_used + 2];
  c->mv_ev_arn[c->mv_ev_used / 4] = 0;
  c->mv_ev_used += 4;
  if(H.mode == 1) {
    const cplx *T = c->fac[harmonic == 1 ? 0 : 1].p;
    const size_t N = (size_t)c->N(harmonic);
    if(c->world == 1) {
      launch_matvec_pairs(H.pplan, H.AB.p, x, T, y, 1, c->st, evm0, evm1, x_staged);
      c->launches += x_staged ? 2 : 3;
    } else {
      need(c->comm != nullptr, "world > 1 but ob_comm_init has not been called");
      launch_matvec_pairs(H.pplan, H.AB.p, x, T, H.pplan.acc, 0, c->st, evm0, evm1, x_staged);
      OB_NCCL(g_nccl.AllReduce((const void *)H.pplan.acc, (void *)H.pplan.acc, 2 * N, ncclDouble, ncclSum, c->comm,
                               c->st));
      launch_pairs_finalize(x, T, H.pplan.acc, N, y, c->st);
      c->launches += 4;
    }
    c->tim[10] = 16.0 * (double)pair_storage_elems(H.pplan) + 32.0 * (double)N;
  } else if(H.mode == 3) {
    const cplx *T = c->fac[harmonic == 1 ? 0 : 1].p;
    const size_t N = (size_t)c->N(harmonic);
    if(c->world == 1) {
      launch_matvec_rot(H.rplan, H.rl, H.rot.p, H.rot_geo, x, T, y, 1, c->st, evm0, evm1);
      c->launches += 2;
    } else {
      need(c->comm != nullptr, "world > 1 but ob_comm_init has not been called");
      launch_matvec_rot(H.rplan, H.rl, H.rot.p, H.rot_geo, x, T, H.rplan.acc, 0, c->st, evm0, evm1);
      OB_NCCL(g_nccl.AllReduce((const void *)H.rplan.acc, (void *)H.rplan.acc, 2 * N, ncclDouble, ncclSum, c->comm,
                               c->st));
      launch_pairs_finalize(x, T, H.rplan.acc, N, y, c->st);
      c->launches += 4;
    }
    c->tim[10] = (double)H.rplan.npairs * (double)H.rl.rec_bytes + 32.0 * (double)N;
  } else if(H.mode == 2) { // matvec of PreconditionedMatrix.cpp:1058-1085 on the compressed blocks
    launch_matvec_aca(H.aca, x, y + (size_t)c->first * 2 * H.n, c->st, evm0, evm1);
    c->launches += H.aca.nch > 1 ? 2 : 1;
    allgather_slices(c, y, 2 * H.n);
    c->tim[10] = 16.0 * H.aca.stored_elems + 32.0 * (double)c->N(harmonic);
  } else {
    launch_matvec(H.plan, H.S.p, x, y + (size_t)c->first * 2 * H.n, c->st, evm0, evm1);
    c->launches += matvec_launches_per_apply(H.plan);
    allgather_slices(c, y, 2 * H.n);
    c->tim[10] = 16.0 * (double)c->Mloc(harmonic) * (double)c->N(harmonic) + 32.0 * (double)c->N(harmonic);
  }
  if(c->trace)
    cudaEventRecord(ev_done, c->st);
}

static void ensure_gmres_ws(ob_ctx *c, int N, int basis) {
  c->V.alloc((size_t)(basis + 1) * N);
  c->w.alloc(N);
  c->h_dev.alloc(basis + 4);
  c->ycoef.alloc(basis + 4);
  c->dot_scratch.alloc(vec_scratch_elems(N, basis + 2));
  c->arn_partial.alloc(arnoldi_scratch_elems(N, basis + 2, c->sm_count));
  if(!c->arn_sync.p) {
    c->arn_sync.alloc(4);
    OB_CUDA(cudaMemsetAsync(c->arn_sync.p, 0, 4 * sizeof(unsigned), c->st));
  }
}

// one fused Arnoldi step on w against V[0..j]; returns through h_dev (see launch_arnoldi_step)
static bool use_fused(ob_ctx *c, int N, int basis) {
  return c->fused_arnoldi && basis + 1 <= 256 && arnoldi_fused_supported(N, c->sm_count);
}
static void fused_step(ob_ctx *c, int harmonic, int j, int mode) {
  const int N = c->N(harmonic);
  HarmonicState &H = c->hs[harmonic - 1];
  cplx *XP = H.mode == 1 ? H.pplan.XP : nullptr, *XS = H.mode == 1 ? H.pplan.XS : nullptr;
  launch_arnoldi_step(c->V.p, N, j, c->w.p, N, mode, c->h_dev.p, c->arn_partial.p, c->arn_sync.p,
                      c->V.p + (size_t)(j + 1) * N, XP, XS, H.n, c->sm_count, c->st);
  c->launches += 1;
  if(c->trace && c->mv_ev_used >= 4) {
    cudaEventRecord(c->mv_ev[c->mv_ev_used - 1], c->st);
    c->mv_ev_arn[c->mv_ev_used / 4 - 1] = 1;
  }
}

static double dev_norm(ob_ctx *c, const cplx *v, int N) {
  launch_multi_dot(v, 0, 1, v, N, c->h_dev.p, c->dot_scratch.p, c->st);
  c->launches += 2;
  cplx r;
  OB_CUDA(cudaMemcpyAsync(&r, c->h_dev.p, sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
  OB_CUDA(cudaStreamSynchronize(c->st));
  return std::sqrt(r.x);
}

struct GmresOut {
  int iters = 0;
  double relres = 0;
  bool converged = false;
};

// Gmres_Zcomp (srcAna/PreconditionedMatrix.cpp:892-985) with the dense device operator.
static GmresOut gmres_zcomp(ob_ctx *c, int harmonic, const cplx *Y, cplx *x, double tol, int maxit, int no_rest) {
  const int N = c->N(harmonic);
  ensure_gmres_ws(c, N, maxit);
  const bool fused = use_fused(c, N, maxit);
  cplx *V = c->V.p, *w = c->w.p;
  OB_CUDA(cudaMemsetAsync(x, 0, (size_t)N * sizeof(cplx), c->st));
  const double abs_y = dev_norm(c, Y, N);
  GmresOut out;
  double err_n = 1.0;
  bool x_zero = true;
  std::vector<hcd> hcol, cs, sn, gi, ym;
  std::vector<std::vector<hcd>> Rcols;
  for(int rest = 1; rest <= no_rest; ++rest) {
    if(err_n <= tol)
      break;
    if(x_zero) { // S * 0 = 0: res = Y
      OB_CUDA(cudaMemcpyAsync(w, Y, (size_t)N * sizeof(cplx), cudaMemcpyDeviceToDevice, c->st));
    } else {
      matvec(c, harmonic, x, c->tmpA.p);
      launch_axpby(mk(1, 0), Y, mk(-1, 0), c->tmpA.p, w, N, c->st);
      c->launches += 1;
    }
    const double beta = dev_norm(c, w, N);
    launch_scale_to(w, 1.0 / beta, V, N, c->st);
    c->launches += 1;
    cs.clear();
    sn.clear();
    Rcols.clear();
    gi.assign(1, hcd(beta, 0));
    int n = 0;
    err_n = 1.0;
    bool staged = false; // v_n already staged for the pair operator by the previous fused step
    while(n < maxit && err_n > tol) {
      matvec(c, harmonic, V + (size_t)n * N, w, staged);
      // modified Gram-Schmidt (:939-943): h_t = v_t^H w ; w -= h_t v_t, sequentially; then v_{n+1} = w / ||w||
      if(fused) {
        fused_step(c, harmonic, n, 1);
        staged = c->hs[harmonic - 1].mode == 1;
      } else {
        for(int t = 0; t <= n; ++t) {
          launch_multi_dot(V + (size_t)t * N, 0, 1, w, N, c->h_dev.p + t, c->dot_scratch.p, c->st);
          launch_multi_axpy(V + (size_t)t * N, 0, 1, c->h_dev.p + t, w, N, c->st);
          c->launches += 3;
        }
        launch_multi_dot(w, 0, 1, w, N, c->h_dev.p + n + 1, c->dot_scratch.p, c->st);
        c->launches += 2;
      }
      OB_CUDA(cudaMemcpyAsync(c->h_pinned, c->h_dev.p, (size_t)(n + 2) * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
      OB_CUDA(cudaStreamSynchronize(c->st));
      hcol.assign(c->h_pinned, c->h_pinned + n + 2);
      const double hn = std::sqrt(hcol[n + 1].real());
      hcol[n + 1] = hn;
      if(!fused) {
        launch_scale_to(w, 1.0 / hn, V + (size_t)(n + 1) * N, N, c->st);
        c->launches += 1;
      }
      // Givens exactly as det_approx (:1087-1133): W = [[conj(c), conj(-s)], [s, c]]
      for(int i = 0; i < n; ++i) {
        hcd a = hcol[i], b = hcol[i + 1];
        hcol[i] = std::conj(cs[i]) * a + std::conj(-sn[i]) * b;
        hcol[i + 1] = sn[i] * a + cs[i] * b;
      }
      hcd cc, ss, temp;
      if(std::abs(hcol[n + 1]) > std::abs(hcol[n])) {
        temp = hcol[n] / hcol[n + 1];
        ss = 1.0 / std::sqrt(1.0 + std::pow(std::abs(temp), 2));
        cc = -temp * ss;
      } else {
        temp = hcol[n + 1] / hcol[n];
        cc = 1.0 / std::sqrt(1.0 + std::pow(std::abs(temp), 2));
        ss = -temp * cc;
      }
      cs.push_back(cc);
      sn.push_back(ss);
      {
        hcd a = hcol[n], b = hcol[n + 1];
        hcol[n] = std::conj(cc) * a + std::conj(-ss) * b;
        hcol[n + 1] = ss * a + cc * b;
      }
      gi.push_back(hcd(0, 0));
      {
        hcd a = gi[n], b = gi[n + 1];
        gi[n] = std::conj(cc) * a + std::conj(-ss) * b;
        gi[n + 1] = ss * a + cc * b;
      }
      Rcols.push_back(hcol);
      err_n = std::abs(gi[n + 1]) / abs_y;
      ++n;
      ++out.iters;
    }
    ym.assign(n, hcd(0, 0));
    for(int i = n - 1; i >= 0; --i) {
      hcd s = gi[i];
      for(int j = i + 1; j < n; ++j)
        s -= Rcols[j][i] * ym[j];
      ym[i] = s / Rcols[i][i];
    }
    if(n > 0) {
      OB_CUDA(cudaMemcpyAsync(c->ycoef.p, ym.data(), (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, c->st));
      launch_combine(V, N, n, c->ycoef.p, x, N, c->st);
      c->launches += 1;
      OB_CUDA(cudaStreamSynchronize(c->st)); // ym is reused
      x_zero = false;
    }
  }
  out.relres = err_n;
  out.converged = err_n <= tol;
  return out;
}

// Belos "GMRES" restated (see include/optimet_b200.h): x0 = b, DGKS, implicit residual / ||r0||.
static GmresOut gmres_belos(ob_ctx *c, int harmonic, const cplx *b, cplx *x, double tol, int max_iters, int num_blocks,
                            int max_restarts) {
  const int N = c->N(harmonic);
  ensure_gmres_ws(c, N, num_blocks);
  const bool fused = use_fused(c, N, num_blocks);
  cplx *V = c->V.p, *w = c->w.p;
  OB_CUDA(cudaMemcpyAsync(x, b, (size_t)N * sizeof(cplx), cudaMemcpyDeviceToDevice, c->st));
  GmresOut out;
  double r0norm = -1, rel = 1;
  std::vector<hcd> h, hh, cs, sn, g, ym;
  std::vector<std::vector<hcd>> Rcols;
  for(int cycle = 0; cycle <= max_restarts && !out.converged && out.iters < max_iters; ++cycle) {
    matvec(c, harmonic, x, c->tmpA.p);
    launch_axpby(mk(1, 0), b, mk(-1, 0), c->tmpA.p, w, N, c->st);
    c->launches += 1;
    const double beta = dev_norm(c, w, N);
    if(r0norm < 0)
      r0norm = beta;
    if(r0norm == 0.0 || beta / r0norm <= tol) {
      out.converged = true;
      rel = r0norm == 0.0 ? 0.0 : beta / r0norm;
      break;
    }
    launch_scale_to(w, 1.0 / beta, V, N, c->st);
    c->launches += 1;
    cs.clear();
    sn.clear();
    Rcols.clear();
    g.assign(1, hcd(beta, 0));
    int j = 0;
    bool staged = false; // v_j already staged for the pair operator by the previous fused step
    bool in_flight = false; // the product with v_j was enqueued ahead of the host's convergence test of step j - 1
    double rel_m1 = 1.0, rel_m2 = 1.0; // implicit residuals of the two previous steps of this cycle
    while(j < num_blocks && out.iters < max_iters) {
      if(!in_flight)
        matvec(c, harmonic, V + (size_t)j * N, w, staged);
      in_flight = false;
      double norm_after;
      if(fused) {
        // classical Gram-Schmidt + DGKS second pass (decided on the device) + normalisation: one launch
        fused_step(c, harmonic, j, 0);
        staged = c->hs[harmonic - 1].mode == 1;
        OB_CUDA(cudaMemcpyAsync(c->h_pinned, c->h_dev.p, (size_t)(j + 3) * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
        OB_CUDA(cudaEventRecord(c->evh, c->st));
        // The host's part of a step (Hessenberg column read-back, Givens rotations, convergence test) sits between two
        // operator applies.  v_{j+1} is already on the device, so the next apply is enqueued BEFORE waiting for the
        // read-back whenever the residual history predicts that step j does not converge (geometric extrapolation of the
        // last two implicit residuals, one decade of margin); a wrong prediction costs one product whose result is never
        // used (it only writes the work vector w).  The iterates, the iteration count and every result are unchanged.
        if(c->speculate && !c->trace && j >= 2 && j + 1 < num_blocks && out.iters + 1 < max_iters &&
           rel_m1 * std::min(1.0, rel_m1 / rel_m2) > 10.0 * tol) {
          matvec(c, harmonic, V + (size_t)(j + 1) * N, w, staged);
          in_flight = true;
        }
        OB_CUDA(cudaEventSynchronize(c->evh));
        h.assign(c->h_pinned, c->h_pinned + j + 3);
        norm_after = std::sqrt(h[j + 2].real());
        h.resize(j + 2);
      } else {
        // pass 1: classical Gram-Schmidt, all dots at once (+ ||w||^2 as the last "dot")
        launch_multi_dot(V, N, j + 1, w, N, c->h_dev.p, c->dot_scratch.p, c->st);
        launch_multi_dot(w, 0, 1, w, N, c->h_dev.p + j + 1, c->dot_scratch.p, c->st);
        launch_multi_axpy(V, N, j + 1, c->h_dev.p, w, N, c->st);
        launch_multi_dot(w, 0, 1, w, N, c->h_dev.p + j + 2, c->dot_scratch.p, c->st);
        c->launches += 7;
        h.assign(j + 3, hcd(0, 0));
        OB_CUDA(cudaMemcpyAsync(h.data(), c->h_dev.p, (size_t)(j + 3) * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
        OB_CUDA(cudaStreamSynchronize(c->st));
        const double norm_before = std::sqrt(h[j + 1].real());
        norm_after = std::sqrt(h[j + 2].real());
        h.resize(j + 2);
        if(norm_after < 0.70710678118654752440 * norm_before) { // DGKS second pass
          launch_multi_dot(V, N, j + 1, w, N, c->h_dev.p, c->dot_scratch.p, c->st);
          launch_multi_axpy(V, N, j + 1, c->h_dev.p, w, N, c->st);
          launch_multi_dot(w, 0, 1, w, N, c->h_dev.p + j + 1, c->dot_scratch.p, c->st);
          c->launches += 5;
          hh.assign(j + 2, hcd(0, 0));
          OB_CUDA(cudaMemcpyAsync(hh.data(), c->h_dev.p, (size_t)(j + 2) * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
          OB_CUDA(cudaStreamSynchronize(c->st));
          for(int t = 0; t <= j; ++t)
            h[t] += hh[t];
          norm_after = std::sqrt(hh[j + 1].real());
        }
        launch_scale_to(w, 1.0 / norm_after, V + (size_t)(j + 1) * N, N, c->st);
        c->launches += 1;
      }
      h[j + 1] = norm_after;
      for(int i = 0; i < j; ++i) {
        hcd a = h[i], bb = h[i + 1];
        h[i] = cs[i] * a + sn[i] * bb;
        h[i + 1] = -std::conj(sn[i]) * a + cs[i] * bb;
      }
      hcd f = h[j], gg = h[j + 1], cc, ss;
      if(gg == hcd(0, 0)) {
        cc = 1;
        ss = 0;
      } else if(f == hcd(0, 0)) {
        cc = 0;
        ss = std::conj(gg) / std::abs(gg);
      } else {
        double d = std::sqrt(std::norm(f) + std::norm(gg));
        cc = std::abs(f) / d;
        ss = (f / std::abs(f)) * std::conj(gg) / d;
      }
      cs.push_back(cc);
      sn.push_back(ss);
      h[j] = cc * f + ss * gg;
      h[j + 1] = 0;
      g.push_back(hcd(0, 0));
      hcd ga = g[j];
      g[j] = cc * ga;
      g[j + 1] = -std::conj(ss) * ga;
      Rcols.push_back(h);
      ++j;
      ++out.iters;
      rel = std::abs(g[j]) / r0norm;
      rel_m2 = rel_m1;
      rel_m1 = rel;
      if(rel <= tol) {
        out.converged = true;
        break;
      }
    }
    ym.assign(j, hcd(0, 0));
    for(int i = j - 1; i >= 0; --i) {
      hcd s = g[i];
      for(int k = i + 1; k < j; ++k)
        s -= Rcols[k][i] * ym[k];
      ym[i] = s / Rcols[i][i];
    }
    if(j > 0) {
      OB_CUDA(cudaMemcpyAsync(c->ycoef.p, ym.data(), (size_t)j * sizeof(cplx), cudaMemcpyHostToDevice, c->st));
      launch_combine(V, N, j, c->ycoef.p, x, N, c->st);
      c->launches += 1;
      OB_CUDA(cudaStreamSynchronize(c->st));
    }
  }
  out.relres = rel;
  return out;
}

// Direct dense solve: the serial reference's S.colPivHouseholderQr().solve(Q) (PreconditionedMatrixSolver.h:58,75) and
// the pzgesv_ route (ScalapackSolver.cpp:54-128).  The dense matrix is assembled into a work buffer (reference layout,
// -T_i [[A^T,B^T],[B^T,A^T]] with identity diagonal blocks), factorised in place and released.
static GmresOut direct_solve(ob_ctx *c, int harmonic, const cplx *rhs, cplx *x) {
  check_harmonic(harmonic);
  need(c->world == 1, "direct solve runs on one GPU (use a GMRES flavour when the matrix is row-sharded)");
  need(c->nobj > 0, "ob_set_cluster has not been called");
  ensure_factors(c);
  HarmonicState &H = c->hs[harmonic - 1];
  const size_t N = (size_t)c->N(harmonic);
  const cplx k = harmonic == 1 ? c->waveK : cscale(c->waveK, 2.0);
  if(c->lu_mat.n < N * N) {
    c->lu_mat.release();
    size_t free_b = 0, total_b = 0;
    OB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if((double)N * (double)N * 16.0 + 64e6 > (double)free_b)
      throw Error("direct solve needs 16 N^2 = " + std::to_string(16.0 * N * N / 1e9) + " GB of device memory, " +
                  std::to_string(free_b / 1e9) + " GB free: use a GMRES flavour");
    c->lu_mat.alloc(N * N);
  }
  VtacTableSet &ts = tables_for(c, H.nMax);
  launch_assemble(ts, c->xyz.p, c->fac[harmonic == 1 ? 0 : 1].p, k, c->nobj, 0, c->nobj, c->lu_mat.p, N, c->st);
  c->launches += 1;
  const int info = lu_solve(c->lu_mat.p, (int)N, N, c->lu, rhs, x, c->sm_count, c->st, c->launches);
  if(!c->keep_matrices)
    c->lu_mat.release();
  if(info != 0)
    throw Error("direct solve: the scattering matrix is singular (zero pivot at column " + std::to_string(info) + ")");
  GmresOut r;
  r.converged = true;
  return r;
}

static GmresOut solve_dev(ob_ctx *c, int harmonic, const cplx *rhs, cplx *x, const ob_gmres_opts *o) {
  need(o != nullptr, "ob_gmres_opts is NULL");
  if(o->flavour == OB_SOLVE_DIRECT)
    return direct_solve(c, harmonic, rhs, x);
  // the public ABI accepts any values: reject what the drivers cannot run (the Hessenberg column of one iteration is
  // staged in a pinned buffer grown to the basis size here)
  need(o->tol > 0.0, "ob_gmres_opts: tol must be positive");
  need(o->max_iters > 0, "ob_gmres_opts: max_iters must be positive");
  need(o->max_restarts >= 0, "ob_gmres_opts: max_restarts must not be negative");
  need(o->flavour != OB_GMRES_BELOS || o->restart > 0, "ob_gmres_opts: restart (Num Blocks) must be positive");
  {
    const size_t basis = (size_t)(o->flavour == OB_GMRES_BELOS ? std::min(o->restart, o->max_iters) : o->max_iters) + 8;
    if(basis > c->h_pinned_cap) {
      if(c->h_pinned)
        cudaFreeHost(c->h_pinned);
      c->h_pinned = nullptr;
      c->h_pinned_cap = 0;
      OB_CUDA(cudaMallocHost(&c->h_pinned, basis * sizeof(hcd)));
      c->h_pinned_cap = basis;
    }
  }
  c->tmpA.alloc(c->N(harmonic));
  GmresOut r;
  struct Flush { // matvec timings are read back when the solve ends (also on the error path)
    ob_ctx *c;
    ~Flush() { flush_matvec_timing(c); }
  } flush_guard{c};
  if(o->flavour == OB_GMRES_ZCOMP)
    r = gmres_zcomp(c, harmonic, rhs, x, o->tol, o->max_iters, o->max_restarts);
  else if(o->flavour == OB_GMRES_BELOS) {
    r = gmres_belos(c, harmonic, rhs, x, o->tol, o->max_iters, o->restart, o->max_restarts);
    if(!r.converged) // srcAna/MatrixBelosSolver.cpp:60-61
      throw Error("Error encountered while solving the linear system");
  } else
    throw Error("unknown GMRES flavour");
  return r;
}

static void source_ff(ob_ctx *c) {
  need(c->have_inc, "ob_set_incident has not been called");
  ensure_factors(c);
  const int N = c->N(1), blk = 2 * c->hs[0].n;
  c->Q.alloc(N);
  VtacTableSet &ts = tables_for(c, c->nMax);
  launch_translate_apply(ts, c->xyz.p, c->waveK, c->first, c->count, c->ainc.p, 0, c->fac[0].p,
                         c->Q.p + (size_t)c->first * blk, c->st);
  c->launches += 1;
  allgather_slices(c, c->Q.p, blk);
}

static ShInputs sh_inputs(ob_ctx *c) {
  ShInputs in;
  in.nobj = c->nobj;
  in.nMax = c->nMax;
  in.nMaxS = c->nMaxS;
  in.omega = c->omega;
  in.eps_b = c->eps_b;
  in.mu_b = c->mu_b;
  in.radius = c->radius.p;
  in.eps = c->mat[0].p;
  in.mu = c->mat[1].p;
  in.eps_SH = c->mat[2].p;
  in.mu_SH = c->mat[3].p;
  in.ksippp = c->mat[4].p;
  in.ksiparppar = c->mat[5].p;
  in.gamma = c->mat[6].p;
  for(int t = 0; t < 9; ++t)
    in.tab[t] = c->cg[t].p;
  return in;
}

static void ensure_cg(ob_ctx *c) {
  if(c->cg_nmax == c->nMax * 1000 + c->nMaxS)
    return;
  const size_t sz = (size_t)flat_max(c->nMaxS) * flat_max(c->nMax) * flat_max(c->nMax);
  double *T[9];
  for(int t = 0; t < 9; ++t) {
    c->cg[t].alloc(sz);
    T[t] = c->cg[t].p;
  }
  launch_cg_tables(c->nMax, c->nMaxS, T, c->st);
  c->launches += 2;
  c->cg_nmax = c->nMax * 1000 + c->nMaxS;
}

// K, K1ana from the conjugated FF internal coefficients (device, full length)
static void source_sh(ob_ctx *c, const cplx *Xint_conj) {
  ensure_factors(c);
  ensure_cg(c);
  const int N = c->N(2), blk = 2 * c->hs[1].n;
  c->Ksrc.alloc(N);
  c->K1ana.alloc(N);
  ShInputs in = sh_inputs(c);
  launch_sh_source(in, c->first, c->count, Xint_conj, c->fac[2].p, c->fac[3].p, c->fac[6].p, c->Ksrc.p, c->K1ana.p,
                   c->st);
  c->launches += 1;
  allgather_slices(c, c->Ksrc.p, blk);
  allgather_slices(c, c->K1ana.p, blk);
}

// Result.cpp:557-794; device vectors; partial sums over local particles, gathered on the host side
// extinction sum of one particle, Re sum_p conj(Q_local[p]) X_sca[p] (Result.cpp:564-571): one thread per local particle,
// the terms in the order and with the roundings of the host loop this kernel replaces (no contraction into FMAs), so the
// 2 x 16 N bytes of coefficients no longer travel to the host for it
__global__ void k_ext_terms(const cplx *__restrict__ q, const cplx *__restrict__ x, int n, int count, double *__restrict__ out) {
  const int jl = blockIdx.x * blockDim.x + threadIdx.x;
  if(jl >= count)
    return;
  const cplx *qq = q + (size_t)jl * 2 * n, *xx = x + (size_t)jl * 2 * n;
  double e = 0.0;
  for(int p = 0; p < n; ++p) {
    const cplx q1 = qq[p], x1 = xx[p], q2 = qq[p + n], x2 = xx[p + n];
    const double r1 = __dadd_rn(__dmul_rn(q1.x, x1.x), __dmul_rn(q1.y, x1.y));
    const double r2 = __dadd_rn(__dmul_rn(q2.x, x2.x), __dmul_rn(q2.y, x2.y));
    e = __dadd_rn(e, __dadd_rn(r1, r2));
  }
  out[jl] = e;
}

static void cross_sections(ob_ctx *c, const cplx *Xsca, const cplx *Xint, const cplx *XscaSH, const cplx *XintSH,
                           bool do_sh, double cs[5]) {
  const int n = c->hs[0].n, blk = 2 * n;
  std::vector<double> part; // per particle (global index), summed in particle order as the reference does
  c->red_d.alloc(3 * (size_t)c->nobj + 16);
  c->red_c.alloc((size_t)c->nobj + 16);
  VtacTableSet &ts = tables_for(c, c->nMax);
  // extinction: Q_local = getIncLocal(R_j)  (Result.cpp:564-571)
  c->tmpB.alloc(std::max(c->N(1), c->N(2)));
  launch_translate_apply(ts, c->xyz.p, c->waveK, c->first, c->count, c->ainc.p, 0, nullptr,
                         c->tmpB.p + (size_t)c->first * blk, c->st);
  launch_sca_sum(ts, c->xyz.p, c->waveK, c->first, c->count, Xsca, c->red_d.p + c->first, c->st);
  c->launches += 2;
  std::vector<double> sca(c->nobj, 0.0), scaSH(c->nobj, 0.0), ext(c->nobj, 0.0), absSH(c->nobj, 0.0);
  if(c->count > 0) {
    k_ext_terms<<<(c->count + 127) / 128, 128, 0, c->st>>>(c->tmpB.p + (size_t)c->first * blk, Xsca + (size_t)c->first * blk, n,
                                                           c->count, c->red_d.p + 2 * (size_t)c->nobj + c->first);
    OB_CUDA(cudaGetLastError());
    c->launches += 1;
  }
  OB_CUDA(cudaMemcpyAsync(ext.data() + c->first, c->red_d.p + 2 * (size_t)c->nobj + c->first, c->count * sizeof(double),
                          cudaMemcpyDeviceToHost, c->st));
  OB_CUDA(cudaMemcpyAsync(sca.data() + c->first, c->red_d.p + c->first, c->count * sizeof(double),
                          cudaMemcpyDeviceToHost, c->st));
  if(do_sh) {
    VtacTableSet &tsS = tables_for(c, c->nMaxS);
    launch_sca_sum(tsS, c->xyz.p, cscale(c->waveK, 2.0), c->first, c->count, XscaSH, c->red_d.p + c->nobj + c->first,
                   c->st);
    ensure_cg(c);
    ShInputs in = sh_inputs(c);
    launch_abs_sh(in, c->first, c->count, Xint, XintSH, c->red_c.p + c->first, c->st);
    c->launches += 2;
    OB_CUDA(cudaMemcpyAsync(scaSH.data() + c->first, c->red_d.p + c->nobj + c->first, c->count * sizeof(double),
                            cudaMemcpyDeviceToHost, c->st));
  }
  std::vector<hcd> acs(c->nobj, hcd(0, 0));
  if(do_sh)
    OB_CUDA(cudaMemcpyAsync(acs.data() + c->first, c->red_c.p + c->first, c->count * sizeof(cplx),
                            cudaMemcpyDeviceToHost, c->st));
  OB_CUDA(cudaStreamSynchronize(c->st));
  for(int jl = 0; jl < c->count; ++jl) {
    if(do_sh) {
      const double mu0 = 4.0 * 3.14159265358979323846 * 1e-7;
      const double eps0 = 1.0 / (mu0 * 299792458.0 * 299792458.0);
      hcd eta = std::sqrt(hcd(c->mu_b.x, c->mu_b.y) / hcd(c->eps_b.x, c->eps_b.y));
      hcd sigma = -hcd(0.0, 1.0) * eps0 * 2.0 * c->omega * (c->h_epsr_SH[c->first + jl] - 1.0); // Result.cpp:784
      absSH[c->first + jl] = std::real((2.0 * eta) * 0.5 * sigma * acs[c->first + jl]);          // :788
    }
  }
  // the four per-particle partial arrays are summed over ranks on the host by the caller (world > 1:
  // each rank returns the partial sums of its own particles; bench/host adaptor adds them).
  double Cext = 0, Csca = 0, CscaSH = 0, CabsSH = 0;
  for(int j = 0; j < c->nobj; ++j) {
    Cext += ext[j];
    Csca += sca[j];
    CscaSH += scaSH[j];
    CabsSH = CabsSH + absSH[j];
  }
  const double k2 = c->waveK.x * c->waveK.x;
  const double mu0 = 4.0 * 3.14159265358979323846 * 1e-7;
  const double eps0 = 1.0 / (mu0 * 299792458.0 * 299792458.0);
  cs[0] = (-1. / k2) * Cext;
  cs[1] = (1. / k2) * Csca;
  cs[2] = cs[0] - cs[1];
  const double ArbCf = (c->eps_b.x / eps0) * (c->mu_b.x / mu0);
  cs[3] = do_sh ? (1.0 / (4.0 * ArbCf)) * CscaSH : 0.0;
  cs[4] = do_sh ? CabsSH : 0.0;
}

static void upload(ob_ctx *c, DevBuf<cplx> &buf, const double *host, size_t n) {
  buf.alloc(n);
  OB_CUDA(cudaMemcpyAsync(buf.p, host, n * sizeof(cplx), cudaMemcpyHostToDevice, c->st));
}
static void download(ob_ctx *c, const cplx *dev, double *host, size_t n) {
  if(!host)
    return;
  OB_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
  OB_CUDA(cudaStreamSynchronize(c->st));
}

struct PhaseTimer {
  ob_ctx *c;
  int slot;
  PhaseTimer(ob_ctx *c_, int s) : c(c_), slot(s) { cudaEventRecord(c->ev0, c->st); }
  void stop() {
    cudaEventRecord(c->ev1, c->st);
    cudaEventSynchronize(c->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->tim[slot] += ms;
  }
};

} // namespace ob

#define OB_BEGIN                                                                                                       \
  if(!ctx)                                                                                                             \
    return 1;                                                                                                          \
  try {                                                                                                                \
    OB_CUDA(cudaSetDevice(ctx->device));
#define OB_END                                                                                                         \
  }                                                                                                                    \
  catch(std::exception & e) {                                                                                          \
    ctx->err = e.what();                                                                                               \
    return 1;                                                                                                          \
  }                                                                                                                    \
  return 0;

extern "C" {

int ob_create(int device, ob_ctx **out) {
  try {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if(e != cudaSuccess || ndev == 0)
      throw Error("no CUDA device available: the B200 path has no CPU fallback");
    if(device < 0 || device >= ndev)
      throw Error("invalid device ordinal");
    OB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    OB_CUDA(cudaGetDeviceProperties(&prop, device));
    if(prop.major < 10)
      throw Error("this library is built for sm_100a (B200) only");
    ob_ctx *c = new ob_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    OB_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    OB_CUDA(cudaEventCreate(&c->ev0));
    OB_CUDA(cudaEventCreate(&c->ev1));
    OB_CUDA(cudaEventCreate(&c->evm0));
    OB_CUDA(cudaEventCreate(&c->evm1));
    OB_CUDA(cudaEventCreate(&c->evt0));
    OB_CUDA(cudaEventCreate(&c->evt1));
    OB_CUDA(cudaEventCreateWithFlags(&c->evh, cudaEventDisableTiming));
    c->mv_ev.resize(256);
    c->mv_ev_arn.assign(64, 0);
    for(auto &e : c->mv_ev)
      OB_CUDA(cudaEventCreate(&e));
    OB_CUDA(cudaMallocHost(&c->h_pinned, 512 * sizeof(hcd)));
    c->h_pinned_cap = 512;
    *out = c;
  } catch(std::exception &e) {
    g_create_error = e.what();
    *out = nullptr;
    return 1;
  }
  return 0;
}

void ob_destroy(ob_ctx *ctx) {
  if(!ctx)
    return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->st);
  if(ctx->comm && g_nccl.CommDestroy)
    g_nccl.CommDestroy(ctx->comm);
  for(int i = 0; i < 2; ++i) {
    ctx->tabs[i].release();
    matvec_plan_release(ctx->hs[i].plan);
    pair_plan_release(ctx->hs[i].pplan);
    rot_plan_release(ctx->hs[i].rplan);
    ctx->hs[i].aca.release();
  }
  ctx->lu.release();
  ctx->aca_scratch.release();
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->evm0);
  cudaEventDestroy(ctx->evm1);
  cudaEventDestroy(ctx->evt0);
  cudaEventDestroy(ctx->evt1);
  for(auto &e : ctx->mv_ev)
    cudaEventDestroy(e);
  if(ctx->h_pinned)
    cudaFreeHost(ctx->h_pinned);
  cudaStream_t st = ctx->st;
  delete ctx;
  cudaStreamDestroy(st);
}

const char *ob_last_error(ob_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int ob_device_info(ob_ctx *ctx, int *sm_count, size_t *free_bytes, size_t *total_bytes) {
  OB_BEGIN
  *sm_count = ctx->sm_count;
  OB_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
  OB_END
}

int ob_partition(int nobj, int world, int rank, int *first, int *count) {
  if(nobj < 0 || world < 1 || rank < 0 || rank >= world)
    return 1;
  partition(nobj, world, rank, *first, *count);
  return 0;
}

int ob_comm_unique_id(char out[128]) {
  if(!g_nccl.load())
    return 1;
  ncclUniqueId id;
  if(g_nccl.GetUniqueId(&id) != ncclSuccess)
    return 1;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out, &id, 128);
  return 0;
}

int ob_comm_init(ob_ctx *ctx, const char uid[128], int rank, int world) {
  OB_BEGIN
  need(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
  ctx->rank = rank;
  ctx->world = world;
  if(world > 1) {
    need(g_nccl.load(), "libnccl.so.2 not found");
    ncclUniqueId id;
    memcpy(&id, uid, 128);
    OB_NCCL(g_nccl.CommInitRank(&ctx->comm, world, id, rank));
  }
  if(ctx->nobj > 0)
    partition(ctx->nobj, world, rank, ctx->first, ctx->count);
  for(int h = 0; h < 2; ++h) { // operators assembled under another partition are stale
    ctx->hs[h].assembled = false;
    ctx->hs[h].pplan_world = ctx->hs[h].rplan_world = -1;
  }
  OB_END
}

int ob_set_shard(ob_ctx *ctx, int rank, int world) {
  OB_BEGIN
  need(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
  need(ctx->comm == nullptr, "ob_set_shard: the context already has a communicator");
  ctx->rank = rank;
  ctx->world = world;
  if(ctx->nobj > 0)
    partition(ctx->nobj, world, rank, ctx->first, ctx->count);
  for(int h = 0; h < 2; ++h) { // operators assembled for another shard are stale
    ctx->hs[h].assembled = false;
    ctx->hs[h].pplan_world = ctx->hs[h].rplan_world = -1;
  }
  OB_END
}

int ob_set_cluster(ob_ctx *ctx, int nobj, const double *xyz_m, const double *radius_m, int nMax, int nMaxS) {
  OB_BEGIN
  need(nobj > 0, "No scatterers defined in input");
  need(nMax >= 1 && nMax <= OB_MAX_NMAX && nMaxS >= 1 && nMaxS <= OB_MAX_NMAX,
       "nMax out of range (1..13 supported by the shared-memory VTAC kernel)");
  // the same cluster as last time (a wavelength sweep calls update() per wavelength): nothing to check or upload again
  if(nobj == ctx->nobj && nMax == ctx->nMax && nMaxS == ctx->nMaxS && ctx->h_xyz.size() == 3 * (size_t)nobj &&
     std::memcmp(ctx->h_xyz.data(), xyz_m, 3 * (size_t)nobj * sizeof(double)) == 0 &&
     std::memcmp(ctx->h_radius.data(), radius_m, (size_t)nobj * sizeof(double)) == 0) {
    ctx->fac_valid = false;
    ctx->hs[0].assembled = ctx->hs[1].assembled = false;
    ctx->have_inc = false;
    return 0;
  }
  // overlap check of Geometry::pushObject (srcAna/Geometry.cpp:39-52), O(N^2) on the host only for small clusters
  if(nobj <= 4096)
    for(int i = 0; i < nobj; ++i)
      for(int j = 0; j < i; ++j) {
        double dx = xyz_m[3 * i] - xyz_m[3 * j], dy = xyz_m[3 * i + 1] - xyz_m[3 * j + 1],
               dz = xyz_m[3 * i + 2] - xyz_m[3 * j + 2];
        if(std::sqrt(dx * dx + dy * dy + dz * dz) <= radius_m[i] + radius_m[j])
          throw Error("The sphere at index " + std::to_string(i) + " overlaps with the one at index " +
                      std::to_string(j));
      }
  ctx->nobj = nobj;
  ctx->nMax = nMax;
  ctx->nMaxS = nMaxS;
  ctx->hs[0].nMax = nMax;
  ctx->hs[0].n = flat_max(nMax);
  ctx->hs[1].nMax = nMaxS;
  ctx->hs[1].n = flat_max(nMaxS);
  partition(nobj, ctx->world, ctx->rank, ctx->first, ctx->count);
  ctx->h_xyz.assign(xyz_m, xyz_m + 3 * (size_t)nobj);
  ctx->h_radius.assign(radius_m, radius_m + nobj);
  ctx->xyz.alloc(3 * (size_t)nobj);
  ctx->radius.alloc(nobj);
  OB_CUDA(cudaMemcpyAsync(ctx->xyz.p, xyz_m, 3 * (size_t)nobj * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  OB_CUDA(cudaMemcpyAsync(ctx->radius.p, radius_m, (size_t)nobj * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  OB_CUDA(cudaStreamSynchronize(ctx->st));
  ctx->fac_valid = false;
  ctx->hs[0].assembled = ctx->hs[1].assembled = false;
  ctx->have_inc = false;
  OB_END
}

int ob_set_frequency(ob_ctx *ctx, double omega, const double waveK[2], const double eps_b[2], const double mu_b[2],
                     const double *eps, const double *mu, const double *eps_SH, const double *mu_SH,
                     const double *ksippp, const double *ksiparppar, const double *gamma) {
  OB_BEGIN
  need(ctx->nobj > 0, "ob_set_cluster has not been called");
  ctx->omega = omega;
  ctx->waveK = mk(waveK[0], waveK[1]);
  ctx->eps_b = mk(eps_b[0], eps_b[1]);
  ctx->mu_b = mk(mu_b[0], mu_b[1]);
  const double *src[7] = {eps, mu, eps_SH, mu_SH, ksippp, ksiparppar, gamma};
  for(int i = 0; i < 7; ++i) {
    need(src[i] != nullptr, "ob_set_frequency: NULL material array");
    upload(ctx, ctx->mat[i], src[i], ctx->nobj);
  }
  const double mu0 = 4.0 * 3.14159265358979323846 * 1e-7;
  const double eps0 = 1.0 / (mu0 * 299792458.0 * 299792458.0);
  ctx->h_epsr_SH.resize(ctx->nobj);
  for(int j = 0; j < ctx->nobj; ++j)
    ctx->h_epsr_SH[j] = hcd(eps_SH[2 * j], eps_SH[2 * j + 1]) / eps0;
  OB_CUDA(cudaStreamSynchronize(ctx->st));
  ctx->have_freq = true;
  ctx->fac_valid = false;
  ctx->hs[0].assembled = ctx->hs[1].assembled = false;
  OB_END
}

int ob_set_incident(ob_ctx *ctx, const double *a_origin, const double *b_origin) {
  OB_BEGIN
  need(ctx->nobj > 0, "ob_set_cluster has not been called");
  const int n = ctx->hs[0].n;
  ctx->ainc.alloc(2 * n);
  OB_CUDA(cudaMemcpyAsync(ctx->ainc.p, a_origin, n * sizeof(cplx), cudaMemcpyHostToDevice, ctx->st));
  OB_CUDA(cudaMemcpyAsync(ctx->ainc.p + n, b_origin, n * sizeof(cplx), cudaMemcpyHostToDevice, ctx->st));
  OB_CUDA(cudaStreamSynchronize(ctx->st));
  ctx->have_inc = true;
  OB_END
}

int ob_vtac(ob_ctx *ctx, const double relR_sph[3], const double k[2], int regular_flag, int nMax, double *A,
            double *B) {
  OB_BEGIN
  need(nMax >= 1 && nMax <= OB_MAX_NMAX, "nMax out of range");
  VtacTableSet &ts = tables_for(ctx, nMax);
  const size_t n = flat_max(nMax);
  DevBuf<cplx> dA, dB;
  dA.alloc(n * n);
  dB.alloc(n * n);
  // Coupling.cpp:86: the ctor flag is inverted when handed to the TA coefficients
  launch_vtac_single(ts, relR_sph[0], relR_sph[1], relR_sph[2], mk(k[0], k[1]), regular_flag ? 0 : 1, dA.p, dB.p,
                     ctx->st);
  ctx->launches += 1;
  download(ctx, dA.p, A, n * n);
  download(ctx, dB.p, B, n * n);
  OB_END
}

int ob_particle_factors(ob_ctx *ctx, int which, double *out) {
  OB_BEGIN
  need(which >= 0 && which < 7, "which must be 0..6");
  ensure_factors(ctx);
  const int nm = (which == 0 || which == 4) ? ctx->nMax : ctx->nMaxS;
  download(ctx, ctx->fac[which].p, out, (size_t)ctx->nobj * 2 * flat_max(nm));
  OB_END
}

int ob_inc_local(ob_ctx *ctx, double *out) {
  OB_BEGIN
  need(ctx->have_inc && ctx->have_freq, "incident field / frequency not set");
  const int N = ctx->N(1), blk = 2 * ctx->hs[0].n;
  ctx->tmpB.alloc(std::max(ctx->N(1), ctx->N(2)));
  VtacTableSet &ts = tables_for(ctx, ctx->nMax);
  launch_translate_apply(ts, ctx->xyz.p, ctx->waveK, ctx->first, ctx->count, ctx->ainc.p, 0, nullptr,
                         ctx->tmpB.p + (size_t)ctx->first * blk, ctx->st);
  ctx->launches += 1;
  allgather_slices(ctx, ctx->tmpB.p, blk);
  download(ctx, ctx->tmpB.p, out, N);
  OB_END
}

int ob_assemble(ob_ctx *ctx, int harmonic) {
  OB_BEGIN
  assemble(ctx, harmonic);
  OB_CUDA(cudaStreamSynchronize(ctx->st));
  OB_END
}

int ob_release_matrix(ob_ctx *ctx, int harmonic) {
  OB_BEGIN
  check_harmonic(harmonic);
  ctx->hs[harmonic - 1].S.release();
  ctx->hs[harmonic - 1].AB.release();
  ctx->hs[harmonic - 1].aca.release();
  rot_records_drop(ctx, harmonic - 1);
  ctx->hs[harmonic - 1].assembled = false;
  OB_END
}

int ob_fetch_block(ob_ctx *ctx, int harmonic, int i, int j, double *out) {
  OB_BEGIN
  check_harmonic(harmonic);
  HarmonicState &H = ctx->hs[harmonic - 1];
  need(H.assembled, "matrix not assembled");
  need(H.mode != 2, "the ACA-compressed operator holds no dense blocks: use ob_aca_block");
  need(H.mode != 3, "the rotated-axial operator holds no dense blocks (operator 0 or 1 rebuild them)");
  if(H.mode == 1) {
    need(i >= 0 && i < ctx->nobj && j >= 0 && j < ctx->nobj, "block index out of range");
    const size_t b2 = (size_t)4 * H.n * H.n;
    ctx->tmpA.alloc(std::max(b2, (size_t)ctx->N(harmonic)));
    launch_pairs_expand_block(H.pplan, H.AB.p, i, j, ctx->fac[harmonic == 1 ? 0 : 1].p, ctx->tmpA.p, ctx->st);
    ctx->launches += 1;
    download(ctx, ctx->tmpA.p, out, b2);
    return 0;
  }
  need(i >= ctx->first && i < ctx->first + ctx->count && j >= 0 && j < ctx->nobj, "block is not local to this rank");
  const size_t b = 2 * H.n, ld = (size_t)ctx->Mloc(harmonic);
  const cplx *src = H.S.p + (size_t)j * b * ld + (size_t)(i - ctx->first) * b;
  OB_CUDA(cudaMemcpy2DAsync(out, b * sizeof(cplx), src, ld * sizeof(cplx), b * sizeof(cplx), b, cudaMemcpyDeviceToHost,
                            ctx->st));
  OB_CUDA(cudaStreamSynchronize(ctx->st));
  OB_END
}

int ob_fetch_matrix(ob_ctx *ctx, int harmonic, double *out) {
  OB_BEGIN
  check_harmonic(harmonic);
  HarmonicState &H = ctx->hs[harmonic - 1];
  need(H.assembled, "matrix not assembled");
  need(H.mode != 2, "the ACA-compressed operator holds no dense matrix: use ob_aca_block");
  need(H.mode != 3, "the rotated-axial operator holds no dense matrix (operator 0 or 1 rebuild it)");
  if(H.mode == 1) { // rebuild the dense reference layout block by block (tests; single rank only)
    need(ctx->world == 1, "ob_fetch_matrix in pair form needs world == 1");
    const size_t b = 2 * (size_t)H.n, ld = (size_t)ctx->N(harmonic);
    DevBuf<cplx> blk;
    blk.alloc(b * b);
    std::vector<hcd> hb(b * b);
    hcd *o = (hcd *)out;
    for(int i = 0; i < ctx->nobj; ++i)
      for(int j = 0; j < ctx->nobj; ++j) {
        launch_pairs_expand_block(H.pplan, H.AB.p, i, j, ctx->fac[harmonic == 1 ? 0 : 1].p, blk.p, ctx->st);
        OB_CUDA(cudaMemcpyAsync(hb.data(), blk.p, b * b * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->st));
        OB_CUDA(cudaStreamSynchronize(ctx->st));
        for(size_t cc = 0; cc < b; ++cc)
          memcpy(o + ((size_t)j * b + cc) * ld + (size_t)i * b, hb.data() + cc * b, b * sizeof(hcd));
      }
    return 0;
  }
  download(ctx, H.S.p, out, (size_t)ctx->Mloc(harmonic) * ctx->N(harmonic));
  OB_END
}

int ob_aca_compress(ob_ctx *ctx, int dim, const double *C, int *rank, double *U, double *V, int *I, int *J) {
  OB_BEGIN
  need(dim >= 2 && dim <= 2 * OB_MAX_FLAT, "ob_aca_compress: dim out of range (2..390)");
  const size_t b2 = (size_t)dim * dim;
  DevBuf<cplx> dC, dU, dV;
  DevBuf<int> dr, dp;
  dC.alloc(b2);
  dU.alloc(b2);
  dV.alloc(b2);
  dr.alloc(1);
  dp.alloc(2 * (size_t)dim);
  OB_CUDA(cudaMemcpyAsync(dC.p, C, b2 * sizeof(cplx), cudaMemcpyHostToDevice, ctx->st));
  aca_compress_single(dC.p, dim, ctx->eps_aca, dU.p, dV.p, dr.p, dp.p, ctx->st);
  ctx->launches += 1;
  int r = 0;
  OB_CUDA(cudaMemcpy(&r, dr.p, sizeof(int), cudaMemcpyDeviceToHost));
  *rank = r;
  need(r >= 2, "ACA_compression: no admissible pivot (the reference reads an uninitialised index there)");
  OB_CUDA(cudaMemcpy(U, dU.p, (size_t)r * dim * sizeof(cplx), cudaMemcpyDeviceToHost));
  OB_CUDA(cudaMemcpy(V, dV.p, (size_t)r * dim * sizeof(cplx), cudaMemcpyDeviceToHost));
  std::vector<int> pv(2 * (size_t)dim);
  OB_CUDA(cudaMemcpy(pv.data(), dp.p, pv.size() * sizeof(int), cudaMemcpyDeviceToHost));
  for(int p = 0; p < r; ++p) { // pivot rows / columns are optional outputs, as in ob_aca_block
    if(I)
      I[p] = pv[p];
    if(J)
      J[p] = pv[dim + p];
  }
  OB_END
}

int ob_aca_block(ob_ctx *ctx, int harmonic, int i, int j, int *rank, double *U, double *V, int *I, int *J) {
  OB_BEGIN
  check_harmonic(harmonic);
  HarmonicState &H = ctx->hs[harmonic - 1];
  need(H.assembled && H.mode == 2 && H.aca.built, "ACA operator not assembled (operator = 2, ob_assemble)");
  need(i >= ctx->first && i < ctx->first + ctx->count && j >= 0 && j < ctx->nobj, "block is not local to this rank");
  const int dim = H.aca.dim;
  const size_t b = (size_t)(i - ctx->first) * ctx->nobj + j;
  const AcaDesc d = H.aca.h_desc[b];
  *rank = d.rank;
  OB_CUDA(cudaStreamSynchronize(ctx->st));
  if(d.rank < 0) {
    OB_CUDA(cudaMemcpy(U, d.U, (size_t)dim * dim * sizeof(cplx), cudaMemcpyDeviceToHost));
  } else if(d.rank > 0) {
    OB_CUDA(cudaMemcpy(U, d.U, (size_t)d.rank * dim * sizeof(cplx), cudaMemcpyDeviceToHost));
    OB_CUDA(cudaMemcpy(V, d.V, (size_t)d.rank * dim * sizeof(cplx), cudaMemcpyDeviceToHost));
    std::vector<int> pv(2 * (size_t)dim);
    OB_CUDA(cudaMemcpy(pv.data(), H.aca.piv + b * 2 * dim, pv.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for(int p = 0; p < d.rank; ++p) {
      if(I)
        I[p] = pv[p];
      if(J)
        J[p] = pv[dim + p];
    }
  }
  OB_END
}

int ob_aca_stats(ob_ctx *ctx, int harmonic, double out[6]) {
  OB_BEGIN
  check_harmonic(harmonic);
  HarmonicState &H = ctx->hs[harmonic - 1];
  need(H.assembled && H.mode == 2 && H.aca.built, "ACA operator not assembled (operator = 2, ob_assemble)");
  out[0] = 16.0 * H.aca.stored_elems;
  out[1] = 16.0 * (double)ctx->Mloc(harmonic) * (double)ctx->N(harmonic);
  out[2] = (double)H.aca.n_lowrank;
  out[3] = (double)H.aca.n_dense;
  out[4] = H.aca.n_lowrank ? H.aca.rank_sum / (double)H.aca.n_lowrank : 0.0;
  out[5] = (double)H.aca.rank_max;
  OB_END
}

int ob_matvec(ob_ctx *ctx, int harmonic, const double *x, double *y) {
  OB_BEGIN
  check_harmonic(harmonic);
  const int N = ctx->N(harmonic);
  upload(ctx, ctx->tmpA, x, N);
  ctx->tmpB.alloc(std::max(ctx->N(1), ctx->N(2)));
  matvec(ctx, harmonic, ctx->tmpA.p, ctx->tmpB.p);
  download(ctx, ctx->tmpB.p, y, N);
  flush_matvec_timing(ctx);
  OB_END
}

// page-lock caller memory the coefficient vectors are copied into (cudaHostRegister): device -> host at PCIe speed
int ob_host_register(ob_ctx *ctx, void *ptr, size_t bytes) {
  OB_BEGIN
  need(ptr != nullptr && bytes > 0, "ob_host_register: empty range");
  OB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  OB_END
}
int ob_host_unregister(ob_ctx *ctx, void *ptr) {
  OB_BEGIN
  if(ptr)
    cudaHostUnregister(ptr);
  OB_END
}

int ob_matvec_partial(ob_ctx *ctx, int harmonic, const double *x, double *acc) {
  OB_BEGIN
  check_harmonic(harmonic);
  HarmonicState &H = ctx->hs[harmonic - 1];
  need(H.assembled, "matrix not assembled (call ob_assemble)");
  need(H.mode == 1 || H.mode == 3, "ob_matvec_partial: only the pair and rotated-axial forms are sharded by pairs");
  const int N = ctx->N(harmonic);
  upload(ctx, ctx->tmpA, x, N);
  const cplx *T = ctx->fac[harmonic == 1 ? 0 : 1].p;
  if(H.mode == 1) {
    launch_matvec_pairs(H.pplan, H.AB.p, ctx->tmpA.p, T, H.pplan.acc, 0, ctx->st);
    download(ctx, H.pplan.acc, acc, N);
  } else {
    launch_matvec_rot(H.rplan, H.rl, H.rot.p, H.rot_geo, ctx->tmpA.p, T, H.rplan.acc, 0, ctx->st);
    download(ctx, H.rplan.acc, acc, N);
  }
  OB_END
}

int ob_source_ff(ob_ctx *ctx, double *Q) {
  OB_BEGIN
  source_ff(ctx);
  download(ctx, ctx->Q.p, Q, ctx->N(1));
  OB_END
}

int ob_set_cg_tables(ob_ctx *ctx, const double *const tables[9]) {
  OB_BEGIN
  need(ctx->nobj > 0, "ob_set_cluster has not been called");
  const size_t sz = (size_t)flat_max(ctx->nMaxS) * flat_max(ctx->nMax) * flat_max(ctx->nMax);
  for(int t = 0; t < 9; ++t) {
    ctx->cg[t].alloc(sz);
    OB_CUDA(cudaMemcpyAsync(ctx->cg[t].p, tables[t], sz * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  }
  OB_CUDA(cudaStreamSynchronize(ctx->st));
  ctx->cg_nmax = ctx->nMax * 1000 + ctx->nMaxS;
  OB_END
}

int ob_build_cg_tables(ob_ctx *ctx) {
  OB_BEGIN
  need(ctx->nobj > 0, "ob_set_cluster has not been called");
  ctx->cg_nmax = -1;
  ensure_cg(ctx);
  OB_END
}

int ob_fetch_cg_table(ob_ctx *ctx, int t, double *out) {
  OB_BEGIN
  need(t >= 0 && t < 9 && ctx->cg_nmax >= 0, "tables not built");
  const size_t sz = (size_t)flat_max(ctx->nMaxS) * flat_max(ctx->nMax) * flat_max(ctx->nMax);
  OB_CUDA(cudaMemcpyAsync(out, ctx->cg[t].p, sz * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  OB_CUDA(cudaStreamSynchronize(ctx->st));
  OB_END
}

int ob_source_sh(ob_ctx *ctx, const double *Xint_conj, double *K, double *K1ana) {
  OB_BEGIN
  upload(ctx, ctx->tmpA, Xint_conj, ctx->N(1));
  source_sh(ctx, ctx->tmpA.p);
  download(ctx, ctx->Ksrc.p, K, ctx->N(2));
  download(ctx, ctx->K1ana.p, K1ana, ctx->N(2));
  OB_END
}

int ob_solve(ob_ctx *ctx, int harmonic, const double *rhs, double *x, const ob_gmres_opts *opts, int *iters,
             double *relres) {
  OB_BEGIN
  check_harmonic(harmonic);
  const int N = ctx->N(harmonic);
  DevBuf<cplx> &X = harmonic == 1 ? ctx->Xsca : ctx->XscaSH;
  X.alloc(N);
  const cplx *b;
  if(rhs) {
    upload(ctx, ctx->tmpB, rhs, N);
    b = ctx->tmpB.p;
  } else {
    DevBuf<cplx> &R = harmonic == 1 ? ctx->Q : ctx->Ksrc;
    need(R.p != nullptr, "no resident source vector for this harmonic");
    b = R.p;
  }
  GmresOut r = solve_dev(ctx, harmonic, b, X.p, opts);
  if(iters)
    *iters = r.iters;
  if(relres)
    *relres = r.relres;
  download(ctx, X.p, x, N);
  OB_END
}

int ob_dense_solve(ob_ctx *ctx, int N, const double *A, const double *b, double *x) {
  OB_BEGIN
  need(N > 0 && A && b && x, "ob_dense_solve: bad arguments");
  DevBuf<cplx> dA, db;
  dA.alloc((size_t)N * N);
  db.alloc(N);
  OB_CUDA(cudaMemcpyAsync(dA.p, A, sizeof(cplx) * (size_t)N * N, cudaMemcpyHostToDevice, ctx->st));
  OB_CUDA(cudaMemcpyAsync(db.p, b, sizeof(cplx) * (size_t)N, cudaMemcpyHostToDevice, ctx->st));
  const int info = lu_solve(dA.p, N, (size_t)N, ctx->lu, db.p, db.p, ctx->sm_count, ctx->st, ctx->launches);
  if(info != 0)
    throw Error("direct solve: the matrix is singular (zero pivot at column " + std::to_string(info) + ")");
  download(ctx, db.p, x, N);
  OB_END
}

int ob_unprecondition_ff(ob_ctx *ctx, const double *X_sca, double *X_int) {
  OB_BEGIN
  ensure_factors(ctx);
  const int N = ctx->N(1);
  upload(ctx, ctx->tmpA, X_sca, N);
  ctx->Xint.alloc(N);
  launch_hadamard(ctx->tmpA.p, ctx->fac[4].p, nullptr, ctx->Xint.p, N, 0, ctx->st);
  ctx->launches += 1;
  download(ctx, ctx->Xint.p, X_int, N);
  OB_END
}

int ob_unprecondition_sh(ob_ctx *ctx, const double *X_sca_SH, const double *K1ana, double *X_int_SH) {
  OB_BEGIN
  ensure_factors(ctx);
  const int N = ctx->N(2);
  upload(ctx, ctx->tmpA, X_sca_SH, N);
  upload(ctx, ctx->tmpB, K1ana, N);
  ctx->XintSH.alloc(N);
  launch_hadamard(ctx->tmpA.p, ctx->fac[5].p, ctx->tmpB.p, ctx->XintSH.p, N, 0, ctx->st);
  ctx->launches += 1;
  download(ctx, ctx->XintSH.p, X_int_SH, N);
  OB_END
}

int ob_cross_sections(ob_ctx *ctx, const double *X_sca, const double *X_int, const double *X_sca_SH,
                      const double *X_int_SH, int do_sh, double cs[5]) {
  OB_BEGIN
  need(ctx->have_inc && ctx->have_freq, "incident field / frequency not set");
  ensure_factors(ctx);
  upload(ctx, ctx->Xsca, X_sca, ctx->N(1));
  if(do_sh) {
    upload(ctx, ctx->Xint, X_int, ctx->N(1));
    upload(ctx, ctx->XscaSH, X_sca_SH, ctx->N(2));
    upload(ctx, ctx->XintSH, X_int_SH, ctx->N(2));
  }
  cross_sections(ctx, ctx->Xsca.p, ctx->Xint.p, ctx->XscaSH.p, ctx->XintSH.p, do_sh != 0, cs);
  OB_END
}

static void release_harmonic(ob_ctx *ctx, int harmonic) {
  HarmonicState &H = ctx->hs[harmonic - 1];
  H.S.release();
  // keep_matrices = 0 alternates the two harmonics in the same storage: park the pair buffer instead of paying a
  // cudaFree + cudaMalloc of ~100 GB per harmonic and step (the next assemble() picks it up when it is large enough)
  if(H.AB.p && !ctx->spare_AB.p) {
    std::swap(H.AB.p, ctx->spare_AB.p);
    std::swap(H.AB.n, ctx->spare_AB.n);
  }
  H.AB.release();
  H.aca.release();
  rot_records_drop(ctx, harmonic - 1);
  H.assembled = false;
}

int ob_fields(ob_ctx *ctx, long npts, const double *pts_sph, const double *X_sca, const double *X_int,
              const double *X_sca_SH, const double *X_int_SH, int do_sh, double *out, int *inner) {
  OB_BEGIN
  need(ctx->have_inc && ctx->have_freq, "incident field / frequency not set");
  need(npts >= 0 && (npts == 0 || (pts_sph && out)), "ob_fields: null buffers");
  const int N1 = ctx->N(1), N2 = ctx->N(2);
  // NULL vectors: the resident solution of the last ob_run
  if(X_sca)
    upload(ctx, ctx->Xsca, X_sca, N1);
  if(X_int)
    upload(ctx, ctx->Xint, X_int, N1);
  need(ctx->Xsca.p && ctx->Xint.p && ctx->Xsca.n >= (size_t)N1, "ob_fields: no FF solution (pass X_sca / X_int or call ob_run)");
  if(do_sh) {
    if(X_sca_SH)
      upload(ctx, ctx->XscaSH, X_sca_SH, N2);
    if(X_int_SH)
      upload(ctx, ctx->XintSH, X_int_SH, N2);
    need(ctx->XscaSH.p && ctx->XintSH.p && ctx->XscaSH.n >= (size_t)N2,
         "ob_fields: no SH solution (pass X_sca_SH / X_int_SH or call ob_run)");
    ensure_cg(ctx);
  }
  FieldInputs in;
  in.nobj = ctx->nobj;
  in.nMax = ctx->nMax;
  in.nMaxS = ctx->nMaxS;
  in.do_sh = do_sh ? 1 : 0;
  in.omega = ctx->omega;
  in.waveK = ctx->waveK;
  in.eps_b = ctx->eps_b;
  in.mu_b = ctx->mu_b;
  in.xyz = ctx->xyz.p;
  in.radius = ctx->radius.p;
  in.eps = ctx->mat[0].p;
  in.mu = ctx->mat[1].p;
  in.eps_SH = ctx->mat[2].p;
  in.mu_SH = ctx->mat[3].p;
  in.gamma = ctx->mat[6].p;
  in.ainc = ctx->ainc.p;
  in.Xsca = ctx->Xsca.p;
  in.Xint = ctx->Xint.p;
  in.XscaSH = ctx->XscaSH.p;
  in.XintSH = ctx->XintSH.p;
  for(int t = 0; t < 9; ++t)
    in.tab[t] = ctx->cg[t].p;
  DevBuf<double> dpts;
  DevBuf<cplx> dout;
  DevBuf<int> dinner;
  dpts.alloc(3 * (size_t)npts);
  dout.alloc(12 * (size_t)npts);
  dinner.alloc((size_t)npts);
  if(npts > 0) {
    OB_CUDA(cudaMemcpyAsync(dpts.p, pts_sph, 3 * (size_t)npts * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    launch_fields(in, npts, dpts.p, dout.p, dinner.p, ctx->st);
    ctx->launches += do_sh ? 2 : 1;
    OB_CUDA(cudaMemcpyAsync(out, dout.p, 12 * (size_t)npts * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->st));
    if(inner)
      OB_CUDA(cudaMemcpyAsync(inner, dinner.p, (size_t)npts * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    OB_CUDA(cudaStreamSynchronize(ctx->st));
  }
  OB_END
}

int ob_run(ob_ctx *ctx, const ob_gmres_opts *opts, int do_sh, double *X_sca, double *X_int, double *X_sca_SH,
           double *X_int_SH, double cs[5], int stats[2]) {
  OB_BEGIN
  need(ctx->have_inc && ctx->have_freq, "incident field / frequency not set");
  for(int i = 0; i < 16; ++i)
    ctx->tim[i] = 0;
  ctx->launches = 0;
  const int N1 = ctx->N(1), N2 = ctx->N(2);
  int it_ff = 0, it_sh = 0;
  // ---- update(): Q, S (PreconditionedMatrixSolver.h:82-100) ----
  {
    PhaseTimer t(ctx, 0);
    ctx->fac_valid = false;
    ensure_factors(ctx);
    source_ff(ctx);
    t.stop();
  }
  const bool direct = opts && opts->flavour == OB_SOLVE_DIRECT; // the LU assembles its own dense work matrix
  if(!ctx->keep_matrices) // one harmonic resident at a time: drop what an earlier step left behind
    release_harmonic(ctx, 2);
  if(!direct) {
    PhaseTimer t(ctx, 1);
    assemble(ctx, 1);
    t.stop();
  }
  // ---- solve(): FF ----
  {
    ctx->Xsca.alloc(N1);
    ctx->Xint.alloc(N1);
    double t0 = ctx->tim[7];
    PhaseTimer t(ctx, 2);
    GmresOut r = solve_dev(ctx, 1, ctx->Q.p, ctx->Xsca.p, opts);
    it_ff = r.iters;
    launch_hadamard(ctx->Xsca.p, ctx->fac[4].p, nullptr, ctx->Xint.p, N1, 0, ctx->st); // Solver.cpp:57-77
    ctx->launches += 1;
    t.stop();
    (void)t0;
  }
  if(do_sh) {
    if(!ctx->keep_matrices)
      release_harmonic(ctx, 1);
    {
      PhaseTimer t(ctx, 3);
      ctx->tmpA.alloc(std::max(N1, N2));
      // X_int_conj (PreconditionedMatrixSolver.h:66-67): a .* b conjugated
      launch_hadamard(ctx->Xsca.p, ctx->fac[4].p, nullptr, ctx->tmpA.p, N1, 1, ctx->st);
      ctx->launches += 1;
      source_sh(ctx, ctx->tmpA.p);
      t.stop();
    }
    if(!direct) {
      PhaseTimer t(ctx, 4);
      assemble(ctx, 2);
      t.stop();
    }
    {
      ctx->XscaSH.alloc(N2);
      ctx->XintSH.alloc(N2);
      PhaseTimer t(ctx, 5);
      GmresOut r = solve_dev(ctx, 2, ctx->Ksrc.p, ctx->XscaSH.p, opts);
      it_sh = r.iters;
      launch_hadamard(ctx->XscaSH.p, ctx->fac[5].p, ctx->K1ana.p, ctx->XintSH.p, N2, 0, ctx->st); // Solver.cpp:95-116
      ctx->launches += 1;
      t.stop();
    }
    if(!ctx->keep_matrices)
      release_harmonic(ctx, 2);
  }
  {
    PhaseTimer t(ctx, 6);
    cross_sections(ctx, ctx->Xsca.p, ctx->Xint.p, ctx->XscaSH.p, ctx->XintSH.p, do_sh != 0, cs);
    t.stop();
  }
  download(ctx, ctx->Xsca.p, X_sca, N1);
  download(ctx, ctx->Xint.p, X_int, N1);
  if(do_sh) {
    download(ctx, ctx->XscaSH.p, X_sca_SH, N2);
    download(ctx, ctx->XintSH.p, X_int_SH, N2);
  }
  if(stats) {
    stats[0] = it_ff;
    stats[1] = it_sh;
  }
  ctx->tim[9] = (double)ctx->launches;
  OB_END
}

int ob_timings(ob_ctx *ctx, double out[16]) {
  if(!ctx)
    return 1;
  for(int i = 0; i < 16; ++i)
    out[i] = ctx->tim[i];
  out[9] = (double)ctx->launches;
  return 0;
}

int ob_timer(ob_ctx *ctx, int op, double *ms) {
  OB_BEGIN
  if(op == 0)
    OB_CUDA(cudaEventRecord(ctx->evt0, ctx->st));
  else {
    OB_CUDA(cudaEventRecord(ctx->evt1, ctx->st));
    OB_CUDA(cudaEventSynchronize(ctx->evt1));
    float f = 0;
    OB_CUDA(cudaEventElapsedTime(&f, ctx->evt0, ctx->evt1));
    if(ms)
      *ms = f;
  }
  OB_END
}

int ob_measure_fp64_peak(ob_ctx *ctx, double *tflops) {
  OB_BEGIN
  *tflops = measure_fp64_peak(ctx->sm_count, ctx->st);
  OB_END
}

int ob_set_option(ob_ctx *ctx, const char *name, double value) {
  OB_BEGIN
  std::string n(name ? name : "");
  if(n == "matvec_variant") {
    ctx->matvec_variant = (int)value;
    for(int h = 0; h < 2; ++h) // only the dense form holds a streaming plan to rebuild; the others re-plan at assembly
      if(ctx->hs[h].assembled && ctx->hs[h].mode == 0 && ctx->hs[h].plan.M > 0 && ctx->hs[h].plan.N > 0)
        matvec_plan(ctx->hs[h].plan, ctx->hs[h].plan.M, ctx->hs[h].plan.N, ctx->hs[h].plan.ld, ctx->sm_count,
                    ctx->matvec_variant);
  } else if(n == "keep_matrices")
    ctx->keep_matrices = value != 0;
  else if(n == "trace_iterations")
    ctx->trace = value != 0;
  else if(n == "fused_arnoldi")
    ctx->fused_arnoldi = value != 0;
  else if(n == "speculate") // 1 (default): the next apply is enqueued ahead of the host's convergence test when safe
    ctx->speculate = value != 0;
  else if(n == "pairs_kb" || n == "pairs_groups") { // tuning: columns per pipeline stage / column groups (0 = auto)
    static int kb = 0, gr = 0;
    (n == "pairs_kb" ? kb : gr) = (int)value;
    pair_plan_tuning(kb, gr);
    for(int h = 0; h < 2; ++h) {
      ctx->hs[h].pplan_world = -1; // force a rebuild of the plan at the next assembly
      ctx->hs[h].assembled = false;
    }
  } else if(n == "operator") { // 0 = dense slab, 1 = compact pair form, 2 = ACA-compressed (reference's ACA path)
    need(value == 0 || value == 1 || value == 2 || value == 3,
         "operator must be 0 (dense), 1 (pairs), 2 (ACA) or 3 (rotated-axial)");
    ctx->operator_mode = (int)value;
    ctx->hs[0].assembled = ctx->hs[1].assembled = false;
  } else if(n == "assemble_minb") { // tuning: resident CTAs per SM k_assemble_pairs is compiled for (0 auto, 2, 3)
    need(value == 0 || value == 2 || value == 3, "assemble_minb must be 0, 2 or 3");
    assemble_pairs_tuning((int)value);
  } else if(n == "rot_share") { // 1 (default): one set of phases / small-d matrices serves both harmonics
    ctx->rot_share = value != 0;
    ctx->hs[0].assembled = ctx->hs[1].assembled = false;
  } else if(n == "rot_assembly") { // 1 = axial-only recursion (default), 0 = vtac_block at theta = 0 (cross-check path)
    need(value == 0 || value == 1, "rot_assembly must be 0 or 1");
    rot_tuning((int)value, -1, -1);
    ctx->hs[0].assembled = ctx->hs[1].assembled = false;
  } else if(n == "rot_rows" || n == "rot_ctas_per_sm") { // tuning of the rotated-axial plan (0 = auto); forces a new plan
    need(value >= 0 && value <= 64, "rot_rows / rot_ctas_per_sm out of range");
    if(n == "rot_rows")
      rot_tuning(-1, (int)value, -1);
    else
      rot_tuning(-1, -1, (int)value);
    for(int h = 0; h < 2; ++h) {
      ctx->hs[h].rplan_world = -1;
      ctx->hs[h].assembled = false;
    }
  } else if(n == "eps_aca") {
    need(value > 0, "eps_aca must be positive");
    ctx->eps_aca = value;
    ctx->hs[0].assembled = ctx->hs[1].assembled = false;
  } else if(n == "aca_budget_mb") {
    ctx->aca_budget = (size_t)std::max(1.0, value) << 20;
  }
  else if(n == "reset_timings") {
    for(int i = 0; i < 16; ++i)
      ctx->tim[i] = 0;
    ctx->launches = 0;
  } else
    throw Error("unknown option " + n);
  OB_END
}

} // extern "C"
