// ob_matvec.cu -- K2: complex-FP64 block matvec y = S x for the GMRES operator.
//   reference: the Belos operator -> pzgemm_ (srcAna/scalapack/Belos.hpp:74-90, scalapack/Matrix.cpp:38-68)
//   and the in-tree matvec (srcAna/PreconditionedMatrix.cpp:1058-1085).
//
// Memory-bound (16 B per 8 flops): the kernel is a TMA streaming pipeline.
//   * S is the local row slab, column-major (ld = local rows): a row tile of TR rows is TR*16
//     contiguous bytes per column -> one cp.async.bulk (UBLKCP) per column, KB columns + the matching
//     x segment per pipeline stage, NS stages in flight per CTA, completion on mbarriers.
//   * one producer warp (one elected lane issues the copies), TR/RPT consumer threads; each consumer
//     owns RPT rows and accumulates over the columns with no cross-thread reduction.
//   * work unit = (row tile, column chunk); persistent CTAs take units round-robin; the pipeline
//     keeps streaming across unit boundaries.  Column chunks make the unit count a near multiple of
//     the SM count; partial sums go to a [chunks][M] buffer and are added in fixed chunk order by a
//     second tiny kernel (deterministic: no atomics).
#include "ob_internal.h"

namespace ob {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra WAIT_DONE;\n"
               "bra WAIT_LOOP;\n"
               "WAIT_DONE:\n"
               "}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// 1-D bulk async copy global -> shared, completion counted on an mbarrier (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}

template <int RPT, int KB, int NS> struct MvCfg {
  static constexpr int CONSUMERS = 128;
  static constexpr int TR = CONSUMERS * RPT;         // rows per tile
  static constexpr int THREADS = CONSUMERS + 32;     // + producer warp
  static constexpr size_t STAGE_BYTES = (size_t)KB * TR * sizeof(cplx) + (size_t)KB * sizeof(cplx);
  static constexpr size_t SMEM = NS * STAGE_BYTES + 2 * NS * sizeof(uint64_t) + 16;
};

template <int RPT, int KB, int NS>
__global__ void __launch_bounds__(MvCfg<RPT, KB, NS>::THREADS, 1)
k_matvec(const cplx *__restrict__ S, size_t ld, const cplx *__restrict__ x, cplx *__restrict__ out, int M, int N,
         int tiles, int chunks, int cols_per_chunk) {
  typedef MvCfg<RPT, KB, NS> C;
  extern __shared__ __align__(128) unsigned char smem[];
  cplx *stage_base = (cplx *)smem;
  uint64_t *full = (uint64_t *)(smem + NS * C::STAGE_BYTES);
  uint64_t *empty = full + NS;
  const int tid = threadIdx.x;
  if(tid == 0) {
    for(int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], C::CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int units = tiles * chunks;
  const size_t stage_elems = C::STAGE_BYTES / sizeof(cplx);

  if(tid >= C::CONSUMERS) {
    // ===== producer warp =====
    if(tid == C::CONSUMERS) {
      const uint64_t pol_s = policy_evict_first(), pol_x = policy_evict_last();
      int s = 0;
      uint32_t ph = 0;
      for(int u = blockIdx.x; u < units; u += gridDim.x) {
        const int tile = u / chunks, chunk = u - tile * chunks;
        const int r0 = tile * C::TR;
        const int rows = min(C::TR, M - r0);
        const int c0 = chunk * cols_per_chunk, c1 = min(N, c0 + cols_per_chunk);
        for(int c = c0; c < c1; c += KB) {
          const int nc = min(KB, c1 - c);
          mbar_wait(&empty[s], ph ^ 1);
          cplx *dst = stage_base + (size_t)s * stage_elems;
          mbar_expect_tx(&full[s], (uint32_t)((size_t)nc * rows * sizeof(cplx) + (size_t)nc * sizeof(cplx)));
          const cplx *src = S + (size_t)c * ld + r0;
#pragma unroll 4
          for(int q = 0; q < nc; ++q)
            bulk_g2s(dst + (size_t)q * C::TR, src + (size_t)q * ld, (uint32_t)(rows * sizeof(cplx)), &full[s], pol_s);
          bulk_g2s(dst + (size_t)KB * C::TR, x + c, (uint32_t)(nc * sizeof(cplx)), &full[s], pol_x);
          if(++s == NS) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else {
    // ===== consumers: thread t owns rows r0 + t + i*CONSUMERS, i < RPT =====
    int s = 0;
    uint32_t ph = 0;
    for(int u = blockIdx.x; u < units; u += gridDim.x) {
      const int tile = u / chunks, chunk = u - tile * chunks;
      const int r0 = tile * C::TR;
      const int c0 = chunk * cols_per_chunk, c1 = min(N, c0 + cols_per_chunk);
      double are[RPT], aim[RPT], bre[RPT], bim[RPT];
#pragma unroll
      for(int i = 0; i < RPT; ++i)
        are[i] = aim[i] = bre[i] = bim[i] = 0.0;
      for(int c = c0; c < c1; c += KB) {
        const int nc = min(KB, c1 - c);
        mbar_wait(&full[s], ph);
        const cplx *tile_s = stage_base + (size_t)s * stage_elems;
        const cplx *xs = tile_s + (size_t)KB * C::TR;
        if(nc == KB) {
#pragma unroll
          for(int q = 0; q < KB; q += 2) {
            const cplx x0 = xs[q], x1 = xs[q + 1];
#pragma unroll
            for(int i = 0; i < RPT; ++i) {
              const cplx a0 = tile_s[(size_t)q * C::TR + tid + i * C::CONSUMERS];
              const cplx a1 = tile_s[(size_t)(q + 1) * C::TR + tid + i * C::CONSUMERS];
              are[i] = fma(a0.x, x0.x, are[i]);
              aim[i] = fma(a0.x, x0.y, aim[i]);
              are[i] = fma(-a0.y, x0.y, are[i]);
              aim[i] = fma(a0.y, x0.x, aim[i]);
              bre[i] = fma(a1.x, x1.x, bre[i]);
              bim[i] = fma(a1.x, x1.y, bim[i]);
              bre[i] = fma(-a1.y, x1.y, bre[i]);
              bim[i] = fma(a1.y, x1.x, bim[i]);
            }
          }
        } else {
          for(int q = 0; q < nc; ++q) {
            const cplx x0 = xs[q];
#pragma unroll
            for(int i = 0; i < RPT; ++i) {
              const cplx a0 = tile_s[(size_t)q * C::TR + tid + i * C::CONSUMERS];
              are[i] = fma(a0.x, x0.x, are[i]);
              aim[i] = fma(a0.x, x0.y, aim[i]);
              are[i] = fma(-a0.y, x0.y, are[i]);
              aim[i] = fma(a0.y, x0.x, aim[i]);
            }
          }
        }
        __syncwarp();
        if((tid & 31) == 0)
          mbar_arrive(&empty[s]);
        if(++s == NS) {
          s = 0;
          ph ^= 1;
        }
      }
      cplx *o = out + (size_t)chunk * M + r0;
#pragma unroll
      for(int i = 0; i < RPT; ++i) {
        int r = tid + i * C::CONSUMERS;
        if(r0 + r < M)
          o[r] = mk(are[i] + bre[i], aim[i] + bim[i]);
      }
    }
  }
}

// y[r] = sum over chunks (fixed order) of partial[chunk][r]
__global__ void k_matvec_reduce(const cplx *__restrict__ partial, int M, int chunks, cplx *__restrict__ y) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if(r >= M)
    return;
  cplx s = partial[r];
  for(int c = 1; c < chunks; ++c)
    s = cadd(s, partial[(size_t)c * M + r]);
  y[r] = s;
}

// ---------------------------------------------------------------------------------------------
struct MvVariant {
  int RPT, KB, NS;
};
// variant 0 is the default; the others exist for on-device tuning (bench.py --matvec-variant)
static const MvVariant kVariants[] = {{1, 16, 6}, {2, 8, 6}, {1, 8, 8}, {2, 16, 3}, {1, 32, 3}, {1, 16, 4}};
static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

template <int RPT, int KB, int NS>
static void launch_variant(MatvecPlan const &p, const cplx *S, const cplx *x, cplx *out, cudaStream_t st) {
  typedef MvCfg<RPT, KB, NS> C;
  auto fn = k_matvec<RPT, KB, NS>;
  static bool attr_set[64] = {false};
  if(first_use_on_device(attr_set))
    OB_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  fn<<<p.grid, C::THREADS, C::SMEM, st>>>(S, p.ld, x, out, p.M, p.N, p.tiles, p.chunks, p.cols_per_chunk);
  OB_CUDA(cudaGetLastError());
}

void matvec_plan(MatvecPlan &p, int M, int N, size_t ld, int sm_count, int variant) {
  matvec_plan_release(p);
  if(variant < 0 || variant >= kNumVariants)
    throw Error("matvec: unknown kernel variant");
  const MvVariant v = kVariants[variant];
  const int TR = 128 * v.RPT;
  p.variant = variant;
  p.M = M;
  p.N = N;
  p.ld = ld;
  p.tiles = (M + TR - 1) / TR;
  // choose the column chunking: enough units to balance the SMs, each unit at least ~8 stages long
  const int kb_total = (N + v.KB - 1) / v.KB;
  int max_chunks = std::max(1, kb_total / 8);
  int best = 1;
  double best_eff = -1;
  for(int ch = 1; ch <= std::min(max_chunks, 64); ++ch) {
    long units = (long)p.tiles * ch;
    long rounds = (units + sm_count - 1) / sm_count;
    double eff = (double)units / (double)(rounds * sm_count);
    // mild preference for fewer chunks (less partial traffic)
    eff -= 0.002 * ch;
    if(eff > best_eff) {
      best_eff = eff;
      best = ch;
    }
  }
  p.chunks = best;
  int kb_per_chunk = (kb_total + p.chunks - 1) / p.chunks;
  p.cols_per_chunk = kb_per_chunk * v.KB;
  p.chunks = (N + p.cols_per_chunk - 1) / p.cols_per_chunk;
  p.grid = (int)std::min<long>((long)p.tiles * p.chunks, sm_count);
  if(p.chunks > 1)
    OB_CUDA(cudaMalloc(&p.partial, (size_t)p.chunks * M * sizeof(cplx)));
}
void matvec_plan_release(MatvecPlan &p) {
  if(p.partial)
    cudaFree(p.partial);
  p.partial = nullptr;
}
size_t matvec_launches_per_apply(MatvecPlan const &p) { return p.chunks > 1 ? 2 : 1; }

void launch_matvec(MatvecPlan const &p, const cplx *S, const cplx *x, cplx *y, cudaStream_t st, cudaEvent_t e0,
                   cudaEvent_t e1) {
  if(p.M == 0)
    return;
  if(e0)
    cudaEventRecord(e0, st);
  cplx *out = p.chunks > 1 ? p.partial : y;
  switch(p.variant) {
  case 0: launch_variant<1, 16, 6>(p, S, x, out, st); break;
  case 1: launch_variant<2, 8, 6>(p, S, x, out, st); break;
  case 2: launch_variant<1, 8, 8>(p, S, x, out, st); break;
  case 3: launch_variant<2, 16, 3>(p, S, x, out, st); break;
  case 4: launch_variant<1, 32, 3>(p, S, x, out, st); break;
  default: launch_variant<1, 16, 4>(p, S, x, out, st); break;
  }
  if(e1)
    cudaEventRecord(e1, st); // brackets the streaming kernel alone (the roofline denominator)
  if(p.chunks > 1) {
    k_matvec_reduce<<<(p.M + 255) / 256, 256, 0, st>>>(p.partial, p.M, p.chunks, y);
    OB_CUDA(cudaGetLastError());
  }
}

} // namespace ob
