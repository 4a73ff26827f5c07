// Peak probes for the roofline denominators MEASURED_PEAKS.json does not carry: FP64 (a register-resident DFMA loop and a
// DMMA.8x8x4 loop) and the int8 tensor cores (back-to-back tcgen05.mma kind::i8 from resident shared memory).
// Used by bench.py / tools to report measured ceilings next to the nominal ones.
#include <cstdint>

#include "common.cuh"

namespace npgp {

__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double seed, double* out) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;  // keep the loop alive
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double seed, double* out) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = seed + threadIdx.x * 1e-3, b = seed - threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

// int8 tensor-core ceiling: every CTA (one per SM) issues `reps` x 8 MMAs of shape 128 x N x 32 (u8 x s8 -> s32) on
// operands that stay in shared memory, round-robin over the 512 / N accumulators that fit TMEM, one commit at the end.
// N = 256 is the densest single-CTA shape (12 KB of operands per 128 cycles of math: tensor bound); N = 64 is the digit
// engine's tile shape (6 KB per 32 cycles: shared-memory-feed bound), optionally with A-operand collector reuse.
__device__ __forceinline__ uint32_t pk_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t pk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
#define PK_MMA(NAME, COLL)                                                                                          \
  __device__ __forceinline__ void NAME(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {                      \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"                                                  \
                 "tcgen05.mma.cta_group::1.kind::i8" COLL " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),                  \
                 "l"(da), "l"(db), "r"(idesc)                                                                       \
                 : "memory");                                                                                       \
  }
PK_MMA(pk_mma, "")
PK_MMA(pk_mma_fill, ".collector::a::fill")
PK_MMA(pk_mma_use, ".collector::a::use")
PK_MMA(pk_mma_last, ".collector::a::lastuse")

template <int N, bool COLL>
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int reps) {
  extern __shared__ __align__(1024) uint8_t pk_sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  constexpr int NB = 8;                 // distinct B tiles (as the digit engine: several B digits per A digit)
  uint8_t* sA = pk_sm;                  // 128 x 32 bytes
  uint8_t* sB = pk_sm + 4096;           // NB x (N x 32 bytes)
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 4096 + NB * N * 32; e += 128) pk_sm[e] = (uint8_t)(e * 2654435761u >> 13);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pk_smem(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pk_smem(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = pk_desc(pk_smem(sA), 128 * 16, 128);
    uint64_t db[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) db[b] = pk_desc(pk_smem(sB + b * N * 32), N * 16, 128);
    constexpr int NACC = 512 / N;
    for (int r = 0; r < reps; ++r) {
      if (pred) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const uint32_t d = tmem + (uint32_t)((b % NACC) * N);
          if (!COLL) pk_mma(d, da, db[b], idesc);
          else if (b == 0) pk_mma_fill(d, da, db[b], idesc);
          else if (b == NB - 1) pk_mma_last(d, da, db[b], idesc);
          else pk_mma_use(d, da, db[b], idesc);
        }
      }
      __syncwarp();
    }
    if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pk_smem(&bar)) : "memory");
    __syncwarp();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                   : "=r"(done)
                   : "r"(pk_smem(&bar)), "r"(0u)
                   : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// The same measurement for a CTA PAIR (cta_group::2): one MMA covers 256 x N x 32 -- 128 rows of A from each CTA's shared
// memory, N/2 rows of B from each -- so per SM the B traffic is halved.  Both CTAs allocate TMEM, the leader issues,
// one multicast commit signals both.
#define PK_MMA2(NAME, COLL)                                                                                         \
  __device__ __forceinline__ void NAME(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {                      \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"                                                  \
                 "tcgen05.mma.cta_group::2.kind::i8" COLL " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),                  \
                 "l"(da), "l"(db), "r"(idesc)                                                                       \
                 : "memory");                                                                                       \
  }
PK_MMA2(pk2_mma, "")
PK_MMA2(pk2_mma_fill, ".collector::a::fill")
PK_MMA2(pk2_mma_use, ".collector::a::use")
PK_MMA2(pk2_mma_last, ".collector::a::lastuse")

__device__ __forceinline__ void pk_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int N, bool COLL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) i8_peak2_kernel(int reps) {
  extern __shared__ __align__(1024) uint8_t pk_sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  constexpr int NB = 8, NH = N / 2;     // this CTA holds half of the rows of every B tile
  uint8_t* sA = pk_sm;                  // 128 x 32 bytes (this CTA's half of the 256 rows)
  uint8_t* sB = pk_sm + 4096;           // NB x (N/2 x 32 bytes)
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int e = tid; e < 4096 + NB * NH * 32; e += 128) pk_sm[e] = (uint8_t)(e * 2654435761u >> 13);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pk_smem(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pk_smem(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  pk_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
    if (rank == 0) {
      const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint64_t da = pk_desc(pk_smem(sA), 128 * 16, 128);
      uint64_t db[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) db[b] = pk_desc(pk_smem(sB + b * NH * 32), NH * 16, 128);
      constexpr int NACC = 512 / N;
      for (int r = 0; r < reps; ++r) {
        if (pred) {
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const uint32_t d = tmem + (uint32_t)((b % NACC) * N);
            if (!COLL) pk2_mma(d, da, db[b], idesc);
            else if (b == 0) pk2_mma_fill(d, da, db[b], idesc);
            else if (b == NB - 1) pk2_mma_last(d, da, db[b], idesc);
            else pk2_mma_use(d, da, db[b], idesc);
          }
        }
        __syncwarp();
      }
      if (pred)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         pk_smem(&bar)),
                     "h"((uint16_t)3)
                     : "memory");
      __syncwarp();
    }
    uint32_t done = 0;
    long spins = 0;
    while (!done && ++spins < (1L << 26))  // bounded: a wrong guess about the pair protocol must not hang the GPU
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                   : "=r"(done)
                   : "r"(pk_smem(&bar)), "r"(0u)
                   : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  pk_cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace npgp

// int8 tensor-core probe: `blocks` CTAs x reps x 8 MMAs of 128 x n_tile x 32; int8 operations launched =
// blocks * reps * 8 * 2 * 128 * n_tile * 32.  n_tile: 256 (ceiling) or 64 (the digit engine's shape); collector: A reuse hints.
extern "C" int npgp_i8_peak_probe(int n_tile, int collector, int blocks, int reps, cudaStream_t stream) {
  if (blocks <= 0 || reps <= 0 || (n_tile != 64 && n_tile != 256)) return NPGP_EINVAL;
  const int smem = 4096 + 8 * n_tile * 32;
#define NPGP_PK(NT, CL)                                                                                              \
  do {                                                                                                               \
    NPGP_CUDA(cudaFuncSetAttribute(npgp::i8_peak_kernel<NT, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    npgp::i8_peak_kernel<NT, CL><<<blocks, 128, smem, stream>>>(reps);                                                \
  } while (0)
  if (n_tile == 256) { if (collector) NPGP_PK(256, true); else NPGP_PK(256, false); }
  else { if (collector) NPGP_PK(64, true); else NPGP_PK(64, false); }
#undef NPGP_PK
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

// mode 0: DFMA, mode 1: DMMA.  Launches `blocks` CTAs of 256 threads running `iters` iterations of 16 independent ops.
// FLOPs launched: mode 0: blocks*256*iters*16*2;  mode 1: blocks*8(warps)*iters*16*(8*8*4*2).
extern "C" int npgp_fp64_peak_probe(int mode, int blocks, int iters, double* out, cudaStream_t stream) {
  if (!out || blocks <= 0 || iters <= 0) return NPGP_EINVAL;
  if (mode == 0) npgp::dfma_peak_kernel<<<blocks, 256, 0, stream>>>(iters, 1.0, out);
  else npgp::dmma_peak_kernel<<<blocks, 256, 0, stream>>>(iters, 1.0, out);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

// CTA-pair variant of npgp_i8_peak_probe: blocks (even) CTAs in clusters of 2, every pair issues reps x 8 MMAs of
// 256 x n_tile x 32 (tcgen05.mma.cta_group::2); int8 operations = (blocks / 2) * reps * 8 * 2 * 256 * n_tile * 32.
extern "C" int npgp_i8_peak_probe_pair(int n_tile, int collector, int blocks, int reps, cudaStream_t stream) {
  if (blocks <= 0 || (blocks & 1) || reps <= 0 || (n_tile != 64 && n_tile != 256)) return NPGP_EINVAL;
  const int smem = 4096 + 8 * (n_tile / 2) * 32;
#define NPGP_PK2(NT, CL)                                                                                               \
  do {                                                                                                                \
    NPGP_CUDA(cudaFuncSetAttribute(npgp::i8_peak2_kernel<NT, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    npgp::i8_peak2_kernel<NT, CL><<<blocks, 128, smem, stream>>>(reps);                                                \
  } while (0)
  if (n_tile == 256) { if (collector) NPGP_PK2(256, true); else NPGP_PK2(256, false); }
  else { if (collector) NPGP_PK2(64, true); else NPGP_PK2(64, false); }
#undef NPGP_PK2
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
