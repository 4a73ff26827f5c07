// Exact-integer (Ozaki-split) evaluation of the row-quadratic contraction  T = K C,  q_i = sum_j T_ij K_ij  on the
// 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM), as an alternative to the FP64 DMMA GEMM of
// dgemm.cu (B200 has no FP64 tcgen05 kind and DMMA peaks at 37 TFLOP/s).
//
// Arithmetic.  Every row i of K and every row j of C (C is symmetric: row j = column j) is scaled by a power of two so
// that its entries lie in [-1, 1] and split into NS = 8 signed 7-bit slices, x = sum_p a_p 2^-(6+7p), a_p in [-64, 64]
// (exact: each step removes the leading 7 bits of the remainder).  Then
//     T_ij = 2^(e_i + f_j - 12) * sum_t 2^(-7t) G_t,   G_t = sum_{p+q=t} (a_p c_q^T)_ij   (int32, exact: |G_t| < 2^25)
// and only t = 0..7 is kept (36 int8 products): the dropped terms are below 2^-52 of (row max)(column max) K, i.e. of
// the order of the rounding error bound of the FP64 product itself.  The eight G_t live in eight TMEM accumulators of
// 128 lanes x 64 columns (all 512 columns), so one CTA owns an SM and a 128 x 64 output tile at a time.
//
// Data flow.  A slicing kernel writes the int8 slices of both operands already in the canonical K-major, no-swizzle
// core-matrix order of the UMMA shared-memory descriptor (8 rows x 16 bytes per core matrix), grouped so that everything
// one pipeline stage needs (32 bytes of K for all 8 slices of a 128-row block) is ONE contiguous range: the producer
// thread moves it with cp.async.bulk (global -> shared, mbarrier completion), no tensor map needed.  Warp roles:
// warp 0 producer, warp 1 MMA issuer (one thread, 36 MMAs per stage, tcgen05.commit releases the stage), warps 2-5
// epilogue (tcgen05.ld, fp64 recombination, scaling, fused row dot, 64-byte row segments of T).
// Descriptor encodings: cute/arch/mma_sm100_desc.hpp of the vendored CUTLASS; verified by tools/probes/umma_i8_probe.cu.
#include <cstdint>

#include "common.cuh"

namespace npgp {

constexpr int OZ_NS = 8;          // slices per operand
constexpr int OZ_BM = 128;        // rows of K per tile (TMEM lanes)
constexpr int OZ_BN = 64;         // columns per tile (TMEM columns per accumulator)
constexpr int OZ_KS = 32;         // bytes of K per pipeline stage = one MMA k-step
constexpr int OZ_STAGES = 4;
constexpr int OZ_A_STAGE = OZ_NS * OZ_BM * OZ_KS;   // 32 KB
constexpr int OZ_B_STAGE = OZ_NS * OZ_BN * OZ_KS;   // 16 KB
constexpr int OZ_THREADS = 192;

// ---------------------------------------------------------------------------------------------------------------------
// slicing: one warp per row of X (R x Kd, fp64, row stride ldx); rows >= R are written as zeros.
// out layout, BR = rows per block (128 for K, 64 for C):
//   byte offset = ((((r / BR) * nks + k / 32) * 8 + p) * 2 + (k % 32) / 16) * (BR * 16) + ((r % BR) / 8) * 128 + (r % 8) * 16 + k % 16
// expo[r]: x = X[r, :] * 2^-expo[r] in [-1, 1]
// ---------------------------------------------------------------------------------------------------------------------
template <int BR>
__global__ void __launch_bounds__(256) oz_slice_kernel(int R, int Rpad, int Kd, const double* __restrict__ X, long ldx,
                                                       int8_t* __restrict__ out, int* __restrict__ expo) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Rpad) return;
  const int r = warp, nks = Kd / OZ_KS;
  double mx = 0.0;
  if (r < R)
    for (int k = lane; k < Kd; k += 32) mx = fmax(mx, fabs(X[(long)r * ldx + k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e = 0;
  if (mx > 0.0) frexp(mx, &e);  // mx = m 2^e, m in [0.5, 1)  =>  |x| 2^-e < 1
  if (lane == 0) expo[r] = e;
  const long blk = (long)(r / BR) * nks;
  const int rin = ((r % BR) / 8) * 128 + (r % 8) * 16;
  for (int k = lane; k < Kd; k += 32) {
    double x = (r < R) ? ldexp(X[(long)r * ldx + k], 6 - e) : 0.0;  // x 2^6, in [-64, 64]
    int8_t* base = out + ((blk + k / OZ_KS) * OZ_NS * 2 + (k % OZ_KS) / 16) * (long)(BR * 16) + rin + (k % 16);
#pragma unroll
    for (int p = 0; p < OZ_NS; ++p) {
      const double a = rint(x);
      base[(long)p * 2 * (BR * 16)] = (int8_t)(int)a;
      x = (x - a) * 128.0;  // exact: removes the leading 7 bits of the remainder
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t oz_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void oz_mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem(b)), "r"(count));
}
__device__ __forceinline__ void oz_mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
        : "=r"(done)
        : "r"(oz_smem(b)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(oz_smem(b)) : "memory");
}
__device__ __forceinline__ void oz_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(oz_smem(dst)),
               "l"(src), "r"(bytes), "r"(oz_smem(b))
               : "memory");
}
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void oz_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  const uint32_t zero = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(zero), "r"(zero), "r"(zero), "r"(zero)
      : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem(b)) : "memory");
}
__device__ __forceinline__ void oz_tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

// ---------------------------------------------------------------------------------------------------------------------
// T (n x N) = K C,  q[i] += sum_j T_ij K_ij   from the sliced operands.  Persistent: grid = #SMs, tiles (rb, cb) in
// row-block-major order so that the CTAs working at the same time share their A row blocks through L2.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_rowquad_kernel(int n, int N, int Kd, const int8_t* __restrict__ As, const int* __restrict__ ea,
                  const int8_t* __restrict__ Bs, const int* __restrict__ eb, const double* __restrict__ Kmat, long ldk,
                  double* __restrict__ T, long ldt, double* __restrict__ q) {
  extern __shared__ __align__(1024) uint8_t oz_sm[];
  uint8_t* sA = oz_sm;                                   // OZ_STAGES x 32 KB
  uint8_t* sB = oz_sm + OZ_STAGES * OZ_A_STAGE;          // OZ_STAGES x 16 KB
  __shared__ __align__(8) uint64_t full[OZ_STAGES], empty[OZ_STAGES], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nks = Kd / OZ_KS, n_rb = (n + OZ_BM - 1) / OZ_BM, n_cb = N / OZ_BN;
  const int n_tiles = n_rb * n_cb;

  if (tid == 0) {
    for (int s = 0; s < OZ_STAGES; ++s) {
      oz_mbar_init(&full[s], 1);
      oz_mbar_init(&empty[s], 1);
    }
    oz_mbar_init(&acc_full, 1);
    oz_mbar_init(&acc_empty, 4);  // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int rb = tile / n_cb, cb = tile % n_cb;
        const int8_t* a = As + (long)rb * nks * OZ_A_STAGE;
        const int8_t* b = Bs + (long)cb * nks * OZ_B_STAGE;
        for (int ks = 0; ks < nks; ++ks) {
          oz_mbar_wait(&empty[stage], phase ^ 1);
          oz_mbar_expect_tx(&full[stage], OZ_A_STAGE + OZ_B_STAGE);
          oz_bulk_g2s(sA + stage * OZ_A_STAGE, a + (long)ks * OZ_A_STAGE, OZ_A_STAGE, &full[stage]);
          oz_bulk_g2s(sB + stage * OZ_B_STAGE, b + (long)ks * OZ_B_STAGE, OZ_B_STAGE, &full[stage]);
          if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // c_format S32 (2) @4, a/b format INT8 (1) @7/@10, both K-major, N>>3 @17, M>>4 @24
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        oz_mbar_wait(&acc_empty, acc_phase ^ 1);  // epilogue has drained the accumulators of the previous tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ks = 0; ks < nks; ++ks) {
          oz_mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = oz_smem(sA + stage * OZ_A_STAGE), b0 = oz_smem(sB + stage * OZ_B_STAGE);
#pragma unroll
          for (int p = 0; p < OZ_NS; ++p) {
            const uint64_t da = oz_desc(a0 + p * (2 * OZ_BM * 16), OZ_BM * 16, 128);
#pragma unroll
            for (int qq = 0; qq < OZ_NS; ++qq) {
              if (p + qq < OZ_NS) {
                const uint64_t db = oz_desc(b0 + qq * (2 * OZ_BN * 16), OZ_BN * 16, 128);
                oz_mma_i8(tmem + (uint32_t)((p + qq) * OZ_BN), da, db, idesc, (ks > 0 || p > 0) ? 1u : 0u);
              }
            }
          }
          oz_commit(&empty[stage]);  // frees the stage when these MMAs have read it
          if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
        }
        oz_commit(&acc_full);
        acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int rb = tile / n_cb, cb = tile % n_cb;
      const int row = rb * OZ_BM + quad * 32 + lane;
      oz_mbar_wait(&acc_full, acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int er = (row < n) ? ea[row] : 0;
      double qsum = 0.0;
      for (int c0 = 0; c0 < OZ_BN; c0 += 8) {
        uint32_t g[OZ_NS][8];
#pragma unroll
        for (int t = 0; t < OZ_NS; ++t)
          oz_tmem_ld8(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * OZ_BN + c0), g[t]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < n) {
          const int col = cb * OZ_BN + c0;
          double out[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            double acc = (double)(int)g[OZ_NS - 1][j];
#pragma unroll
            for (int t = OZ_NS - 2; t >= 0; --t) acc = fma(acc, 0.0078125, (double)(int)g[t][j]);
            out[j] = ldexp(acc, er + eb[col + j] - 12);
          }
          double* tp = T + (long)row * ldt + col;
#pragma unroll
          for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(tp + j) = make_double2(out[j], out[j + 1]);
          if (q) {
            const double* kp = Kmat + (long)row * ldk + col;
#pragma unroll
            for (int j = 0; j < 8; ++j) qsum = fma(out[j], kp[j], qsum);
          }
        }
      }
      if (q && row < n) atomicAdd(&q[row], qsum);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) oz_mbar_arrive(&acc_empty);
      acc_phase ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

constexpr int OZ_SMEM = OZ_STAGES * (OZ_A_STAGE + OZ_B_STAGE) + 1024;

}  // namespace npgp

using namespace npgp;

// workspace: A slices (ceil(n/128)*128 * Kd * 8 bytes) + C slices (N * Kd * 8) + exponents (int per padded row / column)
extern "C" long npgp_rowquad_i8_workspace_bytes(int n, int M) {
  const long npad = ((long)n + OZ_BM - 1) / OZ_BM * OZ_BM;
  return npad * M * OZ_NS + (long)M * M * OZ_NS + (npad + M) * (long)sizeof(int) + 1024;
}

// T (n x M) = K (n x M) @ C (M x M, symmetric);  q[i] += sum_j T_ij K_ij (q zeroed by the caller; NULL to skip).
// M must be a multiple of 64.  Replaces npgp_rowquad on the integer tensor-core path.
extern "C" int npgp_rowquad_i8(int n, int M, const double* K, long ldk, const double* C, long ldc, double* T, long ldt,
                               double* q, void* work, long work_bytes, cudaStream_t stream) {
  if (n < 0 || M < 0) return NPGP_EINVAL;
  if (n == 0 || M == 0) return NPGP_OK;
  if (!K || !C || !T || !work) return NPGP_EINVAL;
  if (M % OZ_BN || (ldt & 1) || (reinterpret_cast<uintptr_t>(T) & 15)) return NPGP_EUNSUPPORTED;
  if (work_bytes < npgp_rowquad_i8_workspace_bytes(n, M)) return NPGP_EWORKSPACE;
  const long npad = ((long)n + OZ_BM - 1) / OZ_BM * OZ_BM;
  int8_t* As = static_cast<int8_t*>(work);
  int8_t* Bs = As + npad * M * OZ_NS;
  int* ea = reinterpret_cast<int*>(Bs + (long)M * M * OZ_NS);
  int* eb = ea + npad;
  oz_slice_kernel<OZ_BM><<<(unsigned)((npad * 32 + 255) / 256), 256, 0, stream>>>(n, (int)npad, M, K, ldk, As, ea);
  NPGP_LAUNCH_CHECK();
  oz_slice_kernel<OZ_BN><<<(unsigned)(((long)M * 32 + 255) / 256), 256, 0, stream>>>(M, M, M, C, ldc, Bs, eb);
  NPGP_LAUNCH_CHECK();
  static bool attr_set = false;
  if (!attr_set) {
    NPGP_CUDA(cudaFuncSetAttribute(oz_rowquad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
    attr_set = true;
  }
  const int tiles = (int)(npad / OZ_BM) * (M / OZ_BN);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  oz_rowquad_kernel<<<grid, OZ_THREADS, OZ_SMEM, stream>>>(n, M, M, As, ea, Bs, eb, K, ldk, T, ldt, q);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
