// Exact-integer (Ozaki-split) evaluation of the two large contractions of the sparse-GP objectives,
//     row-quadratic:  T = K C,  q_i = sum_j T_ij K_ij            (npgp_rowquad_i8)
//     SYRK:           Out = alpha w0 K^T K                        (npgp_syrk_i8)
// on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM), as an alternative to the FP64 DMMA
// GEMM of dgemm.cu (B200 has no FP64 tcgen05 kind and DMMA peaks at 37 TFLOP/s).  Measured at B = 65536, M = 1024:
// 1.95 ms against 4.03 ms (row-quadratic) and 1.05 ms against 2.14 ms (SYRK), results equal to FP64 rounding.
//
// Arithmetic.  Every row i of K and every row j of C (C is symmetric: row j = column j) is scaled by a power of two so
// that its entries lie in [-1, 1] and split into NS = 8 signed 7-bit slices, x = sum_p a_p 2^-(6+7p), a_p in [-64, 64]
// (balanced base-128 digits of rint(x 2^55): exact up to 2^-56 of the row maximum).  Then
//     T_ij = 2^(e_i + f_j - 12) * sum_t 2^(-7t) G_t,   G_t = sum_{p+q=t} (a_p c_q^T)_ij   (int32, exact: |G_t| < 2^25)
// and only t = 0..7 is kept (36 int8 products): the dropped terms are below 2^-52 of (row max)(column max) K, i.e. of
// the order of the rounding error bound of the FP64 product itself.  The eight G_t live in eight TMEM accumulators of
// 128 lanes x 64 columns (all 512 columns), so one CTA owns an SM and a 128 x 64 output tile at a time.
//
// Data flow.  A slicing kernel writes the int8 slices of both operands already in the canonical K-major, no-swizzle
// core-matrix order of the UMMA shared-memory descriptor (8 rows x 16 bytes per core matrix), grouped so that everything
// one pipeline stage needs (32 bytes of K for all 8 slices of a 128-row block) is ONE contiguous range: the producer
// thread moves it with cp.async.bulk (global -> shared, mbarrier completion), no tensor map needed.  Warp roles:
// warp 0 producer, warp 1 MMA issuer (one elected lane, 36 MMAs per stage, tcgen05.commit releases the stage), warps 2-5
// epilogue (tcgen05.ld, integer recombination of the eight accumulators, power-of-two scaling, fused row dot against a K
// tile staged in shared memory while the MMAs run, 64-byte row segments of T).  The main loop runs at 50 cycles per
// 128x64x32 MMA, the shared-memory operand bandwidth (6 KB per MMA at 128 B/clk); the epilogue cannot overlap the next
// tile because the eight accumulators fill TMEM (11 % of the kernel).
// Descriptor encodings: cute/arch/mma_sm100_desc.hpp of the vendored CUTLASS; verified by tools/probes/umma_i8_probe.cu.
#include <cstdint>

#include "common.cuh"

namespace npgp {

constexpr int OZ_NS = 8;          // slices per operand
constexpr int OZ_BM = 128;        // rows of K per tile (TMEM lanes)
constexpr int OZ_BN = 64;         // columns per tile (TMEM columns per accumulator)
constexpr int OZ_KS = 32;         // bytes of K per pipeline stage = one MMA k-step
constexpr int OZ_STAGES = 2;       // row-quadratic kernel: leaves shared memory for a co-resident slicer CTA (3 measured equal)
constexpr int OZ_SYRK_STAGES = 4;  // SYRK kernel (no K tile to stage)
constexpr int OZ_KLD = OZ_BN + 2;  // leading dimension (doubles) of the staged K tile: conflict-free 16-byte row reads
constexpr int OZ_A_STAGE = OZ_NS * OZ_BM * OZ_KS;   // 32 KB
constexpr int OZ_B_STAGE = OZ_NS * OZ_BN * OZ_KS;   // 16 KB
constexpr int OZ_THREADS = 192;
#ifndef OZ_SLICE_MINB
#define OZ_SLICE_MINB 2  // two 512-thread slicer CTAs per SM (64 registers, 132 bytes of spill): the pass is HBM/latency bound
#endif

// ---------------------------------------------------------------------------------------------------------------------
// slicing of X (R x Kd, fp64, row stride ldx); rows >= R are written as zeros.
// out layout, BR = rows per block (128 for K, 64 for C):
//   byte offset = ((((r / BR) * nks + k / 32) * 8 + p) * 2 + (k % 32) / 16) * (BR * 16) + ((r % BR) / 8) * 128 + (r % 8) * 16 + k % 16
// expo[r]: x = X[r, :] * 2^-expo[r] in [-1, 1]
// ---------------------------------------------------------------------------------------------------------------------
// Slices of 16 consecutive K-dimension entries v[j] * sc (|v sc| < 2^55): y = rint(v sc) is written in balanced base-128
// digits, y = sum_p d_p 128^(7-p), d_p in [-64, 64], extracted as the unsigned digits of y + 64 (128^8 - 1)/127 (integer
// shifts and masks only; the FP64 pipe sees one multiply and one conversion per entry).  Digit p of the 16 entries is one
// 16-byte vector, stored at base + p * slice_stride.
__device__ __forceinline__ void oz_store_slices(const double (&v)[16], double sc, int8_t* base, long slice_stride) {
  constexpr unsigned long long BIAS = 64ull * ((1ull << 56) - 1ull) / 127ull;
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const unsigned long long u = (unsigned long long)(__double2ll_rn(v[j] * sc) + (long long)BIAS);
    lo[j] = (uint32_t)(u & 0xFFFFFFFull);  // digits 4..7
    hi[j] = (uint32_t)(u >> 28);           // digits 0..3 (the top one may reach 128)
  }
#pragma unroll
  for (int p = 0; p < OZ_NS; ++p) {
    uint32_t w[4];
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      uint32_t word = 0;
#pragma unroll
      for (int b4 = 0; b4 < 4; ++b4) {
        const int j = g4 * 4 + b4;
        const uint32_t src = (p < 4) ? hi[j] : lo[j];
        const int sh = 7 * (3 - (p & 3));
        const uint32_t dgt = (p == 0) ? (src >> sh) : ((src >> sh) & 127u);
        word |= ((dgt - 64u) & 0xffu) << (8 * b4);
      }
      w[g4] = word;
    }
    *reinterpret_cast<uint4*>(base + (long)p * slice_stride) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// Row-wise slicing (operand rows = rows of X, contraction along the columns).  CTA = 8 consecutive rows x 64 chunks of
// 16 columns (512 threads; longer rows loop): thread (row = t % 8, chunk = t / 8) keeps its 16 values in registers, the
// row maximum is reduced through shuffles + shared memory; the 8 lanes of a row group write 128 contiguous bytes (one
// core matrix) per slice.
template <int BR>
__global__ void __launch_bounds__(512, OZ_SLICE_MINB) oz_slice_kernel(int R, int Rpad, int Kd, const double* __restrict__ X, long ldx,
                                                       int8_t* __restrict__ out, int* __restrict__ expo) {
  __shared__ double smax[16][8];
  __shared__ int sexp[8];
  const int t = threadIdx.x, rr = t & 7, ch = t >> 3, warp = t >> 5;
  const int r = blockIdx.x * 8 + rr;
  const int nks = Kd / OZ_KS, nch = Kd / 16;
  double v[16];
  double mx = 0.0;
  for (int c = ch; c < nch; c += 64) {
    if (r < R) {
      const double2* src = reinterpret_cast<const double2*>(X + (long)r * ldx + c * 16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double2 d = src[j];
        if (c == ch) { v[2 * j] = d.x; v[2 * j + 1] = d.y; }
        mx = fmax(mx, fmax(fabs(d.x), fabs(d.y)));
      }
    } else if (c == ch) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.0;
    }
  }
  mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
  mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
  if ((t & 31) < 8) smax[warp][rr] = mx;
  __syncthreads();
  if (t < 8) {
    double m = 0.0;
#pragma unroll
    for (int w = 0; w < 16; ++w) m = fmax(m, smax[w][t]);
    int e = 0;
    if (m > 0.0) frexp(m, &e);  // m = f 2^e, f in [0.5, 1)  =>  |x| 2^-e < 1
    sexp[t] = e;
    if (blockIdx.x * 8 + t < Rpad) expo[blockIdx.x * 8 + t] = e;
  }
  __syncthreads();
  if (r >= Rpad) return;
  const double sc = __hiloint2double((1023 + 55 - sexp[rr]) << 20, 0);  // 2^(55 - e)
  const long blk = (long)(r / BR) * nks;
  const int rin = ((r % BR) / 8) * 128 + (r % 8) * 16;
  for (int c = ch; c < nch; c += 64) {
    if (c != ch && r < R) {
      const double2* src = reinterpret_cast<const double2*>(X + (long)r * ldx + c * 16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double2 d = src[j];
        v[2 * j] = d.x;
        v[2 * j + 1] = d.y;
      }
    }
    oz_store_slices(v, sc, out + ((blk + c / 2) * OZ_NS * 2 + (c & 1)) * (long)(BR * 16) + rin, 2L * (BR * 16));
  }
}

// Column maxima of |X| (R x Kd) as int64 bit patterns (non-negative doubles order like integers); cmax zeroed by the caller.
// Optionally the same pass also accumulates the weighted column sums wsum[j] += sum_i w_i X_ij (K^T g_mu of the SVGP step,
// which would otherwise be one more read of X).
__global__ void __launch_bounds__(256) oz_colmax_kernel(int R, int Kd, const double* __restrict__ X, long ldx,
                                                        int rows_per_cta, unsigned long long* __restrict__ cmax,
                                                        const double* __restrict__ w, double* __restrict__ wsum) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= Kd) return;
  const int i0 = blockIdx.y * rows_per_cta, i1 = min(R, i0 + rows_per_cta);
  double m = 0.0, acc = 0.0;
  if (w) {
    for (int i = i0; i < i1; ++i) {
      const double v = X[(long)i * ldx + j];
      m = fmax(m, fabs(v));
      acc = fma(w[i], v, acc);
    }
    atomicAdd(&wsum[j], acc);
  } else {
    for (int i = i0; i < i1; ++i) m = fmax(m, fabs(X[(long)i * ldx + j]));
  }
  atomicMax(&cmax[j], (unsigned long long)__double_as_longlong(m));
}

__global__ void oz_exp_from_max_kernel(int n, const unsigned long long* __restrict__ cmax, int* __restrict__ expo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double m = __longlong_as_double((long long)cmax[j]);
  int e = 0;
  if (m > 0.0) frexp(m, &e);
  expo[j] = e;
}

// Transposed slicing for the SYRK: operand rows = columns j of X (R x Kd), contraction along the rows i.  CTA = one
// k-stage (32 rows i) x 128 columns j: coalesced reads along j, transpose through shared memory, thread (j, half) slices
// the 16 consecutive i of its half stage.  Layout as oz_slice_kernel<128> with (row, k) = (j, i).
__global__ void __launch_bounds__(256) oz_slice_t_kernel(int R, int Kd, const double* __restrict__ X, long ldx,
                                                         const int* __restrict__ expo, int8_t* __restrict__ out) {
  __shared__ double tile[32][129];
  const int ks = blockIdx.x, jb = blockIdx.y, nks = gridDim.x;
  for (int e = threadIdx.x; e < 32 * 128; e += 256) {
    const int i = e >> 7, j = e & 127;
    const int gi = ks * 32 + i, gj = jb * 128 + j;
    tile[i][j] = (gi < R && gj < Kd) ? X[(long)gi * ldx + gj] : 0.0;
  }
  __syncthreads();
  const int j = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int gj = jb * 128 + j;
  double v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = tile[half * 16 + k][j];
  const double sc = __hiloint2double((1023 + 55 - ((gj < Kd) ? expo[gj] : 0)) << 20, 0);
  int8_t* base = out + (((long)jb * nks + ks) * OZ_NS * 2 + half) * (long)(OZ_BM * 16) + (j / 8) * 128 + (j % 8) * 16;
  oz_store_slices(v, sc, base, 2L * (OZ_BM * 16));
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
// one elected lane of a fully active warp (the form the compiler recognises: a tcgen05 instruction under this predicate is
// issued once, without the elect-and-retry loop it wraps around the same instruction in a `lane == 0` branch)
__device__ __forceinline__ bool oz_elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
// shared-memory descriptors of one stage differ only in the 14-bit start-address field of the low word
__device__ __forceinline__ uint64_t oz_desc_hi_lo(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
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
                  double* __restrict__ T, long ldt, double* __restrict__ q, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t oz_sm[];
  uint8_t* sA = oz_sm;                                   // OZ_STAGES x 32 KB
  uint8_t* sB = oz_sm + OZ_STAGES * OZ_A_STAGE;          // OZ_STAGES x 16 KB
  double* sK = reinterpret_cast<double*>(oz_sm + OZ_STAGES * (OZ_A_STAGE + OZ_B_STAGE));  // 128 x OZ_KLD tile of K
  __shared__ double scol[OZ_BN];                         // 2^f_j of the tile's columns
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
      long long w_empty = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int rb = tile / n_cb, cb = tile % n_cb;
        const int8_t* a = As + (long)rb * nks * OZ_A_STAGE;
        const int8_t* b = Bs + (long)cb * nks * OZ_B_STAGE;
        for (int ks = 0; ks < nks; ++ks) {
          const long long c0 = dbg ? clock64() : 0;
          oz_mbar_wait(&empty[stage], phase ^ 1);
          if (dbg) w_empty += clock64() - c0;
          oz_mbar_expect_tx(&full[stage], OZ_A_STAGE + OZ_B_STAGE);
          oz_bulk_g2s(sA + stage * OZ_A_STAGE, a + (long)ks * OZ_A_STAGE, OZ_A_STAGE, &full[stage]);
          oz_bulk_g2s(sB + stage * OZ_B_STAGE, b + (long)ks * OZ_B_STAGE, OZ_B_STAGE, &full[stage]);
          if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg && blockIdx.x == 0) dbg[0] = w_empty;
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop (warp-uniform control flow), one elected lane issues =====
    {
      const bool leader = oz_elect_one();
      // c_format S32 (2) @4, a/b format INT8 (1) @7/@10, both K-major, N>>3 @17, M>>4 @24
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
      constexpr uint32_t kHiDesc = (128u >> 4) | (1u << 14);  // stride byte offset 128, descriptor version 1
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      long long w_acc = 0, w_full = 0, t_all = dbg ? clock64() : 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        long long c0 = dbg ? clock64() : 0;
        oz_mbar_wait(&acc_empty, acc_phase ^ 1);  // epilogue has drained the accumulators of the previous tile
        if (dbg) w_acc += clock64() - c0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ks = 0; ks < nks; ++ks) {
          c0 = dbg ? clock64() : 0;
          oz_mbar_wait(&full[stage], phase);
          if (dbg) w_full += clock64() - c0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (leader) {
            const uint32_t a_lo = ((oz_smem(sA + stage * OZ_A_STAGE) >> 4) & 0x3FFFu) | ((uint32_t)(OZ_BM * 16 >> 4) << 16);
            const uint32_t b_lo = ((oz_smem(sB + stage * OZ_B_STAGE) >> 4) & 0x3FFFu) | ((uint32_t)(OZ_BN * 16 >> 4) << 16);
            const uint32_t first = ks > 0 ? 1u : 0u;
#pragma unroll
            for (int p = 0; p < OZ_NS; ++p) {
              const uint64_t da = oz_desc_hi_lo(a_lo + p * (2 * OZ_BM * 16 >> 4), kHiDesc);
#pragma unroll
              for (int qq = 0; qq < OZ_NS; ++qq) {
                if (p + qq < OZ_NS) {
                  const uint64_t db = oz_desc_hi_lo(b_lo + qq * (2 * OZ_BN * 16 >> 4), kHiDesc);
                  oz_mma_i8(tmem + (uint32_t)((p + qq) * OZ_BN), da, db, idesc, p > 0 ? 1u : first);
                }
              }
            }
            oz_commit(&empty[stage]);  // frees the stage when these MMAs have read it
          }
          __syncwarp();
          if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
        }
        if (leader) oz_commit(&acc_full);
        __syncwarp();
        acc_phase ^= 1;
      }
      if (dbg && blockIdx.x == 0 && leader) {
        dbg[1] = w_acc;
        dbg[2] = w_full;
        dbg[3] = clock64() - t_all;
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3, et = tid - 64;  // et: 0..127
    uint32_t acc_phase = 0;
    long long w_accfull = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int rb = tile / n_cb, cb = tile % n_cb;
      const int row = rb * OZ_BM + quad * 32 + lane;
      // while the MMAs of this tile run: stage the K tile (row dot) and the column scales
      if (q) {
        for (int rr = et >> 5; rr < OZ_BM; rr += 4) {  // one row (512 contiguous bytes) per warp instruction
          const int gr = rb * OZ_BM + rr;
          double2 d = make_double2(0.0, 0.0);
          if (gr < n) d = *reinterpret_cast<const double2*>(Kmat + (long)gr * ldk + cb * OZ_BN + 2 * lane);
          *reinterpret_cast<double2*>(sK + rr * OZ_KLD + 2 * lane) = d;
        }
      }
      if (et < OZ_BN) scol[et] = __hiloint2double((1023 + eb[cb * OZ_BN + et]) << 20, 0);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const long long c0 = dbg ? clock64() : 0;
      oz_mbar_wait(&acc_full, acc_phase);
      if (dbg) w_accfull += clock64() - c0;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int er = (row < n) ? ea[row] : 0;
      const double rs = __hiloint2double((1023 + er - 33) << 20, 0);  // 2^(e_i - 12 - 21)
      const double* krow = sK + (quad * 32 + lane) * OZ_KLD;
      double qsum = 0.0;
      for (int c0i = 0; c0i < OZ_BN; c0i += 8) {
        uint32_t g[OZ_NS][8];
#pragma unroll
        for (int t = 0; t < OZ_NS; ++t)
          oz_tmem_ld8(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * OZ_BN + c0i), g[t]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        double out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // sum_t 2^(-7t) G_t = 2^-21 hi + 2^-49 lo with hi, lo exact 64-bit integers (< 2^47)
          long long hi = (int)g[0][j], lo = (int)g[4][j];
#pragma unroll
          for (int t = 1; t < 4; ++t) {
            hi = hi * 128 + (int)g[t][j];
            lo = lo * 128 + (int)g[4 + t][j];
          }
          const double val = fma((double)lo, 3.7252902984619140625e-9 /* 2^-28 */, (double)hi);
          out[j] = val * rs * scol[c0i + j];
        }
        if (row < n) {
          double* tp = T + (long)row * ldt + cb * OZ_BN + c0i;
#pragma unroll
          for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(tp + j) = make_double2(out[j], out[j + 1]);
          if (q) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const double2 kk = *reinterpret_cast<const double2*>(krow + c0i + j);
              qsum = fma(out[j], kk.x, qsum);
              qsum = fma(out[j + 1], kk.y, qsum);
            }
          }
        }
      }
      if (q && row < n) atomicAdd(&q[row], qsum);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) oz_mbar_arrive(&acc_empty);
      acc_phase ^= 1;
      asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone is done with sK / scol before the next tile overwrites them
    }
    if (dbg && blockIdx.x == 0 && warp == 2 && lane == 0) dbg[4] = w_accfull;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------------------------------------------------
// Out (M x M, upper 128 x 64 tiles) += alpha * w0 * X^T X from the transposed slices (operand rows = columns of X).  The
// contraction runs over the R rows of X; it is cut into chunks of `stages_per_chunk` k-stages (<= 256 = 8192 rows, so that
// |G_t| < 8 * 8192 * 2^12 = 2^28 stays exact in int32) and every (tile, chunk) work item adds its FP64 result atomically.
// Both operands come from the same 128-row-block layout: the B operand (64 columns) is one half of a block, fetched as 16
// pieces of 1 KB per stage.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_syrk_kernel(int M, int nks_total, int stages_per_chunk, const int8_t* __restrict__ Xs, const int* __restrict__ ex,
               double alpha, const double* __restrict__ w0, const double* __restrict__ uniform_count,
               double uniform_target, double* __restrict__ Out, long ldo) {
  if (uniform_count && *uniform_count != uniform_target) return;  // unequal weights: the FP64 weighted kernel handles it
  extern __shared__ __align__(1024) uint8_t oz_sm[];
  uint8_t* sA = oz_sm;
  uint8_t* sB = oz_sm + OZ_SYRK_STAGES * OZ_A_STAGE;
  __shared__ __align__(8) uint64_t full[OZ_SYRK_STAGES], empty[OZ_SYRK_STAGES], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ double scol[OZ_BN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rb = M / OZ_BM, n_cb = M / OZ_BN;
  const int n_tiles = n_rb * n_cb - n_rb * (n_rb - 1);  // sum_rb (n_cb - 2 rb): tiles with cb >= 2 rb
  const int n_chunks = (nks_total + stages_per_chunk - 1) / stages_per_chunk;
  const int n_items = n_tiles * n_chunks;
  auto decode = [&](int item, int& rb, int& cb, int& k0, int& k1) {
    int tl = item / n_chunks;
    const int chunk = item % n_chunks;
    rb = 0;
    while (tl >= n_cb - 2 * rb) {
      tl -= n_cb - 2 * rb;
      ++rb;
    }
    cb = 2 * rb + tl;
    k0 = chunk * stages_per_chunk;
    k1 = min(nks_total, k0 + stages_per_chunk);
  };

  if (tid == 0) {
    for (int s = 0; s < OZ_SYRK_STAGES; ++s) {
      oz_mbar_init(&full[s], 1);
      oz_mbar_init(&empty[s], 1);
    }
    oz_mbar_init(&acc_full, 1);
    oz_mbar_init(&acc_empty, 4);
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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int rb, cb, k0, k1;
        decode(item, rb, cb, k0, k1);
        const int8_t* a = Xs + (long)rb * nks_total * OZ_A_STAGE;
        const int8_t* b = Xs + (long)(cb >> 1) * nks_total * OZ_A_STAGE + (cb & 1) * (OZ_BN * 16);
        for (int ks = k0; ks < k1; ++ks) {
          oz_mbar_wait(&empty[stage], phase ^ 1);
          oz_mbar_expect_tx(&full[stage], OZ_A_STAGE + OZ_B_STAGE);
          oz_bulk_g2s(sA + stage * OZ_A_STAGE, a + (long)ks * OZ_A_STAGE, OZ_A_STAGE, &full[stage]);
#pragma unroll
          for (int pp = 0; pp < 2 * OZ_NS; ++pp)  // (slice, plane) pieces: 64 rows x 16 bytes each
            oz_bulk_g2s(sB + stage * OZ_B_STAGE + pp * (OZ_BN * 16), b + (long)ks * OZ_A_STAGE + pp * (OZ_BM * 16),
                        OZ_BN * 16, &full[stage]);
          if (++stage == OZ_SYRK_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = oz_elect_one();
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
      constexpr uint32_t kHiDesc = (128u >> 4) | (1u << 14);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int rb, cb, k0, k1;
        decode(item, rb, cb, k0, k1);
        oz_mbar_wait(&acc_empty, acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ks = k0; ks < k1; ++ks) {
          oz_mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (leader) {
            const uint32_t a_lo = ((oz_smem(sA + stage * OZ_A_STAGE) >> 4) & 0x3FFFu) | ((uint32_t)(OZ_BM * 16 >> 4) << 16);
            const uint32_t b_lo = ((oz_smem(sB + stage * OZ_B_STAGE) >> 4) & 0x3FFFu) | ((uint32_t)(OZ_BN * 16 >> 4) << 16);
            const uint32_t first = ks > k0 ? 1u : 0u;
#pragma unroll
            for (int p = 0; p < OZ_NS; ++p) {
              const uint64_t da = oz_desc_hi_lo(a_lo + p * (2 * OZ_BM * 16 >> 4), kHiDesc);
#pragma unroll
              for (int qq = 0; qq < OZ_NS; ++qq) {
                if (p + qq < OZ_NS) {
                  const uint64_t db = oz_desc_hi_lo(b_lo + qq * (2 * OZ_BN * 16 >> 4), kHiDesc);
                  oz_mma_i8(tmem + (uint32_t)((p + qq) * OZ_BN), da, db, idesc, p > 0 ? 1u : first);
                }
              }
            }
            oz_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == OZ_SYRK_STAGES) { stage = 0; phase ^= 1; }
        }
        if (leader) oz_commit(&acc_full);
        __syncwarp();
        acc_phase ^= 1;
      }
    }
  } else {
    const int quad = warp & 3, et = tid - 64;
    uint32_t acc_phase = 0;
    const double aw = alpha * (w0 ? w0[0] : 1.0);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int rb, cb, k0, k1;
      decode(item, rb, cb, k0, k1);
      const int row = rb * OZ_BM + quad * 32 + lane;
      if (et < OZ_BN) scol[et] = __hiloint2double((1023 + ex[cb * OZ_BN + et]) << 20, 0);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      oz_mbar_wait(&acc_full, acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const double rs = aw * __hiloint2double((1023 + ex[row] - 33) << 20, 0);
      for (int c0i = 0; c0i < OZ_BN; c0i += 8) {
        uint32_t g[OZ_NS][8];
#pragma unroll
        for (int t = 0; t < OZ_NS; ++t)
          oz_tmem_ld8(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * OZ_BN + c0i), g[t]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        double* op = Out + (long)row * ldo + cb * OZ_BN + c0i;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          long long hi = (int)g[0][j], lo = (int)g[4][j];
#pragma unroll
          for (int t = 1; t < 4; ++t) {
            hi = hi * 128 + (int)g[t][j];
            lo = lo * 128 + (int)g[4 + t][j];
          }
          const double val = fma((double)lo, 3.7252902984619140625e-9 /* 2^-28 */, (double)hi);
          atomicAdd(op + j, val * rs * scol[c0i + j]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) oz_mbar_arrive(&acc_empty);
      acc_phase ^= 1;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

__global__ void oz_symmetrize_upper_kernel(int M, double* C, long ldc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r < M && c < M && c < r) C[(long)r * ldc + c] = C[(long)c * ldc + r];
}

constexpr int OZ_SMEM = OZ_STAGES * (OZ_A_STAGE + OZ_B_STAGE) + OZ_BM * OZ_KLD * 8 + 1024;

}  // namespace npgp

using namespace npgp;

static long long* g_oz_dbg = nullptr;
// debugging aid: device buffer of 8 counters filled by CTA 0 (cycles the producer / MMA issuer / epilogue spend waiting)
extern "C" void npgp_rowquad_i8_debug(long long* dev_counters) { g_oz_dbg = dev_counters; }

// workspace: A slices (ceil(n/128)*128 * Kd * 8 bytes) + C slices (N * Kd * 8) + exponents (int per padded row / column)
extern "C" long npgp_rowquad_i8_workspace_bytes(int n, int M) {
  const long npad = ((long)n + OZ_BM - 1) / OZ_BM * OZ_BM;
  return npad * M * OZ_NS + (long)M * M * OZ_NS + (npad + M) * (long)sizeof(int) + 1024;
}

// T (n x M) = K (n x M) @ C (M x M, symmetric);  q[i] += sum_j T_ij K_ij (q zeroed by the caller; NULL to skip).
// M must be a multiple of 64.  Replaces npgp_rowquad on the integer tensor-core path.
static int rowquad_i8_impl(int n, int M, const double* K, long ldk, const double* C, long ldc, double* T, long ldt,
                           double* q, void* work, long work_bytes, cudaStream_t stream, bool slice, bool gemm = true);

// The slicing passes alone (K and C into `work`); follow with npgp_rowquad_i8_gemm_only.  Lets the caller put other
// HBM-bound work on a second stream exactly under the tensor-core kernel.
extern "C" int npgp_rowquad_i8_slice_only(int n, int M, const double* K, long ldk, const double* C, long ldc, void* work,
                                          long work_bytes, cudaStream_t stream) {
  if (!K || !C || !work) return (n == 0 || M == 0) ? NPGP_OK : NPGP_EINVAL;
  return rowquad_i8_impl(n, M, K, ldk, C, ldc, const_cast<double*>(K), 2, nullptr, work, work_bytes, stream, true, false);
}

extern "C" int npgp_rowquad_i8(int n, int M, const double* K, long ldk, const double* C, long ldc, double* T, long ldt,
                               double* q, void* work, long work_bytes, cudaStream_t stream) {
  return rowquad_i8_impl(n, M, K, ldk, C, ldc, T, ldt, q, work, work_bytes, stream, true);
}

// The tensor-core kernel alone, on the slices a previous npgp_rowquad_i8 call with the same shapes left in `work`
// (measurement helper: times the GEMM without the slicing passes).
extern "C" int npgp_rowquad_i8_gemm_only(int n, int M, const double* K, long ldk, double* T, long ldt, double* q,
                                         void* work, long work_bytes, cudaStream_t stream) {
  return rowquad_i8_impl(n, M, K, ldk, K, 0, T, ldt, q, work, work_bytes, stream, false);
}

static int rowquad_i8_impl(int n, int M, const double* K, long ldk, const double* C, long ldc, double* T, long ldt,
                           double* q, void* work, long work_bytes, cudaStream_t stream, bool slice, bool gemm) {
  if (n < 0 || M < 0) return NPGP_EINVAL;
  if (n == 0 || M == 0) return NPGP_OK;
  if (!K || !C || !T || !work) return NPGP_EINVAL;
  // 16-byte vector accesses: T rows, the K tile of the row dot and the slicers' reads of K and C
  if (M % OZ_BN || (ldt & 1) || (reinterpret_cast<uintptr_t>(T) & 15) || (ldk & 1) ||
      (reinterpret_cast<uintptr_t>(K) & 15) || (slice && ((ldc & 1) || (reinterpret_cast<uintptr_t>(C) & 15))))
    return NPGP_EUNSUPPORTED;
  if (work_bytes < npgp_rowquad_i8_workspace_bytes(n, M)) return NPGP_EWORKSPACE;
  const long npad = ((long)n + OZ_BM - 1) / OZ_BM * OZ_BM;
  int8_t* As = static_cast<int8_t*>(work);
  int8_t* Bs = As + npad * M * OZ_NS;
  int* ea = reinterpret_cast<int*>(Bs + (long)M * M * OZ_NS);
  int* eb = ea + npad;
  if (slice) {
    oz_slice_kernel<OZ_BM><<<(unsigned)(npad / 8), 512, 0, stream>>>(n, (int)npad, M, K, ldk, As, ea);
    NPGP_LAUNCH_CHECK();
    oz_slice_kernel<OZ_BN><<<(unsigned)(M / 8), 512, 0, stream>>>(M, M, M, C, ldc, Bs, eb);
    NPGP_LAUNCH_CHECK();
  }
  if (!gemm) return NPGP_OK;
  static bool attr_set = false;
  if (!attr_set) {
    NPGP_CUDA(cudaFuncSetAttribute(oz_rowquad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
    attr_set = true;
  }
  const int tiles = (int)(npad / OZ_BM) * (M / OZ_BN);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  oz_rowquad_kernel<<<grid, OZ_THREADS, OZ_SMEM, stream>>>(n, M, M, As, ea, Bs, eb, K, ldk, T, ldt, q, g_oz_dbg);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

// workspace: transposed slices (M * ceil(n/32)*32 * 8 bytes) + column maxima (8 bytes each) + exponents
extern "C" long npgp_syrk_i8_workspace_bytes(int n, int M) {
  const long npad = ((long)n + OZ_KS - 1) / OZ_KS * OZ_KS;
  return npad * M * OZ_NS + (long)M * (sizeof(unsigned long long) + sizeof(int)) + 1024;
}

// Out (M x M, symmetric) = alpha * w0 * K^T K with w0 = *w0_dev (NULL: 1), K (n x M); M must be a multiple of 128.
// The unweighted / equal-weights case of npgp_wsyrk on the integer tensor cores (exact Ozaki split).
// uniform_count / uniform_target (optional): the kernel only runs when *uniform_count == uniform_target (device-side
// gate, see npgp_wsyrk_hint); accumulate != 0: add to Out instead of overwriting it.
// phase: 0 = slice K and run; 1 = slicing passes only (column maxima, exponents, transposed slices into `work`, so that
// they can run on another stream under an unrelated kernel); 2 = tensor-core kernel only, on the slices a phase-1 call
// with the same K left in `work`.
static int syrk_i8_impl(int n, int M, double alpha, const double* K, long ldk, const double* w0_dev,
                        const double* uniform_count, double uniform_target, int accumulate, int phase, double* Out,
                        long ldo, void* work, long work_bytes, const double* colw, double* colwsum, cudaStream_t stream);

extern "C" int npgp_syrk_i8(int n, int M, double alpha, const double* K, long ldk, const double* w0_dev,
                            const double* uniform_count, double uniform_target, int accumulate, int phase, double* Out,
                            long ldo, void* work, long work_bytes, cudaStream_t stream) {
  return syrk_i8_impl(n, M, alpha, K, ldk, w0_dev, uniform_count, uniform_target, accumulate, phase, Out, ldo, work,
                      work_bytes, nullptr, nullptr, stream);
}

// The slicing passes of npgp_syrk_i8 (phase 1) with the weighted column sums wsum[j] = sum_i w_i K_ij (overwritten) fused
// into the column-maximum pass: the SVGP step gets K^T g_mu from the read of K it needs anyway.
extern "C" int npgp_syrk_i8_prepare(int n, int M, const double* K, long ldk, const double* w, double* wsum, void* work,
                                    long work_bytes, cudaStream_t stream) {
  if ((w == nullptr) != (wsum == nullptr)) return NPGP_EINVAL;
  return syrk_i8_impl(n, M, 1.0, K, ldk, nullptr, nullptr, 0.0, 0, 1, nullptr, 0, work, work_bytes, w, wsum, stream);
}

static int syrk_i8_impl(int n, int M, double alpha, const double* K, long ldk, const double* w0_dev,
                        const double* uniform_count, double uniform_target, int accumulate, int phase, double* Out,
                        long ldo, void* work, long work_bytes, const double* colw, double* colwsum, cudaStream_t stream) {
  if (n < 0 || M < 0) return NPGP_EINVAL;
  if (M == 0) return NPGP_OK;
  if ((phase != 1 && !Out) || (n > 0 && !K) || !work || phase < 0 || phase > 2) return NPGP_EINVAL;
  if (M % OZ_BM) return NPGP_EUNSUPPORTED;
  if (work_bytes < npgp_syrk_i8_workspace_bytes(n, M)) return NPGP_EWORKSPACE;
  const long npad = ((long)n + OZ_KS - 1) / OZ_KS * OZ_KS;
  const int nks = (int)(npad / OZ_KS);
  int8_t* Xs = static_cast<int8_t*>(work);
  unsigned long long* cmax = reinterpret_cast<unsigned long long*>(Xs + npad * M * OZ_NS);
  int* ex = reinterpret_cast<int*>(cmax + M);
  if (phase != 1 && !accumulate) NPGP_CUDA(cudaMemset2DAsync(Out, sizeof(double) * ldo, 0, sizeof(double) * M, M, stream));
  if (colwsum) NPGP_CUDA(cudaMemsetAsync(colwsum, 0, sizeof(double) * M, stream));
  if (n == 0) return NPGP_OK;
  if (phase != 2) {
    NPGP_CUDA(cudaMemsetAsync(cmax, 0, sizeof(unsigned long long) * M, stream));
    const int rows_per_cta = 256;
    dim3 grid(ceil_div(M, 256), ceil_div(n, rows_per_cta));
    oz_colmax_kernel<<<grid, 256, 0, stream>>>(n, M, K, ldk, rows_per_cta, cmax, colw, colwsum);
    NPGP_LAUNCH_CHECK();
    oz_exp_from_max_kernel<<<ceil_div(M, 256), 256, 0, stream>>>(M, cmax, ex);
    NPGP_LAUNCH_CHECK();
    dim3 gs(nks, M / OZ_BM);
    oz_slice_t_kernel<<<gs, 256, 0, stream>>>(n, M, K, ldk, ex, Xs);
    NPGP_LAUNCH_CHECK();
  }
  if (phase == 1) return NPGP_OK;
  static bool attr_set = false;
  constexpr int smem = OZ_SYRK_STAGES * (OZ_A_STAGE + OZ_B_STAGE) + 1024;
  if (!attr_set) {
    NPGP_CUDA(cudaFuncSetAttribute(oz_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int n_rb = M / OZ_BM, n_cb = M / OZ_BN;
  const int n_tiles = n_rb * n_cb - n_rb * (n_rb - 1);
  // chunks of at most 256 stages (int32 exactness); among the admissible chunk counts take the one whose work items fill
  // whole rounds of the persistent CTAs best (cost = rounds x stages per chunk)
  const int min_chunks = ceil_div(nks, 256);
  int best_chunks = min_chunks;
  long best_cost = -1;
  for (int c = min_chunks; c <= min_chunks + 64 && c <= nks; ++c) {
    const int spc_c = ceil_div(nks, c);
    if (spc_c < 16 && c > min_chunks) break;
    const long rounds = ((long)n_tiles * ceil_div(nks, spc_c) + kNumSMs - 1) / kNumSMs;
    const long cost = rounds * (spc_c + 4);  // + epilogue, in units of a k-stage
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_chunks = c;
    }
  }
  const int spc = ceil_div(nks, best_chunks);
  const int items = n_tiles * ceil_div(nks, spc);
  oz_syrk_kernel<<<items < kNumSMs ? items : kNumSMs, OZ_THREADS, smem, stream>>>(M, nks, spc, Xs, ex, alpha, w0_dev, uniform_count,
                                                                                       uniform_target, Out, ldo);
  NPGP_LAUNCH_CHECK();
  dim3 blk(32, 8), grd(ceil_div(M, 32), ceil_div(M, 8));
  oz_symmetrize_upper_kernel<<<grd, blk, 0, stream>>>(M, Out, ldo);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
