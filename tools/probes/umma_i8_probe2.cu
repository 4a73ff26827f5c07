// Probe 2 for the byte-digit int8 engine (csrc/oz8.cu).  Checks on real hardware, against an integer reference:
//   1. mixed signedness: A unsigned 8-bit x B signed 8-bit (idesc a_format = 0, b_format = 1), full value range
//   2. MN-major operands: the SAME 8 x 16-byte core matrices read with the contraction running over the core-matrix ROWS
//      (idesc a_major = b_major = 1), both assignments of LBO / SBO
//   3. A-operand collector reuse (.collector::a::fill / ::use / ::lastuse): three products with one A, three B
//   4. issue-rate micro-benchmark: cycles per 128x64x32 MMA from resident shared memory, 28 MMAs per "k-step" exactly as
//      the engine issues them (7 A digits x the B digits with p + q <= 6), with and without collector reuse, 1 CTA per SM
//      on all SMs -> the measured int8 tensor ceiling used as roofline denominator
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/probes/umma_i8_probe2 tools/probes/umma_i8_probe2.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int MM = 128, NN = 64, KK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

#define MMA_I8(COLL)                                                                                                   \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                       \
               "tcgen05.mma.cta_group::1.kind::i8" COLL " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db),      \
               "r"(idesc), "r"(acc)                                                                                    \
               : "memory")

__device__ __forceinline__ void mma_plain(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { MMA_I8(""); }
__device__ __forceinline__ void mma_fill(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { MMA_I8(".collector::a::fill"); }
__device__ __forceinline__ void mma_use(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { MMA_I8(".collector::a::use"); }
__device__ __forceinline__ void mma_last(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { MMA_I8(".collector::a::lastuse"); }

__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  int spins = 0;
  while (!done && spins < (1 << 24)) {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    ++spins;
  }
  return done != 0;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// mode 0: K-major (A: [r][k], core rows = r), a u8 / b s8
// mode 1/2: MN-major, core rows = k; LBO/SBO assignment 1: LBO = K-group stride, SBO = MN-group stride; 2: swapped
// mode 3: collector test (K-major, three B tiles of 64 rows -> three accumulators)
__global__ void __launch_bounds__(128) probe(const uint8_t* A, const int8_t* B, int32_t* D, int mode, int a_signed, int b_signed) {
  __shared__ __align__(128) uint8_t sA[MM * KK];
  __shared__ __align__(128) int8_t sB[3 * NN * KK];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = (mode == 3) ? 3 : 1;
  if (mode == 0 || mode == 3) {
    for (int e = tid; e < MM * KK; e += 128) {
      const int r = e / KK, k = e % KK;
      sA[(k / 16) * (MM * 16) + (r / 8) * 128 + (r % 8) * 16 + (k % 16)] = A[r * KK + k];
    }
    for (int e = tid; e < nb * NN * KK; e += 128) {
      const int t = e / (NN * KK), r = (e / KK) % NN, k = e % KK;
      sB[t * NN * KK + (k / 16) * (NN * 16) + (r / 8) * 128 + (r % 8) * 16 + (k % 16)] = B[e];
    }
  } else {
    // MN-major: element (mn, k) at (mn/16)*512 + (k/8)*128 + (k%8)*16 + mn%16   (K = 32 -> 4 K-groups per 16-wide MN group)
    for (int e = tid; e < MM * KK; e += 128) {
      const int r = e / KK, k = e % KK;
      sA[(r / 16) * 512 + (k / 8) * 128 + (k % 8) * 16 + (r % 16)] = A[r * KK + k];
    }
    for (int e = tid; e < NN * KK; e += 128) {
      const int r = e / KK, k = e % KK;
      sB[(r / 16) * 512 + (k / 8) * 128 + (k % 8) * 16 + (r % 16)] = B[r * KK + k];
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base;
  if (warp == 0) {
    if (elect_one()) {
      // c S32; a / b format: 0 = unsigned 8 bit, 1 = signed 8 bit
      uint32_t idesc = (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(NN >> 3) << 17) |
                       ((uint32_t)(MM >> 4) << 24);
      if (mode == 0) {
        mma_plain(tbase, make_desc(smem_u32(sA), MM * 16, 128), make_desc(smem_u32(sB), NN * 16, 128), idesc, 0);
      } else if (mode == 1 || mode == 2) {
        idesc |= (1u << 15) | (1u << 16);
        const uint32_t lbo = (mode == 1) ? 128 : 512, sbo = (mode == 1) ? 512 : 128;
        mma_plain(tbase, make_desc(smem_u32(sA), lbo, sbo), make_desc(smem_u32(sB), lbo, sbo), idesc, 0);
      } else {
        const uint64_t da = make_desc(smem_u32(sA), MM * 16, 128);
        mma_fill(tbase + 0 * NN, da, make_desc(smem_u32(sB + 0 * NN * KK), NN * 16, 128), idesc, 0);
        mma_use(tbase + 1 * NN, da, make_desc(smem_u32(sB + 1 * NN * KK), NN * 16, 128), idesc, 0);
        mma_last(tbase + 2 * NN, da, make_desc(smem_u32(sB + 2 * NN * KK), NN * 16, 128), idesc, 0);
      }
      commit(&mbar);
    }
    __syncwarp();
  }
  if (!wait_bar(&mbar, 0) && lane == 0) printf("warp %d: mbarrier wait timed out\n", warp);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int t = 0; t < nb; ++t)
    for (int c0 = 0; c0 < NN; c0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * NN + c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(t * MM + warp * 32 + lane) * NN + c0 + j] = (int32_t)v[j];
    }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
}

// issue-rate benchmark: `reps` k-steps of 28 MMAs (A digit p = 0..6, B digit q = 0..6-p, accumulator p+q) from resident
// shared memory (contents irrelevant).  variant 0: plain; 1: collector::a fill/use/lastuse per A digit; 2: MN-major plain
constexpr int NSD = 7;
__global__ void __launch_bounds__(128, 1) rate(int reps, int variant, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* sA = sm;                       // 7 x 4 KB
  uint8_t* sB = sm + NSD * MM * KK;       // 7 x 2 KB
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < NSD * (MM + NN) * KK; e += 128) sm[e] = (uint8_t)(e * 7 + 3);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const bool leader = elect_one();
    uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(MM >> 4) << 24);
    uint32_t lboA = MM * 16, sboA = 128, lboB = NN * 16, sboB = 128;
    if (variant == 2) {
      idesc |= (1u << 15) | (1u << 16);
      lboA = lboB = 128;
      sboA = sboB = 512;
    }
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (leader) {
#pragma unroll
        for (int p = 0; p < NSD; ++p) {
          const uint64_t da = make_desc(smem_u32(sA + p * MM * KK), lboA, sboA);
#pragma unroll
          for (int q = 0; q + p < NSD; ++q) {
            const uint64_t db = make_desc(smem_u32(sB + q * NN * KK), lboB, sboB);
            const uint32_t d = tbase + (uint32_t)((p + q) * NN);
            const uint32_t acc = (r > 0 || p > 0) ? 1u : 0u;
            if (variant != 1 || p == NSD - 1) mma_plain(d, da, db, idesc, acc);
            else if (q == 0) mma_fill(d, da, db, idesc, acc);
            else if (q + p == NSD - 1) mma_last(d, da, db, idesc, acc);
            else mma_use(d, da, db, idesc, acc);
          }
        }
      }
      __syncwarp();
    }
    if (leader) commit(&mbar);
    __syncwarp();
    wait_bar(&mbar, 0);
    t1 = clock64();
    if (leader && blockIdx.x == 0) out[variant] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main() {
  std::vector<uint8_t> hA(MM * KK);
  std::vector<int8_t> hB(3 * NN * KK);
  srand(1);
  for (auto& v : hA) v = (uint8_t)(rand() % 256);
  for (auto& v : hB) v = (int8_t)(rand() % 256 - 128);
  hA[0] = 255; hA[1] = 255; hB[0] = -128; hB[1] = 127;
  uint8_t* dA;
  int8_t* dB;
  int32_t* dD;
  cudaMalloc(&dA, hA.size());
  cudaMalloc(&dB, hB.size());
  cudaMalloc(&dD, sizeof(int32_t) * 3 * MM * NN);
  cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
  const char* names[4] = {"K-major u8 x s8", "MN-major (LBO=K-group 128, SBO=MN-group 512)", "MN-major (LBO=512, SBO=128)",
                          "collector fill/use/lastuse"};
  for (int mode = 0; mode < 4; ++mode)
    for (int combo = 0; combo < (mode == 0 ? 4 : 1); ++combo) {
      const int a_signed = mode == 0 ? (combo >> 1) : 0, b_signed = mode == 0 ? (combo & 1) : 1;
      cudaMemset(dD, 0xff, sizeof(int32_t) * 3 * MM * NN);
      probe<<<1, 128>>>(dA, dB, dD, mode, a_signed, b_signed);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
      const int nb = mode == 3 ? 3 : 1;
      std::vector<int32_t> hD(3 * MM * NN);
      cudaMemcpy(hD.data(), dD, sizeof(int32_t) * 3 * MM * NN, cudaMemcpyDeviceToHost);
      long bad = 0;
      for (int t = 0; t < nb; ++t)
        for (int i = 0; i < MM; ++i)
          for (int j = 0; j < NN; ++j) {
            int32_t ref = 0;
            for (int k = 0; k < KK; ++k) {
              const int32_t av = a_signed ? (int32_t)(int8_t)hA[i * KK + k] : (int32_t)hA[i * KK + k];
              const int32_t bv = b_signed ? (int32_t)hB[t * NN * KK + j * KK + k] : (int32_t)(uint8_t)hB[t * NN * KK + j * KK + k];
              ref += av * bv;
            }
            if (ref != hD[(t * MM + i) * NN + j]) ++bad;
          }
      printf("mode %d [%s] a %s x b %s: %ld of %d entries wrong\n", mode, names[mode], a_signed ? "s8" : "u8",
             b_signed ? "s8" : "u8", bad, nb * MM * NN);
    }
  long long* dout;
  cudaMalloc(&dout, 8 * sizeof(long long));
  cudaMemset(dout, 0, 8 * sizeof(long long));
  const int smem = NSD * (MM + NN) * KK;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 4096;
  for (int grid : {1, 148})
    for (int variant = 0; variant < 3; ++variant) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      rate<<<grid, 128, smem>>>(64, variant, dout);  // warm-up
      cudaEventRecord(a);
      rate<<<grid, 128, smem>>>(reps, variant, dout);
      cudaEventRecord(b);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("rate: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      long long cyc;
      cudaMemcpy(&cyc, dout + variant, sizeof(cyc), cudaMemcpyDeviceToHost);
      const double mmas = (double)reps * 28;
      printf("rate grid=%3d variant %d (%s): %.2f cycles per 128x64x32 MMA (CTA 0), %.3f ms, %.1f int8 TOP/s chip-wide\n", grid,
             variant, variant == 0 ? "plain" : variant == 1 ? "collector-a" : "MN-major", cyc / mmas, ms,
             grid * mmas * 2.0 * MM * NN * KK / (ms * 1e-3) / 1e12);
    }
  return 0;
}
