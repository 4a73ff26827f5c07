// Probe: one tcgen05.mma.kind::i8 (M=128, N=64, K=32) from shared memory (K-major, no swizzle) into TMEM, read back with
// tcgen05.ld and compared with an integer reference.  Groundwork for an exact int8 (Ozaki) replacement of the FP64 DMMA
// GEMMs.  Descriptor encodings: cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor) of the vendored CUTLASS.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/probes/umma_i8_probe tools/probes/umma_i8_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int MM = 128, NN = 64, KK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);                 // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;       // leading byte offset, bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;       // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                                 // version = 1 (Blackwell)
  return d;                                               // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

__global__ void __launch_bounds__(128) probe(const int8_t* A, const int8_t* B, int32_t* D, int swap_lbo_sbo) {
  __shared__ __align__(128) int8_t sA[MM * KK];
  __shared__ __align__(128) int8_t sB[NN * KK];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // canonical K-major no-swizzle layout: offset(r, c16) = c16 * (rows*16) + (r/8)*128 + (r%8)*16
  for (int e = tid; e < MM * KK; e += 128) {
    const int r = e / KK, k = e % KK;
    sA[(k / 16) * (MM * 16) + (r / 8) * 128 + (r % 8) * 16 + (k % 16)] = A[r * KK + k];
  }
  for (int e = tid; e < NN * KK; e += 128) {
    const int r = e / KK, k = e % KK;
    sB[(k / 16) * (NN * 16) + (r / 8) * 128 + (r % 8) * 16 + (k % 16)] = B[r * KK + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base;
  if (tid == 0) {
    uint32_t lboA = MM * 16, sboA = 128, lboB = NN * 16, sboB = 128;
    if (swap_lbo_sbo) { uint32_t t = lboA; lboA = sboA; sboA = t; t = lboB; lboB = sboB; sboB = t; }
    const uint64_t da = make_desc(smem_u32(sA), lboA, sboA), db = make_desc(smem_u32(sB), lboB, sboB);
    // instruction descriptor: c_format S32 (2) @4, a_format INT8 (1) @7, b_format INT8 (1) @10, K-major both, N>>3 @17, M>>4 @24
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(MM >> 4) << 24);
    const uint32_t zero = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tbase),
        "l"(da), "l"(db), "r"(idesc), "r"(zero), "r"(zero), "r"(zero), "r"(zero), "r"(zero)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                 : "memory");
  }
  // wait for the MMA (phase 0)
  {
    uint32_t done = 0;
    int spins = 0;
    while (!done && spins < (1 << 22)) {
      asm volatile(
          "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&mbar)), "r"(0u)
          : "memory");
      ++spins;
    }
    if (!done && lane == 0) printf("warp %d: mbarrier wait timed out\n", warp);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w .. 32w+31, 64 columns, 8 at a time
  for (int c0 = 0; c0 < NN; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * NN + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64));
}

int main() {
  std::vector<int8_t> hA(MM * KK), hB(NN * KK);
  srand(1);
  for (auto& v : hA) v = (int8_t)(rand() % 127 - 63);
  for (auto& v : hB) v = (int8_t)(rand() % 127 - 63);
  int8_t *dA, *dB;
  int32_t* dD;
  cudaMalloc(&dA, hA.size());
  cudaMalloc(&dB, hB.size());
  cudaMalloc(&dD, sizeof(int32_t) * MM * NN);
  cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
  for (int swap = 0; swap < 2; ++swap) {
    cudaMemset(dD, 0xff, sizeof(int32_t) * MM * NN);
    probe<<<1, 128>>>(dA, dB, dD, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("swap=%d: CUDA error %s\n", swap, cudaGetErrorString(e)); return 1; }
    std::vector<int32_t> hD(MM * NN);
    cudaMemcpy(hD.data(), dD, sizeof(int32_t) * MM * NN, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < MM; ++i)
      for (int j = 0; j < NN; ++j) {
        int32_t ref = 0;
        for (int k = 0; k < KK; ++k) ref += (int32_t)hA[i * KK + k] * (int32_t)hB[j * KK + k];
        if (ref != hD[i * NN + j]) ++bad;
      }
    printf("swap_lbo_sbo=%d: %ld of %d entries wrong; D[0][0..3] = %d %d %d %d\n", swap, bad, MM * NN, hD[0], hD[1], hD[2], hD[3]);
  }
  return 0;
}
