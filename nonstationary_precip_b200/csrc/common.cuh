// Shared device helpers and the C-ABI return-code convention (see include/npgp.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NPGP_OK 0
#define NPGP_EINVAL (-1)      // bad argument (null pointer, non-positive size, unsupported dimension)
#define NPGP_EUNSUPPORTED (-2)
#define NPGP_EWORKSPACE (-3)  // workspace too small

// diagnostic only: number of kernel launches issued through this library (read by bench.py for "gpu_launches")
extern "C" long npgp_launch_counter;

#define NPGP_LAUNCH_CHECK()                       \
  do {                                            \
    ++npgp_launch_counter;                        \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

#define NPGP_CUDA(call)                           \
  do {                                            \
    cudaError_t e__ = (call);                     \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

namespace npgp {

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum for blockDim.x <= 1024; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* smem /* >= 32 doubles */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (lane < (int)((blockDim.x + 31) >> 5)) ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

__device__ __forceinline__ void st_v2(double* p, double a, double b) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// ---- tiling shared by the pairwise (Gibbs / RBF) tile kernels ----
constexpr int kNT = 128;         // threads per CTA
constexpr int kCPT = 2;          // columns per thread (16-byte stores)
constexpr int kTJ = kNT * kCPT;  // columns per CTA
constexpr int kTI = 32;          // rows staged in shared memory per chunk

// upstream gradient of a pairwise kernel matrix:  G_ij = rowscale_i * Gm_ij + rowvec_i * colvec_j  (either part optional)
struct GSpec {
  const double* Gm;
  long ldg;
  const double* rowscale;
  const double* rowvec;
  const double* colvec;
};

// rows handled by one CTA: enough CTAs for `waves_target` CTAs per SM, at least one staged chunk each
static inline int pick_rows_per_cta(int n1, int col_tiles, int waves_target) {
  long want = (long)kNumSMs * waves_target;
  long row_ctas = (want + col_tiles - 1) / col_tiles;
  long rpc = (n1 + row_ctas - 1) / row_ctas;
  rpc = ((rpc + kTI - 1) / kTI) * kTI;
  if (rpc < kTI) rpc = kTI;
  return (int)rpc;
}

}  // namespace npgp
