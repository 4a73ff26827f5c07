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

// Row-side reduction of the pairwise backward kernels.  Every lane holds NRC partial sums for the current row (over its
// own columns); the warp needs their sums over the 32 lanes.  A shuffle tree costs 5*NRC add+shuffle steps per row.
// Instead each lane parks its partials in a warp-private, padded shared-memory panel and every RB = 32/NRC rows one
// lane per (row, component) adds up the 32 entries of a panel line (conflict free: line length 33): about
// 32/RB adds per row instead of 5*NRC.
template <int NRC>
struct RowReducer {
  static constexpr int RB = (32 / NRC) > 0 ? (32 / NRC) : 1;   // rows per batch
  static constexpr int LINES = RB * NRC;
  static constexpr int PANEL = LINES * 33;                      // doubles per warp
  double* panel;  // this warp's panel
  int lane;
  __device__ __forceinline__ RowReducer(double* smem_base, int warp, int lane_) : panel(smem_base + warp * PANEL), lane(lane_) {}
  __device__ __forceinline__ void put(int r, int comp, double v) { panel[((r % RB) * NRC + comp) * 33 + lane]  = v; }
  // call after the puts of row r; `out(row, comp, sum)` is invoked by one lane per finished (row, comp)
  template <class F>
  __device__ __forceinline__ void flush_if_due(int r, int nr, F out) {
    const int slot = r % RB;
    if (slot != RB - 1 && r != nr - 1) return;
    __syncwarp();
    const int nvals = (slot + 1) * NRC;
    for (int v = lane; v < nvals; v += 32) {
      const double* line = panel + v * 33;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        a0 += line[k];
        a1 += line[k + 1];
        a2 += line[k + 2];
        a3 += line[k + 3];
      }
      out(r - slot + v / NRC, v % NRC, (a0 + a1) + (a2 + a3));
    }
    __syncwarp();
  }
};

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

// Register queue that keeps the upstream-gradient entries G[i .. i+DEPTH-1, columns of this thread] in flight, so that
// the global-load latency of the backward kernels' one G read per pair is covered by DEPTH rows of pair math.
#ifndef NPGP_GPF_DEPTH
#define NPGP_GPF_DEPTH 4
#endif
template <int CPT, int DEPTH = NPGP_GPF_DEPTH>
struct GPrefetch {
  const double* base;  // &G[0, jbase] or nullptr
  long ldg;
  int row_end;
  bool v[CPT], vec;
  double q[DEPTH][CPT];
  __device__ __forceinline__ GPrefetch(const GSpec& g, int jbase, const bool* valid, int vec_ok, int row_begin,
                                       int row_end_)
      : base(g.Gm ? g.Gm + jbase : nullptr), ldg(g.ldg), row_end(row_end_) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) v[c] = valid[c];
    vec = (CPT == 2) && vec_ok && valid[CPT - 1];
#pragma unroll
    for (int k = 0; k < DEPTH; ++k) fetch(row_begin + k, q[k]);
  }
  __device__ __forceinline__ void fetch(int i, double* out) const {
#pragma unroll
    for (int c = 0; c < CPT; ++c) out[c] = 0.0;
    if (base == nullptr || i >= row_end) return;
    const double* grow = base + (long)i * ldg;
    if (vec) {
      const double2 t = *reinterpret_cast<const double2*>(grow);
      out[0] = t.x;
      out[CPT - 1] = t.y;
    } else {
#pragma unroll
      for (int c = 0; c < CPT; ++c)
        if (v[c]) out[c] = grow[c];
    }
  }
  // returns row i's entries and issues the load of row i + DEPTH
  __device__ __forceinline__ void pop(int i, double* out) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) out[c] = q[0][c];
#pragma unroll
    for (int k = 0; k + 1 < DEPTH; ++k)
#pragma unroll
      for (int c = 0; c < CPT; ++c) q[k][c] = q[k + 1][c];
    fetch(i + DEPTH, q[DEPTH - 1]);
  }
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
