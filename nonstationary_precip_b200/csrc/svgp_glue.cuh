// Small kernels shared by the C orchestrations of the SVGP-Gibbs step (svgp_step.cu) and of the DSVI layer (dsvi_layer.cu):
// O(M^2) / O(M) glue of the replicated Z-side chains.  `static`: every translation unit gets its own copy.
#pragma once
#include "common.cuh"

namespace npgp {

// dst = tril(src) (M x M, contiguous)
static __global__ void svgp_tril_copy_kernel(int M, const double* __restrict__ src, double* __restrict__ dst) {
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
  if (r < M && c < M) dst[(long)r * M + c] = (c <= r) ? src[(long)r * M + c] : 0.0;
}

static __global__ void svgp_add_diag_kernel(int M, double* __restrict__ A, long lda, double v) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < M) A[(long)i * lda + i] += v;
}

// out (d x M): out[k][j] = lam[k]
static __global__ void svgp_bcast_rows_kernel(int d, int M, const double* __restrict__ lam, double* __restrict__ out) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j < M)
    for (int k = 0; k < d; ++k) out[(long)k * M + j] = lam[k];
}

// X <- -Phi(X + m dm^T): lower triangle, diagonal halved, upper zeroed
static __global__ void svgp_addr_phi_kernel(int M, double* __restrict__ X, const double* __restrict__ m, const double* __restrict__ dm) {
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
  if (r >= M || c >= M) return;
  double* p = X + (long)r * M + c;
  const double v = *p + m[r] * dm[c];
  *p = (c < r) ? -v : ((c == r) ? -0.5 * v : 0.0);
}

// out = 0.5 (A + A^T)
static __global__ void svgp_sym_avg_kernel(int M, const double* __restrict__ A, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int r = bx + k, c = by + threadIdx.x;  // transposed block
    tile[k][threadIdx.x] = (r < M && c < M) ? A[(long)r * M + c] : 0.0;
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int r = by + k, c = bx + threadIdx.x;
    if (r < M && c < M) out[(long)r * M + c] = 0.5 * (A[(long)r * M + c] + tile[threadIdx.x][k]);
  }
}

// Skinny triangular products of the M x k (k <= 4) solves against the inverse factor P (lower): the generic DMMA GEMM spends
// ~30 us of pure latency on them.  Y = P R: one warp per output row (fixed lane / shuffle order).
template <int KP>
static __global__ void __launch_bounds__(256) svgp_tri_skinny_n_kernel(int M, const double* __restrict__ P, const double* __restrict__ R,
                                                                       double* __restrict__ Y) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= M) return;
  double a[KP], b[KP];
#pragma unroll
  for (int c = 0; c < KP; ++c) a[c] = b[c] = 0.0;
  int j = lane;
  for (; j + 32 <= i; j += 64) {  // two independent chains
    const double p0 = P[(long)i * M + j], p1 = P[(long)i * M + j + 32];
#pragma unroll
    for (int c = 0; c < KP; ++c) {
      a[c] = fma(p0, R[(long)j * KP + c], a[c]);
      b[c] = fma(p1, R[(long)(j + 32) * KP + c], b[c]);
    }
  }
  if (j <= i) {
    const double p0 = P[(long)i * M + j];
#pragma unroll
    for (int c = 0; c < KP; ++c) a[c] = fma(p0, R[(long)j * KP + c], a[c]);
  }
#pragma unroll
  for (int c = 0; c < KP; ++c) {
    const double t = warp_sum(a[c] + b[c]);
    if (lane == 0) Y[(long)i * KP + c] = t;
  }
}

// W = P^T T: one warp per output row j (= column j of P), lanes over the rows i = j + lane, j + lane + 32, ...; the column
// access is a 32-line gather per load, but P was written a moment ago and sits in L2, and 1024 warps keep ~M^2/2 gathered
// sectors in flight (a first version with a coalesced walk down the rows, 32 CTAs x 128 dependent iterations, took 60 us).
// Fixed lane / shuffle order: bitwise reproducible, no atomics.
template <int KP>
static __global__ void __launch_bounds__(256) svgp_tri_skinny_t_kernel(int M, const double* __restrict__ P, const double* __restrict__ T,
                                                                       double* __restrict__ W) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= M) return;
  double a[KP], b[KP];
#pragma unroll
  for (int c = 0; c < KP; ++c) a[c] = b[c] = 0.0;
  int i = j + lane;
  for (; i + 32 < M; i += 64) {  // two independent chains
    const double p0 = P[(long)i * M + j], p1 = P[(long)(i + 32) * M + j];
#pragma unroll
    for (int c = 0; c < KP; ++c) {
      a[c] = fma(p0, T[(long)i * KP + c], a[c]);
      b[c] = fma(p1, T[(long)(i + 32) * KP + c], b[c]);
    }
  }
  if (i < M) {
    const double p0 = P[(long)i * M + j];
#pragma unroll
    for (int c = 0; c < KP; ++c) a[c] = fma(p0, T[(long)i * KP + c], a[c]);
  }
#pragma unroll
  for (int c = 0; c < KP; ++c) {
    const double t = warp_sum(a[c] + b[c]);
    if (lane == 0) W[(long)j * KP + c] = t;
  }
}

}  // namespace npgp
