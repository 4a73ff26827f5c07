// Blocked FP64 Cholesky of the M x M inducing covariance and the inverse of its factor, M up to a few thousand.
//
//   npgp_potrf_inv_lower:  A = L L^T (L returned in A, strict upper triangle zero),  P = L^-1 (lower),
//                          info = 0 or (1-based) index of the first non-positive pivot (LAPACK convention) so the
//                          host can reproduce psd_safe_cholesky's jitter ladder.
//
// Replaces psd_safe_cholesky + triangular_solve(eye, chol) (reference models/gibbs_kernels.py:197-208) and the
// Cholesky inside GPyTorch's VariationalStrategy (models/dgps.py:29-33).
//
// The factorisation is latency bound (M^3/3 flops is nothing; the critical path is M/64 dependent panel steps), so the
// design minimises the number and the length of dependent launches:
//   * ONE kernel per 64-column panel step j.  The CTA that owns trailing tile (i,k) recomputes the two panel tiles it
//     needs (L_ij = A_ij P_jj^T, L_kj = A_kj P_jj^T: redundant 64^3 DMMA work instead of a separate TRSM launch and a
//     grid-wide dependency), applies A_ik -= L_ij L_kj^T, and the CTA of tile (j+1,j+1) goes on to factor AND invert the
//     freshly updated diagonal block, so the next step can start as soon as this kernel retires (look-ahead of 1).
//   * The 64x64 diagonal block is factored in shared memory by a blocked (8-wide) right-looking elimination whose loop
//     bodies are small and re-executed (a fully unrolled 64-column version measured 110 us: instruction-fetch bound),
//     one rsqrt per pivot, and inverted by recursive doubling 8 -> 16 -> 32 -> 64 (see factor_invert_64).
//   * L^-1 is assembled by recursive doubling, inv([[L11,0],[L21,L22]]) = [[P11,0],[-P22 L21 P11, P22]]: two batched
//     kernels per level (all block pairs of a level in one launch), skipping the structurally zero blocks.
// All tile products run on the FP64 tensor pipe (DMMA.8x8x4).
#include "chol_tiles.cuh"
#include "common.cuh"

namespace npgp {

// first diagonal block
__global__ void __launch_bounds__(CT) potrf_first_kernel(int M, const double* __restrict__ A, long lda,
                                                         double* __restrict__ L, long ldl, double* __restrict__ P,
                                                         long ldp, int* __restrict__ info) {
  extern __shared__ double sm[];
  double *sL = sm, *sX = sm + TILE_SMEM, *tmp = sm + 2 * TILE_SMEM;
  load_tile(sL, A, lda, 0, 0, M, true);
  __syncthreads();
  factor_invert_64(sL, sX, tmp, 0, info);
  store_tile(sL, L, ldl, 0, 0, M);
  store_tile(sX, P, ldp, 0, 0, M);
}

// panel step j: for every trailing tile (i,k), j < k <= i:  A_ik -= (A_ij P_jj^T)(A_kj P_jj^T)^T; tiles with k == j+1
// also publish L_ij; tile (j+1,j+1) additionally factors + inverts the updated block.
__global__ void __launch_bounds__(CT) potrf_step_kernel(int M, int nblk, int j, double* __restrict__ A, long lda,
                                                        double* __restrict__ L, long ldl, double* __restrict__ P,
                                                        long ldp, int* __restrict__ info) {
  extern __shared__ double sm[];
  double *bufA = sm, *bufB = sm + TILE_SMEM, *bufP = sm + 2 * TILE_SMEM;
  // tile index -> (a, b), 0 <= b <= a
  int t = blockIdx.x, a = 0;
  while ((a + 1) * (a + 2) / 2 <= t) ++a;
  const int b = t - a * (a + 1) / 2;
  const int i = j + 1 + a, k = j + 1 + b;
  const WarpPos p;
  load_tile(bufP, P, ldp, j * TB, j * TB, M, false);
  load_tile(bufA, A, lda, i * TB, j * TB, M, false);
  if (k != i) load_tile(bufB, A, lda, k * TB, j * TB, M, false);
  __syncthreads();
  double accI[2][4][2], accK[2][4][2];
  acc_zero(accI);
  mma_64<false, true>(accI, bufA, bufP, p);  // L_ij = A_ij P_jj^T
  if (k != i) {
    acc_zero(accK);
    mma_64<false, true>(accK, bufB, bufP, p);
  }
  __syncthreads();
  acc_to_smem(accI, bufA, p, 1.0);
  if (k != i) acc_to_smem(accK, bufB, p, 1.0);
  __syncthreads();
  if (b == 0) store_tile(bufA, L, ldl, i * TB, j * TB, M);
  double acc[2][4][2];
  acc_zero(acc);
  mma_64<false, true>(acc, bufA, (k != i) ? bufB : bufA, p);  // L_ij L_kj^T
  if (a == 0) {
    // diagonal tile of the next panel: update in shared memory, then factor + invert
    load_tile(bufP, A, lda, i * TB, i * TB, M, true);  // bufP is free (all reads of P_jj happened before the barrier)
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int r = p.wm0 + mt * 8 + p.g, c = p.wn0 + nt * 8 + 2 * p.t4;
        bufP[r * TLD + c] -= acc[mt][nt][0];
        bufP[r * TLD + c + 1] -= acc[mt][nt][1];
      }
    __syncthreads();
    factor_invert_64(bufP, bufB, bufA, i * TB, info);  // bufA (L_ij, no longer needed) serves as scratch
    store_tile(bufP, L, ldl, i * TB, i * TB, M);
    store_tile(bufB, P, ldp, i * TB, i * TB, M);
  } else {
    acc_axpy_global(acc, A, lda, i * TB, k * TB, M, p, -1.0, true);
  }
}

// recursive-doubling level, phase 0: T = L21 P11 for every block pair of the level (blockIdx.z = pair);
// phase 1: P21 = -P22 T.  Blocks are b x b (b a multiple of 64), tiles 64 x 64.
__global__ void __launch_bounds__(CT) trinv_level_kernel(int M, int b, int phase, const double* __restrict__ Lm,
                                                         long ldl, double* __restrict__ P, long ldp,
                                                         double* __restrict__ work) {
  extern __shared__ double sm[];
  double *bufA = sm, *bufB = sm + TILE_SMEM;
  const int s = blockIdx.z * 2 * b;     // pair start
  const int r1 = s + b;                 // first row of block 2
  if (r1 >= M) return;
  const int b2 = min(b, M - r1);
  const int tr = blockIdx.y, tc = blockIdx.x;  // tile row within block 2, tile col within block 1
  if (tr * TB >= b2) return;
  const WarpPos p;
  double* T = work + (long)blockIdx.z * b * b;  // (b2 x b), ld = b
  double acc[2][4][2];
  acc_zero(acc);
  const int nkb = b / TB;
  // phase 0: k runs over block-1 tiles with k >= tc (P11 lower);  phase 1: over block-2 tiles with k <= tr (P22 lower)
  const int k_lo = (phase == 0) ? tc : 0;
  const int k_hi = (phase == 0) ? nkb : min(tr + 1, (b2 + TB - 1) / TB);
  for (int kb = k_lo; kb < k_hi; ++kb) {
    __syncthreads();
    if (phase == 0) {
      load_tile(bufA, Lm, ldl, r1 + tr * TB, s + kb * TB, M, false);
      load_tile(bufB, P, ldp, s + kb * TB, s + tc * TB, M, false);
    } else {
      load_tile(bufA, P, ldp, r1 + tr * TB, r1 + kb * TB, M, false);
      // T tile (rows kb, cols tc) of the b2 x b workspace; reuse load_tile with a virtual M bound
      for (int e = threadIdx.x; e < TB * TB; e += CT) {
        const int r = e >> 6, c = e & 63;
        const int rr = kb * TB + r;
        bufB[r * TLD + c] = (rr < b2) ? T[(long)rr * b + tc * TB + c] : 0.0;
      }
    }
    __syncthreads();
    mma_64<false, false>(acc, bufA, bufB, p);
  }
  if (phase == 0) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int r = tr * TB + p.wm0 + mt * 8 + p.g, c = tc * TB + p.wn0 + nt * 8 + 2 * p.t4;
        if (r < b2) {
          T[(long)r * b + c] = acc[mt][nt][0];
          T[(long)r * b + c + 1] = acc[mt][nt][1];
        }
      }
  } else {
    acc_axpy_global(acc, P, ldp, r1 + tr * TB, s + tc * TB, M, p, -1.0, false);
  }
}

__global__ void copy_matrix_kernel(int M, const double* __restrict__ S, long lds, double* __restrict__ D, long ldd) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * M) return;
  const int r = (int)(idx / M), c = (int)(idx % M);
  D[(long)r * ldd + c] = S[(long)r * lds + c];
}

__global__ void zero_matrix_kernel(int M, double* __restrict__ D, long ldd) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * M) return;
  D[(long)(idx / M) * ldd + (idx % M)] = 0.0;
}

int dgemm_impl(int transA, int transB, int M, int N, int K, double alpha, const double* A, long lda, const double* B,
               long ldb, double beta, double* C, long ldc, int tri_a, int tri_b, int out_tri, cudaStream_t st);  // dgemm.cu

constexpr int FIRST_SMEM = 3 * TILE_SMEM * (int)sizeof(double);
constexpr int STEP_SMEM = 3 * TILE_SMEM * (int)sizeof(double);
constexpr int LEVEL_SMEM = 2 * TILE_SMEM * (int)sizeof(double);

static long work_doubles(int M) {
  const long Mp = ((long)M + TB - 1) / TB * TB;
  return Mp * Mp + Mp * Mp / 2 + (long)TB * Mp;
}

int potrf_inv_impl(int M, double* A, long lda, double* P, long ldp, double* work, int* info, cudaStream_t st) {
  // (the attribute is per device and the call is cheap: set on every call, so a process that touches a second GPU is fine)
  NPGP_CUDA(cudaFuncSetAttribute(potrf_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FIRST_SMEM));
  NPGP_CUDA(cudaFuncSetAttribute(potrf_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STEP_SMEM));
  NPGP_CUDA(cudaFuncSetAttribute(trinv_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEVEL_SMEM));
  const int nblk = (M + TB - 1) / TB;
  double* L = work;  // M x M, ld = M
  const long ldl = M + (M & 1);
  double* T = work + (long)M * ldl;
  const unsigned nbk = (unsigned)(((long)M * M + 255) / 256);
  NPGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  NPGP_CUDA(cudaMemsetAsync(L, 0, sizeof(double) * M * ldl, st));
  zero_matrix_kernel<<<nbk, 256, 0, st>>>(M, P, ldp);
  NPGP_LAUNCH_CHECK();
  potrf_first_kernel<<<1, CT, FIRST_SMEM, st>>>(M, A, lda, L, ldl, P, ldp, info);
  NPGP_LAUNCH_CHECK();
  for (int j = 0; j + 1 < nblk; ++j) {
    const int r = nblk - 1 - j;
    potrf_step_kernel<<<r * (r + 1) / 2, CT, STEP_SMEM, st>>>(M, nblk, j, A, lda, L, ldl, P, ldp, info);
    NPGP_LAUNCH_CHECK();
  }
  for (int b = TB; b < M; b *= 2) {
    const int pairs = (M + 2 * b - 1) / (2 * b);
    if (b >= 256) {
      // large blocks: the pipelined (cp.async, split-K) DMMA GEMM of dgemm.cu, one call per block pair and phase
      for (int pr = 0; pr < pairs; ++pr) {
        const int s0 = pr * 2 * b, r1 = s0 + b;
        if (r1 >= M) continue;
        const int b2 = min(b, M - r1);
        double* Tp = T + (long)pr * b * b;
        int rc = dgemm_impl(0, 0, b2, b, b, 1.0, L + (long)r1 * ldl + s0, ldl, P + (long)s0 * ldp + s0, ldp, 0.0, Tp, b,
                            0, 1, 0, st);
        if (rc != NPGP_OK) return rc;
        rc = dgemm_impl(0, 0, b2, b, b2, -1.0, P + (long)r1 * ldp + r1, ldp, Tp, b, 0.0, P + (long)r1 * ldp + s0, ldp,
                        1, 0, 0, st);
        if (rc != NPGP_OK) return rc;
      }
      continue;
    }
    dim3 grid(b / TB, b / TB, pairs);
    trinv_level_kernel<<<grid, CT, LEVEL_SMEM, st>>>(M, b, 0, L, ldl, P, ldp, T);
    NPGP_LAUNCH_CHECK();
    trinv_level_kernel<<<grid, CT, LEVEL_SMEM, st>>>(M, b, 1, L, ldl, P, ldp, T);
    NPGP_LAUNCH_CHECK();
  }
  copy_matrix_kernel<<<nbk, 256, 0, st>>>(M, L, ldl, A, lda);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

}  // namespace npgp

using namespace npgp;

extern "C" long npgp_potrf_workspace_bytes(int M) { return work_doubles(M) * (long)sizeof(double); }

extern "C" int npgp_potrf_inv_lower(int M, double* A, long lda, double* P, long ldp, void* work, long work_bytes,
                                    int* info, cudaStream_t stream) {
  if (M < 0) return NPGP_EINVAL;
  if (M == 0) return NPGP_OK;
  if (!A || !P || !info || !work) return NPGP_EINVAL;
  if ((lda & 1) || (ldp & 1) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(P) & 15))
    return NPGP_EUNSUPPORTED;
  if (work_bytes < npgp_potrf_workspace_bytes(M)) return NPGP_EWORKSPACE;
  return potrf_inv_impl(M, A, lda, P, ldp, static_cast<double*>(work), info, stream);
}
