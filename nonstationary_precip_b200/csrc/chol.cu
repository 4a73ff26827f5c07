// Blocked FP64 Cholesky of the M x M inducing covariance and the inverse of its factor, M up to a few thousand.
//
//   npgp_potrf_inv_lower:  A = L L^T (in place, strict upper triangle zeroed),  P = L^-1 (lower, upper zeroed),
//                          info = 0 or (1-based) index of the first non-positive pivot (LAPACK convention) so the
//                          host can reproduce psd_safe_cholesky's jitter ladder.
//
// Replaces psd_safe_cholesky + triangular_solve(eye, chol) (reference models/gibbs_kernels.py:197-208) and the
// Cholesky inside GPyTorch's VariationalStrategy (models/dgps.py:29-33).
//
// Right-looking blocked algorithm, NB = 64: the diagonal block is factored AND inverted by one CTA in shared memory;
// the panel solve is then a GEMM with the inverted block, the trailing update a lower-triangular rank-64 DMMA update.
// L^-1 is assembled afterwards by recursive doubling:  inv([[L11,0],[L21,L22]]) = [[P11,0],[-P22 L21 P11, P22]],
// i.e. two DMMA GEMMs per block pair and level.
#include "common.cuh"

namespace npgp {

int dgemm_impl(int transA, int transB, int M, int N, int K, double alpha, const double* A, long lda, const double* B,
               long ldb, double beta, double* C, long ldc, int tri_a, int tri_b, int out_tri, cudaStream_t st);

constexpr int NB = 64;
constexpr int LDB = NB + 1;

// One CTA (256 threads).  A11 (nb x nb at A, leading dim lda) -> L11 written back (upper zeroed); inverse to Pd (ldp).
__global__ void __launch_bounds__(256) potrf_diag_kernel(int nb, double* __restrict__ A, long lda,
                                                         double* __restrict__ Pd, long ldp, int global_offset,
                                                         int* __restrict__ info) {
  extern __shared__ double dyn_smem[];
  double(*s)[LDB] = reinterpret_cast<double(*)[LDB]>(dyn_smem);
  double(*x)[LDB] = reinterpret_cast<double(*)[LDB]>(dyn_smem + NB * LDB);
  const int tid = threadIdx.x;
  for (int t = tid; t < NB * NB; t += 256) {
    const int r = t / NB, c = t % NB;
    s[r][c] = (r < nb && c <= r) ? A[(long)r * lda + c] : ((r == c) ? 1.0 : 0.0);
    x[r][c] = 0.0;
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (tid == 0) {
      const double dj = s[j][j];
      if (!(dj > 0.0)) atomicCAS(info, 0, global_offset + j + 1);
      s[j][j] = sqrt(dj);
    }
    __syncthreads();
    const double inv = 1.0 / s[j][j];
    for (int i = j + 1 + tid; i < nb; i += 256) s[i][j] *= inv;
    __syncthreads();
    // trailing update of the lower triangle: s[i][k] -= s[i][j] * s[k][j],  j < k <= i < nb
    const int rem = nb - j - 1;
    for (int t = tid; t < rem * rem; t += 256) {
      const int i = j + 1 + t / rem, k = j + 1 + t % rem;
      if (k <= i) s[i][k] = fma(-s[i][j], s[k][j], s[i][k]);
    }
    __syncthreads();
  }
  // inverse by rows: X[i][c] = (delta_ic - sum_{k=c}^{i-1} L[i][k] X[k][c]) / L[i][i]; thread c owns column c
  for (int i = 0; i < nb; ++i) {
    if (tid <= i) {
      const int c = tid;
      double a0 = 0.0, a1 = 0.0;
      int k = c;
      for (; k + 1 < i; k += 2) {
        a0 = fma(s[i][k], x[k][c], a0);
        a1 = fma(s[i][k + 1], x[k + 1][c], a1);
      }
      if (k < i) a0 = fma(s[i][k], x[k][c], a0);
      x[i][c] = (((i == c) ? 1.0 : 0.0) - (a0 + a1)) / s[i][i];
    }
    __syncthreads();
  }
  for (int t = tid; t < nb * nb; t += 256) {
    const int r = t / nb, c = t % nb;
    A[(long)r * lda + c] = (c <= r) ? s[r][c] : 0.0;
    Pd[(long)r * ldp + c] = (c <= r) ? x[r][c] : 0.0;
  }
}

// zero the strict upper triangle outside the diagonal blocks (diag blocks are written clean by potrf_diag_kernel)
__global__ void zero_upper_blocks_kernel(int M, double* A, long lda, double* P, long ldp) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= M || c >= M) return;
  if (c / NB > r / NB) {
    if (A) A[(long)r * lda + c] = 0.0;
    if (P) P[(long)r * ldp + c] = 0.0;
  }
}

constexpr int DIAG_SMEM = 2 * NB * LDB * (int)sizeof(double);

int potrf_inv_impl(int M, double* A, long lda, double* P, long ldp, double* work, int* info, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    NPGP_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
    attr_set = true;
  }
  NPGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  for (int j0 = 0; j0 < M; j0 += NB) {
    const int nb = (M - j0 < NB) ? (M - j0) : NB;
    double* A11 = A + (long)j0 * lda + j0;
    double* P11 = P + (long)j0 * ldp + j0;
    potrf_diag_kernel<<<1, 256, DIAG_SMEM, st>>>(nb, A11, lda, P11, ldp, j0, info);
    NPGP_LAUNCH_CHECK();
    const int m2 = M - j0 - nb;
    if (m2 <= 0) break;
    double* A21 = A + (long)(j0 + nb) * lda + j0;
    double* A22 = A + (long)(j0 + nb) * lda + (j0 + nb);
    // L21 = A21 * P11^T   (in place: each CTA reads and writes only its own 128 rows, N = nb <= 128 is one tile)
    int rc = dgemm_impl(0, 1, m2, nb, nb, 1.0, A21, lda, P11, ldp, 0.0, A21, lda, 0, 0, 0, st);
    if (rc) return rc;
    // A22 -= L21 L21^T  (lower tiles only)
    rc = dgemm_impl(0, 1, m2, m2, nb, -1.0, A21, lda, A21, lda, 1.0, A22, lda, 0, 0, 1, st);
    if (rc) return rc;
  }
  {
    dim3 blk(32, 8), grd(ceil_div(M, 32), ceil_div(M, 8));
    zero_upper_blocks_kernel<<<grd, blk, 0, st>>>(M, A, lda, P, ldp);
    NPGP_LAUNCH_CHECK();
  }
  // assemble P = L^-1 by recursive doubling; `work` holds T = L21 P11 (at most (M/2)^2 doubles, ld = b)
  for (int b = NB; b < M; b *= 2) {
    for (int s = 0; s + b < M; s += 2 * b) {
      const int b2 = (M - s - b < b) ? (M - s - b) : b;
      const double* L21 = A + (long)(s + b) * lda + s;
      const double* P11 = P + (long)s * ldp + s;
      const double* P22 = P + (long)(s + b) * ldp + (s + b);
      double* P21 = P + (long)(s + b) * ldp + s;
      int rc = dgemm_impl(0, 0, b2, b, b, 1.0, L21, lda, P11, ldp, 0.0, work, b, 0, 1, 0, st);
      if (rc) return rc;
      rc = dgemm_impl(0, 0, b2, b, b2, -1.0, P22, ldp, work, b, 0.0, P21, ldp, 1, 0, 0, st);
      if (rc) return rc;
    }
  }
  return NPGP_OK;
}

}  // namespace npgp

using namespace npgp;

extern "C" long npgp_potrf_workspace_bytes(int M) {
  long h = (M + 1) / 2 + NB;
  return h * h * (long)sizeof(double);
}

extern "C" int npgp_potrf_inv_lower(int M, double* A, long lda, double* P, long ldp, void* work, long work_bytes,
                                    int* info, cudaStream_t stream) {
  if (M < 0) return NPGP_EINVAL;
  if (M == 0) return NPGP_OK;
  if (!A || !P || !info || !work) return NPGP_EINVAL;
  if ((lda & 1) || (ldp & 1)) return NPGP_EUNSUPPORTED;
  if (work_bytes < npgp_potrf_workspace_bytes(M)) return NPGP_EWORKSPACE;
  return potrf_inv_impl(M, A, lda, P, ldp, static_cast<double*>(work), info, stream);
}
