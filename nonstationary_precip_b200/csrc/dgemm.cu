// FP64 tensor-core contractions (DMMA.8x8x4 through mma.sync.m8n8k4.f64 -- tcgen05 has no FP64 kind on sm_100a).
//
//   npgp_dgemm    C = alpha op(A) op(B) + beta C            (row-major, triangular-operand / triangular-output skipping)
//   npgp_rowquad  T = K C,  q_i = sum_j T_ij K_ij           (whitened-SVGP predictive variance term, fused row-dot)
//   npgp_wsyrk    Out = K^T diag(w) K                       (dL/dC of the SVGP ELBO, SGPR Phi = Kzx Kxz), split over rows
//
// These replace the cuBLAS/MAGMA calls GPyTorch issues for  A = L^-1 Kzx,  A^T (S - I) A  and  Kzx Kxz
// (reference models/gibbs_kernels.py:222-232 via LowRankRootLazyTensor; models/dgps.py:25-35 via VariationalStrategy).
//
// Tiling: 128x128 CTA tile, BK = 16, 3-stage cp.async pipeline, 8 warps each owning a 64x32 warp tile = 8x4 DMMA tiles
// (64 accumulator doubles per thread).  Shared-memory tiles are stored either K-contiguous [128][16+4] or MN-contiguous
// [16][128+4], whichever matches the operand's global layout; with a leading dimension = 4 (mod 16) doubles both
// fragment-load patterns are bank-conflict free for 64-bit accesses.
#include "common.cuh"

namespace npgp {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3, GEMM_THREADS = 256;
constexpr int LDS_K = BK + 4;     // K-contiguous tile: [128][20]
constexpr int LDS_MN = 128 + 4;   // MN-contiguous tile: [16][132]
constexpr int TILE_DOUBLES = (128 * LDS_K > BK * LDS_MN) ? 128 * LDS_K : BK * LDS_MN;  // 2560
constexpr int STAGE_DOUBLES = 2 * TILE_DOUBLES + BK;  // A tile, B tile, weights
constexpr int GEMM_SMEM = STAGES * STAGE_DOUBLES * 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// K-contiguous operand: global G[row][k], leading dim ld.  Tile rows [row0,row0+128), k in [k0,k0+16).
__device__ __forceinline__ void load_tile_kcontig(double* s, const double* __restrict__ G, long ld, int row0, int nrows,
                                                  int k0, int kend) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int c = threadIdx.x + GEMM_THREADS * it;
    const int r = c >> 3, kc = (c & 7) * 2;
    const int row = row0 + r, k = k0 + kc;
    int bytes = 0;
    if (row < nrows) bytes = max(0, min(2, kend - k)) * 8;
    const double* src = bytes ? (G + (long)row * ld + k) : G;
    cp_async16(s + r * LDS_K + kc, src, bytes);
  }
}

// MN-contiguous operand: global G[k][col], leading dim ld.  Tile k in [k0,k0+16), cols [col0,col0+128).
__device__ __forceinline__ void load_tile_mncontig(double* s, const double* __restrict__ G, long ld, int col0,
                                                   int ncols, int k0, int kend) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int c = threadIdx.x + GEMM_THREADS * it;
    const int kk = c >> 6, cc = (c & 63) * 2;
    const int k = k0 + kk, col = col0 + cc;
    int bytes = 0;
    if (k < kend) bytes = max(0, min(2, ncols - col)) * 8;
    const double* src = bytes ? (G + (long)k * ld + col) : G;
    cp_async16(s + kk * LDS_MN + cc, src, bytes);
  }
}

struct GemmParams {
  int M, N, K;
  const double* A;
  long lda;
  const double* B;
  long ldb;
  double* C;
  long ldc;
  double alpha, beta;
  int tri_a, tri_b;  // structure of op(A) (M x K) / op(B) (K x N): 0 dense, 1 lower (col <= row), 2 upper (col >= row)
  int out_tri;       // 0 all tiles, 1 only tiles touching the lower triangle, 2 only the upper
  int splits;        // split-K factor (gridDim.z); > 1 => atomicAdd epilogue, C must be pre-scaled by the caller
  const double* w;   // optional weights along K (scales op(A)[m][k] by w[k])
  const double* dotK;  // optional: q[row] += sum_col acc[row][col] * dotK[row][col]
  long lddot;
  double* q;
};

template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(GEMM_THREADS, 1) dgemm_kernel(GemmParams p) {
  extern __shared__ __align__(16) double smem[];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (p.out_tri == 1 && n0 > m0 + BM - 1) return;
  if (p.out_tri == 2 && n0 + BN - 1 < m0) return;

  // k-range of this tile: triangular operands, then split-K
  int kb = 0, ke = p.K;
  if (p.tri_a == 1) ke = min(ke, m0 + BM);   // op(A) lower: k <= m
  if (p.tri_a == 2) kb = max(kb, m0);        // op(A) upper: k >= m
  if (p.tri_b == 1) kb = max(kb, n0);        // op(B) lower: k >= n   (row index k, col index n)
  if (p.tri_b == 2) ke = min(ke, n0 + BN);   // op(B) upper: k <= n
  kb = (kb / BK) * BK;
  if (p.splits > 1) {
    const int span = ke - kb;
    int per = (span + p.splits - 1) / p.splits;
    per = ((per + BK - 1) / BK) * BK;
    const int b2 = kb + (int)blockIdx.z * per;
    ke = min(ke, b2 + per);
    kb = b2;
  }
  const int nk = (ke > kb) ? (ke - kb + BK - 1) / BK : 0;
  if (nk == 0 && p.splits > 1) return;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int wm0 = (warp >> 2) * 64, wn0 = (warp & 3) * 32;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_stage = [&](int stage, int k0) {
    double* sA = smem + stage * STAGE_DOUBLES;
    double* sB = sA + TILE_DOUBLES;
    double* sW = sB + TILE_DOUBLES;
    if (A_KCONTIG) load_tile_kcontig(sA, p.A, p.lda, m0, p.M, k0, ke);
    else load_tile_mncontig(sA, p.A, p.lda, m0, p.M, k0, ke);
    if (B_KCONTIG) load_tile_kcontig(sB, p.B, p.ldb, n0, p.N, k0, ke);
    else load_tile_mncontig(sB, p.B, p.ldb, n0, p.N, k0, ke);
    if (p.w && tid < BK) sW[tid] = (k0 + tid < ke) ? p.w[k0 + tid] : 0.0;
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, kb + s * BK);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, kb + nxt * BK);
      cp_async_commit();
    }
    const double* sA = smem + (kt % STAGES) * STAGE_DOUBLES;
    const double* sB = sA + TILE_DOUBLES;
    const double* sW = sB + TILE_DOUBLES;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[8], b[4];
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {
        const int m = wm0 + mt * 8 + g;
        a[mt] = A_KCONTIG ? sA[m * LDS_K + kk + t4] : sA[(kk + t4) * LDS_MN + m];
      }
      if (p.w) {
        const double wk = sW[kk + t4];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) a[mt] *= wk;
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = wn0 + nt * 8 + g;
        b[nt] = B_KCONTIG ? sB[n * LDS_K + kk + t4] : sB[(kk + t4) * LDS_MN + n];
      }
#pragma unroll
      for (int mt = 0; mt < 8; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
  }
  cp_async_wait<0>();

  // epilogue
  const bool vec_c = (p.ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) {
    const int row = m0 + wm0 + mt * 8 + g;
    double rowdot = 0.0;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = n0 + wn0 + nt * 8 + 2 * t4;
      if (row < p.M && col < p.N) {
        const bool two = (col + 1 < p.N);
        double v0 = p.alpha * acc[mt][nt][0], v1 = p.alpha * acc[mt][nt][1];
        double* cp = p.C + (long)row * p.ldc + col;
        if (p.splits > 1) {
          atomicAdd(cp, v0);
          if (two) atomicAdd(cp + 1, v1);
        } else {
          if (p.beta != 0.0) {
            v0 = fma(p.beta, cp[0], v0);
            if (two) v1 = fma(p.beta, cp[1], v1);
          }
          if (two && vec_c) st_v2(cp, v0, v1);
          else {
            cp[0] = v0;
            if (two) cp[1] = v1;
          }
        }
        if (p.dotK) {
          const double* kp = p.dotK + (long)row * p.lddot + col;
          rowdot = fma(v0, kp[0], rowdot);
          if (two) rowdot = fma(v1, kp[1], rowdot);
        }
      }
    }
    if (p.dotK) {
      rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 1);
      rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 2);
      if (t4 == 0 && row < p.M) atomicAdd(&p.q[row], rowdot);
    }
  }
}

__global__ void scale_matrix_kernel(int M, int N, double* C, long ldc, double beta, int out_tri) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * N) return;
  const int r = (int)(idx / N), c = (int)(idx % N);
  (void)out_tri;
  double* p = C + (long)r * ldc + c;
  *p = (beta == 0.0) ? 0.0 : beta * *p;
}

// copy one triangle of a square matrix onto the other (make symmetric)
__global__ void symmetrize_kernel(int M, double* C, long ldc, int from_upper) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= M || c >= M || c <= r) return;  // (r,c) strictly upper
  if (from_upper) C[(long)c * ldc + r] = C[(long)r * ldc + c];
  else C[(long)r * ldc + c] = C[(long)c * ldc + r];
}

static int launch_gemm(bool a_kc, bool b_kc, const GemmParams& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    NPGP_CUDA(cudaFuncSetAttribute(dgemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    NPGP_CUDA(cudaFuncSetAttribute(dgemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    NPGP_CUDA(cudaFuncSetAttribute(dgemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    NPGP_CUDA(cudaFuncSetAttribute(dgemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM), p.splits);
  if (a_kc && b_kc) dgemm_kernel<true, true><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(p);
  else if (a_kc) dgemm_kernel<true, false><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(p);
  else if (b_kc) dgemm_kernel<false, true><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(p);
  else dgemm_kernel<false, false><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(p);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int dgemm_impl(int transA, int transB, int M, int N, int K, double alpha, const double* A, long lda, const double* B,
               long ldb, double beta, double* C, long ldc, int tri_a, int tri_b, int out_tri, cudaStream_t st) {
  if (M < 0 || N < 0 || K < 0) return NPGP_EINVAL;
  if (M == 0 || N == 0) return NPGP_OK;
  if (!C || (K > 0 && (!A || !B))) return NPGP_EINVAL;
  if ((lda & 1) || (ldb & 1) || !aligned16(A) || !aligned16(B)) return NPGP_EUNSUPPORTED;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.alpha = alpha; p.beta = beta; p.tri_a = tri_a; p.tri_b = tri_b; p.out_tri = out_tri; p.splits = 1;
  // small outputs with a long K: split K so that the 148 SMs have work
  const long tiles = (long)ceil_div(M, BM) * ceil_div(N, BN);
  if (tiles * 2 <= kNumSMs && K >= 8 * BK * 4) {
    int s = (int)(kNumSMs / tiles);
    s = min(s, K / (8 * BK));
    s = min(s, 16);
    if (s > 1) p.splits = s;
  }
  if (p.splits > 1) {
    const long tot = (long)M * N;
    scale_matrix_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(M, N, C, ldc, beta, out_tri);
    NPGP_LAUNCH_CHECK();
  }
  return launch_gemm(transA == 0, transB != 0, p, st);
}

}  // namespace npgp

using namespace npgp;

extern "C" int npgp_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double* A, long lda,
                          const double* B, long ldb, double beta, double* C, long ldc, int tri_a, int tri_b,
                          int out_tri, cudaStream_t stream) {
  return dgemm_impl(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri_a, tri_b, out_tri, stream);
}

// T (n x M) = K (n x M) @ C (M x M);  q[i] = sum_j T_ij K_ij   (q must be zeroed by the caller; pass q = NULL to skip)
extern "C" int npgp_rowquad(int n, int M, const double* K, long ldk, const double* C, long ldc, double* T, long ldt,
                            double* q, cudaStream_t stream) {
  if (n < 0 || M < 0) return NPGP_EINVAL;
  if (n == 0 || M == 0) return NPGP_OK;
  if (!K || !C || !T) return NPGP_EINVAL;
  if ((ldk & 1) || (ldc & 1) || !aligned16(K) || !aligned16(C)) return NPGP_EUNSUPPORTED;
  GemmParams p{};
  p.M = n; p.N = M; p.K = M; p.A = K; p.lda = ldk; p.B = C; p.ldb = ldc; p.C = T; p.ldc = ldt;
  p.alpha = 1.0; p.beta = 0.0; p.splits = 1;
  if (q) { p.dotK = K; p.lddot = ldk; p.q = q; }
  return launch_gemm(true, false, p, stream);
}

// Out (M x M, symmetric) = alpha * K^T diag(w) K, K is (n x M); w may be NULL.  Out is overwritten.
extern "C" int npgp_wsyrk(int n, int M, double alpha, const double* K, long ldk, const double* w, double* Out, long ldo,
                          cudaStream_t stream) {
  if (n < 0 || M < 0) return NPGP_EINVAL;
  if (M == 0) return NPGP_OK;
  if (!Out || (n > 0 && !K)) return NPGP_EINVAL;
  if ((ldk & 1) || !aligned16(K)) return NPGP_EUNSUPPORTED;
  GemmParams p{};
  p.M = M; p.N = M; p.K = n; p.A = K; p.lda = ldk; p.B = K; p.ldb = ldk; p.C = Out; p.ldc = ldo;
  p.alpha = alpha; p.beta = 0.0; p.out_tri = 2; p.w = w;
  const int tm = ceil_div(M, BM);
  const long tiles = (long)tm * (tm + 1) / 2;
  int s = (int)max(1L, (long)(2 * kNumSMs) / tiles / 2);
  s = min(s, max(1, n / (16 * BK)));
  p.splits = max(1, min(s, 32));
  if (p.splits > 1) {
    const long tot = (long)M * M;
    scale_matrix_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(M, M, Out, ldo, 0.0, 0);
    NPGP_LAUNCH_CHECK();
  }
  int rc = launch_gemm(false, false, p, stream);
  if (rc) return rc;
  dim3 blk(32, 8), grd(ceil_div(M, 32), ceil_div(M, 8));
  symmetrize_kernel<<<grd, blk, 0, stream>>>(M, Out, ldo, 1);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_symmetrize(int M, double* C, long ldc, int from_upper, cudaStream_t stream) {
  if (M <= 0 || !C) return M == 0 ? NPGP_OK : NPGP_EINVAL;
  dim3 blk(32, 8), grd(ceil_div(M, 32), ceil_div(M, 8));
  symmetrize_kernel<<<grd, blk, 0, stream>>>(M, C, ldc, from_upper);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
