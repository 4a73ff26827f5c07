// Dataflow (tile-task) Cholesky + inverse factor of the M x M inducing covariance in ONE kernel launch.
//
//   npgp_potrf_inv_flow:  A = L L^T in place (strict upper triangle zeroed),  P = L^-1 (lower),  info as LAPACK.
//
// Same contract and same reference lines as npgp_potrf_inv_lower (psd_safe_cholesky + triangular_solve(eye, chol),
// models/gibbs_kernels.py:197-208; the Cholesky of GPyTorch's VariationalStrategy, models/dgps.py:29-33).  That version
// spends 16 dependent kernel launches per factorisation at M = 1024 (one per 64-column panel) plus the launches of the
// recursive-doubling inverse: 0.66 ms, all of it latency.  Here every 64 x 64 tile is owned by one CTA that stays resident
// and waits on per-tile flags in global memory, so a dependent step costs an L2 round trip instead of a kernel boundary,
// and the inverse is assembled WHILE the factorisation proceeds:
//   L-task (i,k), k <= i:  A_ik -= sum_{j<k} L_ij L_kj^T as the L_ij / L_kj become ready;  then
//                          i == k: factor + invert the block in shared memory (chol_tiles.cuh) -> L_kk, P_kk
//                          i >  k: L_ik = A_ik P_kk^T
//   P-task (i,k), k <  i:  P_ik = -P_ii sum_{j=k}^{i-1} L_ij P_jk     (forward substitution by tiles)
// Tasks are numbered so that every dependency points to a LOWER block index (L tasks column by column, then P tasks row by
// row): with the hardware's in-order CTA dispatch a waiting CTA only ever waits for CTAs that are already running or
// finished, so the scheme cannot deadlock even when not all M^2/4096 CTAs are co-resident (at M = 1024: 256 CTAs, two per
// SM, all resident).  Waits are bounded: after ~1 s without progress a CTA raises the abort flag, every waiter falls
// through, and *info = -1.
// Critical path per 64 columns: factor+invert (13 us) -> flag -> L_{k+1,k} (one 64^3 DMMA product) -> flag -> update of
// A_{k+1,k+1} (one product) -> next factor.  All tile products run on the FP64 tensor pipe (DMMA.8x8x4).
#include "chol_tiles.cuh"
#include "common.cuh"

namespace npgp {

constexpr int FLOW_SMEM = 3 * TILE_SMEM * (int)sizeof(double);
constexpr long FLOW_SPIN_LIMIT = 1L << 22;

__device__ __forceinline__ int flow_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void flow_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// all threads call; thread 0 polls the flag (bounded), then the CTA synchronises
__device__ __forceinline__ void flow_wait(const int* flag, int* abort_flag, int* info) {
  if (threadIdx.x == 0) {
    long spins = 0;
    while (flow_ld_acquire(flag) == 0) {
      if (++spins > FLOW_SPIN_LIMIT || (((spins & 255) == 0) && flow_ld_acquire(abort_flag) != 0)) {
        flow_st_release(abort_flag, 1);
        atomicCAS(info, 0, -1);
        break;
      }
      if (spins > 64) __nanosleep(40);
    }
  }
  __syncthreads();
}

// publish: every thread's global stores of this CTA become visible before the flag does
__device__ __forceinline__ void flow_signal(int* flag) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) flow_st_release(flag, 1);
}

// tile written by another CTA: read through L2 (never a stale L1 line)
__device__ __forceinline__ void load_tile_cg(double* s, const double* __restrict__ G, long ld, int r0, int c0, int M) {
  for (int e = threadIdx.x; e < TB * TB / 2; e += CT) {
    const int r = e >> 5, c = (e & 31) * 2;
    const int gr = r0 + r, gc = c0 + c;
    double v0 = 0.0, v1 = 0.0;
    if (gr < M) {
      if (gc + 1 < M) {
        const double2 t = __ldcg(reinterpret_cast<const double2*>(G + (long)gr * ld + gc));
        v0 = t.x;
        v1 = t.y;
      } else if (gc < M) {
        v0 = __ldcg(G + (long)gr * ld + gc);
      }
    }
    s[r * TLD + c] = v0;
    s[r * TLD + c + 1] = v1;
  }
}

__device__ __forceinline__ void zero_tile_global(double* __restrict__ G, long ld, int r0, int c0, int M) {
  for (int e = threadIdx.x; e < TB * TB; e += CT) {
    const int r = r0 + (e >> 6), c = c0 + (e & 63);
    if (r < M && c < M) G[(long)r * ld + c] = 0.0;
  }
}

// full 64 x 64 tile (leading dimension 64 in global memory) <-> shared tile: the hand-over of S_{k+1,k} (below)
__device__ __forceinline__ void store_full(const double* s, double* __restrict__ G) {
  for (int e = threadIdx.x; e < TB * TB; e += CT) G[e] = s[(e >> 6) * TLD + (e & 63)];
}
__device__ __forceinline__ void load_full_cg(double* s, const double* __restrict__ G) {
  for (int e = threadIdx.x; e < TB * TB / 2; e += CT) {
    const double2 t = __ldcg(reinterpret_cast<const double2*>(G) + e);
    const int r = e >> 5, c = (e & 31) * 2;
    s[r * TLD + c] = t.x;
    s[r * TLD + c + 1] = t.y;
  }
}

static inline long flow_flag_ints(long nb) { return 2 * nb * nb + 16 + nb; }
static inline long flow_scratch_offset(long nb) { return (flow_flag_ints(nb) * (long)sizeof(int) + 255) & ~255L; }

// flags: [0, nb^2) L tiles (i * nb + k), [nb^2, 2 nb^2) P tiles, [2 nb^2] abort, [2 nb^2 + 1, + nb) S tiles;
// scratch (after the flags, 256-byte aligned): nb tiles of 64 x 64 doubles.
// Hand-over of the sub-diagonal tile: the task of tile (k+1,k) publishes S = A_{k+1,k} - sum_{j<k} L_{k+1,j} L_{k,j}^T as soon as
// it has it (long before P_kk exists); the task of the NEXT diagonal tile (k+1,k+1) then forms L_{k+1,k} = S P_kk^T itself the
// moment P_kk is published, instead of waiting for the (k+1,k) task to compute, store and flag the same product: one flag hop,
// one global store and one L2 load less on the critical path of every column (same operands, same MMA order: same bits).
// Up to kFlowBatch independent matrices of the same order are factored by ONE launch: CTA x works on matrix x % n, task
// x / n.  Interleaving keeps every matrix's own task order (dependencies still point to lower block indices), and the early
// (critical-path) tasks of ALL matrices are resident from the start -- two concurrent single-matrix launches instead
// queue the second matrix's CTAs behind the first one's resident waiters (0.62 ms for two against 0.43 for one).
constexpr int kFlowBatch = 4;
struct FlowBatch {
  double* A[kFlowBatch];
  double* P[kFlowBatch];
  int* flags[kFlowBatch];
  int* info[kFlowBatch];
  double* S[kFlowBatch];
  int n;
};

__global__ void __launch_bounds__(CT, 2) potrf_flow_kernel(int M, int nb, const FlowBatch fb, long lda, long ldp) {
  const int which = blockIdx.x % fb.n;
  double* __restrict__ A = fb.A[which];
  double* __restrict__ P = fb.P[which];
  int* __restrict__ flags = fb.flags[which];
  int* __restrict__ info = fb.info[which];
  double* __restrict__ Sg = fb.S[which];
  extern __shared__ double sm[];
  double *bufA = sm, *bufB = sm + TILE_SMEM, *bufC = sm + 2 * TILE_SMEM;
  int* Lf = flags;
  int* Pf = flags + nb * nb;
  int* abort_flag = flags + 2 * nb * nb;
  int* Sf = flags + 2 * nb * nb + 1;
  const WarpPos p;
  const int nL = nb * (nb + 1) / 2;
  int t = blockIdx.x / fb.n;
  if (t < nL) {
    // ---- L task: tile (i,k), column-major numbering
    int k = 0;
    while (t >= nb - k) {
      t -= nb - k;
      ++k;
    }
    const int i = k + t;
    // everything that depends on nobody happens before the first wait: the own tile of A goes to shared memory (bufC) and
    // the mirrored, strictly upper tiles of both outputs are zeroed
    load_tile(bufC, A, lda, i * TB, k * TB, M, i == k);
    if (i != k) {
      zero_tile_global(A, lda, k * TB, i * TB, M);
      zero_tile_global(P, ldp, k * TB, i * TB, M);
    }
    double upd[2][4][2];
    acc_zero(upd);
    for (int j = 0; j < k; ++j) {
      if (i == k && j == k - 1) {
        // last update of a diagonal tile: L_{k,k-1} = S P_{k-1,k-1}^T is formed here (see the note above the kernel)
        flow_wait(&Sf[k - 1], abort_flag, info);
        load_full_cg(bufA, Sg + (long)(k - 1) * TB * TB);
        flow_wait(&Lf[(k - 1) * nb + (k - 1)], abort_flag, info);
        load_tile_cg(bufB, P, ldp, (k - 1) * TB, (k - 1) * TB, M);
        __syncthreads();
        double lk[2][4][2];
        acc_zero(lk);
        mma_64<false, true>(lk, bufA, bufB, p);
        __syncthreads();
        acc_to_smem(lk, bufA, p, 1.0);
        __syncthreads();
        mma_64<false, true>(upd, bufA, bufA, p);
        __syncthreads();
        continue;
      }
      flow_wait(&Lf[i * nb + j], abort_flag, info);
      if (i != k) flow_wait(&Lf[k * nb + j], abort_flag, info);
      load_tile_cg(bufA, A, lda, i * TB, j * TB, M);
      if (i != k) load_tile_cg(bufB, A, lda, k * TB, j * TB, M);
      __syncthreads();
      mma_64<false, true>(upd, bufA, (i != k) ? bufB : bufA, p);  // L_ij L_kj^T
      __syncthreads();
    }
    // own tile minus the accumulated update
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int r = p.wm0 + mt * 8 + p.g, c = p.wn0 + nt * 8 + 2 * p.t4;
        bufC[r * TLD + c] -= upd[mt][nt][0];
        bufC[r * TLD + c + 1] -= upd[mt][nt][1];
      }
    __syncthreads();
    if (i == k + 1) {  // publish S for the next diagonal task
      store_full(bufC, Sg + (long)k * TB * TB);
      flow_signal(&Sf[k]);
    }
    if (i == k) {
      factor_invert_64(bufC, bufB, bufA, k * TB, info);
      store_tile(bufB, P, ldp, k * TB, k * TB, M);
      store_tile(bufC, A, lda, k * TB, k * TB, M);
    } else {
      flow_wait(&Lf[k * nb + k], abort_flag, info);
      load_tile_cg(bufB, P, ldp, k * TB, k * TB, M);
      __syncthreads();
      double acc[2][4][2];
      acc_zero(acc);
      mma_64<false, true>(acc, bufC, bufB, p);  // L_ik = A_ik P_kk^T
      acc_axpy_global(acc, A, lda, i * TB, k * TB, M, p, 1.0, false);
    }
    flow_signal(&Lf[i * nb + k]);
  } else {
    // ---- P task: tile (i,k), k < i, numbered row by row
    t -= nL;
    int i = 1;
    while (t >= i) {
      t -= i;
      ++i;
    }
    const int k = t;
    double upd[2][4][2];
    acc_zero(upd);
    for (int j = k; j < i; ++j) {
      flow_wait(&Lf[i * nb + j], abort_flag, info);
      flow_wait((j == k) ? &Lf[k * nb + k] : &Pf[j * nb + k], abort_flag, info);
      load_tile_cg(bufA, A, lda, i * TB, j * TB, M);
      load_tile_cg(bufB, P, ldp, j * TB, k * TB, M);
      __syncthreads();
      mma_64<false, false>(upd, bufA, bufB, p);  // L_ij P_jk
      __syncthreads();
    }
    flow_wait(&Lf[i * nb + i], abort_flag, info);
    load_tile_cg(bufA, P, ldp, i * TB, i * TB, M);
    acc_to_smem(upd, bufB, p, 1.0);
    __syncthreads();
    double acc[2][4][2];
    acc_zero(acc);
    mma_64<false, false>(acc, bufA, bufB, p);
    acc_axpy_global(acc, P, ldp, i * TB, k * TB, M, p, -1.0, false);  // P_ik = -P_ii (sum)
    flow_signal(&Pf[i * nb + k]);
  }
}

}  // namespace npgp

using namespace npgp;

// flags (ints) + nb scratch tiles, nb = ceil(M / 64)
extern "C" long npgp_potrf_flow_workspace_bytes(int M) {
  const long nb = ((long)M + TB - 1) / TB;
  return flow_scratch_offset(nb) + nb * TB * TB * (long)sizeof(double);
}

// A (M x M, symmetric, lower part used) -> L in place (strict upper triangle zeroed), P = L^-1; *info = 0, the 1-based index
// of the first non-positive pivot, or -1 if the dataflow wait timed out.  One kernel launch (+ one memset of the flags).
extern "C" int npgp_potrf_inv_flow(int M, double* A, long lda, double* P, long ldp, void* work, long work_bytes, int* info,
                                   cudaStream_t stream) {
  if (M < 0) return NPGP_EINVAL;
  if (M == 0) return NPGP_OK;
  if (!A || !P || !info || !work) return NPGP_EINVAL;
  if ((lda & 1) || (ldp & 1) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(P) & 15))
    return NPGP_EUNSUPPORTED;
  if (work_bytes < npgp_potrf_flow_workspace_bytes(M)) return NPGP_EWORKSPACE;
  const int nb = (M + TB - 1) / TB;
  NPGP_CUDA(cudaFuncSetAttribute(potrf_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FLOW_SMEM));
  if (reinterpret_cast<uintptr_t>(work) & 15) return NPGP_EUNSUPPORTED;
  NPGP_CUDA(cudaMemsetAsync(work, 0, (size_t)flow_flag_ints(nb) * sizeof(int), stream));  // the flags; the scratch tiles need none
  NPGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int), stream));
  FlowBatch fb;
  fb.n = 1;
  fb.A[0] = A, fb.P[0] = P, fb.flags[0] = static_cast<int*>(work), fb.info[0] = info;
  fb.S[0] = reinterpret_cast<double*>(static_cast<char*>(work) + flow_scratch_offset(nb));
  potrf_flow_kernel<<<nb * nb, CT, FLOW_SMEM, stream>>>(M, nb, fb, lda, ldp);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

// The same for n <= 4 independent matrices of order M in ONE launch (host arrays of device pointers; work[i]: flags of matrix
// i, npgp_potrf_flow_workspace_bytes(M) bytes each; info[i] as above).  All matrices share lda / ldp.
extern "C" int npgp_potrf_inv_flow_batch(int n, int M, double* const* A, long lda, double* const* P, long ldp,
                                         void* const* work, long work_bytes, int* const* info, cudaStream_t stream) {
  if (n < 1 || n > kFlowBatch || M < 0 || !A || !P || !work || !info) return NPGP_EINVAL;
  if (M == 0) return NPGP_OK;
  if ((lda & 1) || (ldp & 1)) return NPGP_EUNSUPPORTED;
  if (work_bytes < npgp_potrf_flow_workspace_bytes(M)) return NPGP_EWORKSPACE;
  FlowBatch fb;
  fb.n = n;
  for (int i = 0; i < n; ++i) {
    if (!A[i] || !P[i] || !work[i] || !info[i]) return NPGP_EINVAL;
    if ((reinterpret_cast<uintptr_t>(A[i]) & 15) || (reinterpret_cast<uintptr_t>(P[i]) & 15) ||
        (reinterpret_cast<uintptr_t>(work[i]) & 15))
      return NPGP_EUNSUPPORTED;
    fb.A[i] = A[i], fb.P[i] = P[i], fb.flags[i] = static_cast<int*>(work[i]), fb.info[i] = info[i];
  }
  const int nb = (M + TB - 1) / TB;
  for (int i = 0; i < n; ++i) fb.S[i] = reinterpret_cast<double*>(static_cast<char*>(work[i]) + flow_scratch_offset(nb));
  NPGP_CUDA(cudaFuncSetAttribute(potrf_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FLOW_SMEM));
  for (int i = 0; i < n; ++i) {
    NPGP_CUDA(cudaMemsetAsync(work[i], 0, (size_t)flow_flag_ints(nb) * sizeof(int), stream));
    NPGP_CUDA(cudaMemsetAsync(info[i], 0, sizeof(int), stream));
  }
  potrf_flow_kernel<<<n * nb * nb, CT, FLOW_SMEM, stream>>>(M, nb, fb, lda, ldp);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
