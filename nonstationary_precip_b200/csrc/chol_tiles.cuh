// 64 x 64 FP64 tile primitives of the blocked Cholesky kernels (chol.cu, chol_flow.cu): tile loads / stores, DMMA tile
// products, and the in-shared-memory factor-and-invert of a diagonal block.
#pragma once
#include "common.cuh"

namespace npgp {

constexpr int TB = 64;    // tile edge
constexpr int TLD = 68;   // shared-memory leading dimension (= 4 mod 16: conflict-free fragment loads)
constexpr int CT = 256;   // threads per CTA (8 warps, warp tile 16 x 32)
constexpr int TILE_SMEM = TB * TLD;  // doubles

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct WarpPos {
  int wm0, wn0, g, t4;
  __device__ WarpPos() {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    wm0 = (w >> 1) * 16;
    wn0 = (w & 1) * 32;
    g = lane >> 2;
    t4 = lane & 3;
  }
};

// global (rows r0.., cols c0.. of an M x M matrix) -> shared 64x64 tile; out-of-range entries are 0 (1 on the diagonal
// when `ident`), so ragged edges behave like an identity-padded matrix
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ G, long ld, int r0, int c0, int M,
                                          bool ident) {
  for (int e = threadIdx.x; e < TB * TB / 2; e += CT) {
    const int r = e >> 5, c = (e & 31) * 2;
    const int gr = r0 + r, gc = c0 + c;
    double v0 = 0.0, v1 = 0.0;
    if (gr < M) {
      if (gc + 1 < M) {
        const double2 t = *reinterpret_cast<const double2*>(G + (long)gr * ld + gc);
        v0 = t.x;
        v1 = t.y;
      } else if (gc < M) {
        v0 = G[(long)gr * ld + gc];
      }
    }
    if (ident) {
      if (gr >= M && gr == gc) v0 = 1.0;
      if (gr >= M && gr == gc + 1) v1 = 1.0;
    }
    s[r * TLD + c] = v0;
    s[r * TLD + c + 1] = v1;
  }
}

__device__ __forceinline__ void store_tile(const double* s, double* __restrict__ G, long ld, int r0, int c0, int M) {
  for (int e = threadIdx.x; e < TB * TB; e += CT) {
    const int r = e >> 6, c = e & 63;
    if (r0 + r < M && c0 + c < M) G[(long)(r0 + r) * ld + c0 + c] = s[r * TLD + c];
  }
}

// acc += op(A) op(B) for 64x64x64 tiles in shared memory.  AT: A stored [k][m]; BT: B stored [n][k].
template <bool AT, bool BT>
__device__ __forceinline__ void mma_64(double (&acc)[2][4][2], const double* sA, const double* sB, const WarpPos& p) {
#pragma unroll 4
  for (int kk = 0; kk < TB; kk += 4) {
    double a[2], b[4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int m = p.wm0 + mt * 8 + p.g;
      a[mt] = AT ? sA[(kk + p.t4) * TLD + m] : sA[m * TLD + kk + p.t4];
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int n = p.wn0 + nt * 8 + p.g;
      b[nt] = BT ? sB[n * TLD + kk + p.t4] : sB[(kk + p.t4) * TLD + n];
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}

__device__ __forceinline__ void acc_zero(double (&acc)[2][4][2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

__device__ __forceinline__ void acc_to_smem(const double (&acc)[2][4][2], double* s, const WarpPos& p, double scale) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int r = p.wm0 + mt * 8 + p.g, c = p.wn0 + nt * 8 + 2 * p.t4;
      s[r * TLD + c] = scale * acc[mt][nt][0];
      s[r * TLD + c + 1] = scale * acc[mt][nt][1];
    }
}

// out = base(global tile) + scale * acc, written back to global (guarded)
__device__ __forceinline__ void acc_axpy_global(const double (&acc)[2][4][2], double* __restrict__ G, long ld, int r0,
                                                int c0, int M, const WarpPos& p, double scale, bool add_base) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int r = r0 + p.wm0 + mt * 8 + p.g, c = c0 + p.wn0 + nt * 8 + 2 * p.t4;
      if (r < M) {
        double* q = G + (long)r * ld + c;
        if (c < M) q[0] = (add_base ? q[0] : 0.0) + scale * acc[mt][nt][0];
        if (c + 1 < M) q[1] = (add_base ? q[1] : 0.0) + scale * acc[mt][nt][1];
      }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 64 x 64 diagonal block (symmetric input in s, lower part used): Cholesky factor in place (upper zeroed) and the
// inverse of the factor in x.  All 256 threads.  Blocked by 8 so that every loop body is small and re-executed (a fully
// unrolled 64-column elimination is instruction-fetch bound): the 8x8 pivot block is factored and inverted serially in
// the registers of one thread (the critical path is the 64 dependent rsqrt's anyway), the panel below it is one row
// per thread, the trailing update one element per thread.  The inverse is then assembled by recursive doubling
// (8 -> 16 -> 32 -> 64) with element-per-thread products.  tmp: >= 32*32 doubles.
// ---------------------------------------------------------------------------------------------------------------------
// 8x8 pivot block, one thread, registers.  (Measured alternatives, both slower on B200: a float-seeded rsqrt with two
// Newton steps, 2330 vs 1660 cycles, and a square-root-free elimination that keeps only a reciprocal on the pivot chain,
// 1880 cycles -- the block is bound by the ~400 instructions one thread has to issue, not by the rsqrt latency.  Round 2:
// the same block on 8 lanes of a warp (row per lane, shuffles for the pivot column, ~150 warp instructions) is slower too:
// the dependent shuffles cost more than the issue slots they save, potrf at M = 1024 went 0.460 -> 0.482 ms.)
__device__ __forceinline__ void chol8_serial(double* s, double* sInv, int c0, int global_offset,
                                             int* __restrict__ info) {
  double a[8][8], inv[8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) a[r][c] = s[(c0 + r) * TLD + c0 + c];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double dj = a[j][j];
    if (!(dj > 0.0)) atomicCAS(info, 0, global_offset + c0 + j + 1);
    inv[j] = rsqrt(dj);
    a[j][j] = dj * inv[j];
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[i][j] *= inv[j];
#pragma unroll
    for (int i = j + 1; i < 8; ++i)
#pragma unroll
      for (int k = j + 1; k <= i; ++k) a[i][k] = fma(-a[i][j], a[k][j], a[i][k]);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    sInv[c0 + r] = inv[r];
#pragma unroll
    for (int c = 0; c < 8; ++c) s[(c0 + r) * TLD + c0 + c] = (c <= r) ? a[r][c] : 0.0;
  }
}

// column c (0..7) of the inverse of the 8x8 lower block at (c0,c0), by forward substitution; one thread per column
__device__ __forceinline__ void inv8_column(const double* s, const double* sInv, double* x, int c0, int c) {
  double xc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    double acc = (r == c) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < r; ++k) acc = fma(-s[(c0 + r) * TLD + c0 + k], (k >= c) ? xc[k] : 0.0, acc);
    xc[r] = (r >= c) ? acc * sInv[c0 + r] : 0.0;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) x[(c0 + r) * TLD + c0 + c] = xc[r];
}

// rank-8 update of one 8x8 tile on the tensor pipe: s[ri.., ck..] -= Lr Lc^T with Lr = s[ri.., c0..c0+7], Lc = s[ck.., c0..c0+7]
__device__ __forceinline__ void rank8_tile_update(double* s, int ri, int ck, int c0, int g, int t4) {
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
  for (int kk = 0; kk < 8; kk += 4)
    dmma884(acc0, acc1, s[(ri + g) * TLD + c0 + kk + t4], s[(ck + g) * TLD + c0 + kk + t4]);
  double* q = s + (ri + g) * TLD + ck + 2 * t4;
  q[0] -= acc0;
  q[1] -= acc1;
}

static __device__ void factor_invert_64(double* s, double* x, double* tmp, int global_offset, int* __restrict__ info) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  constexpr int NWARP = CT / 32;
  double* sInv = tmp + 32 * 32;  // 64 reciprocal pivots (tmp holds >= 64*68 doubles)
  for (int e = tid; e < TB * TB; e += CT) x[(e >> 6) * TLD + (e & 63)] = 0.0;
  if (tid == 0) chol8_serial(s, sInv, 0, global_offset, info);
  __syncthreads();
  // Two barriers per 8 columns.  Phase A: panel rows by forward substitution against the 8x8 pivot factor, one row per
  // thread.  Phase B: rank-8 trailing update, one 8x8 tile per warp and DMMA pair; warp 0 takes the NEXT pivot tile first
  // and one of its threads factors it at once (look-ahead), so the serial 8-pivot chain (the critical path: 64
  // dependent rsqrt's) runs under the other warps' tiles and under the inversion of the current pivot block (warp 7).
  for (int c0 = 0; c0 < TB; c0 += 8) {
    const int r1 = c0 + 8, n = TB - r1, nt = n >> 3;
    if (tid < n) {
      const int r = r1 + tid;
      double l[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        double acc = s[r * TLD + c0 + c];
#pragma unroll
        for (int k = 0; k < c; ++k) acc = fma(-l[k], s[(c0 + c) * TLD + c0 + k], acc);
        l[c] = acc * sInv[c0 + c];
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        s[r * TLD + c0 + c] = l[c];
        s[(c0 + c) * TLD + r] = 0.0;  // zero the mirrored (upper) entries
      }
    }
    __syncthreads();
    if (warp == 0) {
      if (nt > 0) {
        rank8_tile_update(s, r1, r1, c0, g, t4);
        __syncwarp();
        if (lane == 0) chol8_serial(s, sInv, r1, global_offset, info);
      }
    } else {
      if (warp == NWARP - 1 && lane < 8) inv8_column(s, sInv, x, c0, lane);
      __syncwarp();
      // lower-triangular tiles (ti >= tk) of the trailing block except (0,0), round-robin over warps 1..NWARP-1
      const int ntiles = nt * (nt + 1) / 2;
      for (int e = warp; e < ntiles; e += NWARP - 1) {
        int ti = 0;
        while ((ti + 1) * (ti + 2) / 2 <= e) ++ti;
        const int tk = e - ti * (ti + 1) / 2;
        rank8_tile_update(s, r1 + 8 * ti, r1 + 8 * tk, c0, g, t4);
      }
    }
    __syncthreads();
  }
  // inverse by doubling on the tensor pipe: for block pairs of size sz, T = L21 X11, then X21 = -X22 T (8x8 output
  // tiles round-robin over the warps; structurally zero k-ranges skipped)
  for (int sz = 8; sz < TB; sz *= 2) {
    const int tps = sz >> 3, tiles = (TB / (2 * sz)) * tps * tps;
    for (int e = warp; e < tiles; e += NWARP) {
      const int pr = e / (tps * tps), q = e % (tps * tps), tr = q / tps, tc = q % tps;
      const int b0 = pr * 2 * sz;
      double acc0 = 0.0, acc1 = 0.0;
      for (int kk = 8 * tc; kk < sz; kk += 4)
        dmma884(acc0, acc1, s[(b0 + sz + 8 * tr + g) * TLD + b0 + kk + t4], x[(b0 + kk + t4) * TLD + b0 + 8 * tc + g]);
      double* q2 = tmp + pr * sz * sz + (8 * tr + g) * sz + 8 * tc + 2 * t4;
      q2[0] = acc0;
      q2[1] = acc1;
    }
    __syncthreads();
    for (int e = warp; e < tiles; e += NWARP) {
      const int pr = e / (tps * tps), q = e % (tps * tps), tr = q / tps, tc = q % tps;
      const int b0 = pr * 2 * sz;
      double acc0 = 0.0, acc1 = 0.0;
      for (int kk = 0; kk < 8 * (tr + 1); kk += 4)
        dmma884(acc0, acc1, x[(b0 + sz + 8 * tr + g) * TLD + b0 + sz + kk + t4],
                tmp[pr * sz * sz + (kk + t4) * sz + 8 * tc + g]);
      double* q2 = x + (b0 + sz + 8 * tr + g) * TLD + b0 + 8 * tc + 2 * t4;
      q2[0] = -acc0;
      q2[1] = -acc1;
    }
    __syncthreads();
  }
}


}  // namespace npgp
