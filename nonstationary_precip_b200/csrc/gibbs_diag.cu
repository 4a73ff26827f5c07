// Fused diagonal-Gibbs cross-covariance tiles: forward K(X,Z) and analytic backward.
// Replaces GibbsKernel.forward (reference models/gibbs_kernels.py:135-162) and its autograd graph.
//
// Tiling: a CTA of 128 threads owns 256 consecutive columns (2 per thread -> 16-byte coalesced stores, a warp writes
// 512 contiguous bytes of a row) and walks down rows.  Per-column data (z, l^2, c) lives in registers, per-row data is
// staged in shared memory 32 rows at a time and read by broadcast.  The kernel is store-only in the forward direction
// (8 B / pair) and FP64-pipe co-limited (about 40 DP instructions per pair at D=3).
#include <cstdlib>
#include "common.cuh"
#include "pairmath.cuh"

namespace npgp {


template <int D>
__device__ __forceinline__ void stage_rows_diag(int n1, int i0, const double* __restrict__ x1,
                                                const double* __restrict__ ell1, double (*sx)[D], double (*sa)[D],
                                                double (*sl)[D], double* sc) {
  for (int r = threadIdx.x; r < kTI; r += kNT) {
    const int i = i0 + r;
    double prod = 1.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double xv = (i < n1) ? x1[(long)i * D + d] : 0.0;
      const double lv = (i < n1) ? ell1[(long)d * n1 + i] : 1.0;
      sx[r][d] = xv;
      sa[r][d] = lv * lv;
      if (sl) sl[r][d] = lv;
      prod *= 1.4142135623730951 * lv;  // 2^(D/2) prod l
    }
    sc[r] = sqrt(prod);
  }
}

template <int D, bool HAS_U>
__global__ void __launch_bounds__(kNT) gibbs_diag_fwd_kernel(int n1, int n2, const double* __restrict__ x1,
                                                             const double* __restrict__ ell1,
                                                             const double* __restrict__ x2,
                                                             const double* __restrict__ ell2,
                                                             const double* __restrict__ scale, double* __restrict__ K,
                                                             long ldk, int vec_ok, const double* __restrict__ u,
                                                             double* __restrict__ Ku, int rows_per_cta) {
  __shared__ double sx[kTI][D], sa[kTI][D], sc[kTI];
  __shared__ double sexp[256];
  load_exp_table(sexp);
  const int jbase = blockIdx.y * kTJ + threadIdx.x * kCPT;
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(n1, row_begin + rows_per_cta);
  const double s = scale ? *scale : 1.0;

  double z[kCPT][D], b[kCPT][D], cj[kCPT], uj[kCPT];
  bool valid[kCPT];
#pragma unroll
  for (int c = 0; c < kCPT; ++c) {
    const int j = jbase + c;
    valid[c] = j < n2;
    double prod = 1.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      z[c][d] = valid[c] ? x2[(long)j * D + d] : 0.0;
      const double lv = valid[c] ? ell2[(long)d * n2 + j] : 1.0;
      b[c][d] = lv * lv;
      prod *= 1.4142135623730951 * lv;
    }
    cj[c] = sqrt(prod) * s;  // outputscale folded into the column constant
    uj[c] = (HAS_U && valid[c]) ? u[j] : 0.0;
  }

  for (int i0 = row_begin; i0 < row_end; i0 += kTI) {
    __syncthreads();
    stage_rows_diag<D>(n1, i0, x1, ell1, sx, sa, (double(*)[D]) nullptr, sc);
    __syncthreads();
    const int nr = min(kTI, row_end - i0);
#pragma unroll 2
    for (int r = 0; r < nr; ++r) {
      const double k0 = gibbs_diag_eval<D>(sx[r], sa[r], sc[r], z[0], b[0], cj[0], sexp);
      const double k1 = gibbs_diag_eval<D>(sx[r], sa[r], sc[r], z[1], b[1], cj[1], sexp);
      double* krow = K + (long)(i0 + r) * ldk + jbase;
      if (vec_ok && valid[1]) {
        st_v2(krow, k0, k1);
      } else {
        if (valid[0]) krow[0] = k0;
        if (valid[1]) krow[1] = k1;
      }
      if (HAS_U) {
        double p = warp_sum(fma(k0, uj[0], k1 * uj[1]));
        if ((threadIdx.x & 31) == 0) atomicAdd(&Ku[i0 + r], p);
      }
    }
  }
}


template <int D, bool DX1, bool DX2, int CPT>
__global__ void __launch_bounds__(kNT) gibbs_diag_bwd_kernel(int n1, int n2, const double* __restrict__ x1,
                                                             const double* __restrict__ ell1,
                                                             const double* __restrict__ x2,
                                                             const double* __restrict__ ell2,
                                                             const double* __restrict__ scale, GSpec g, int vec_ok,
                                                             double* __restrict__ d_ell1, double* __restrict__ d_x1,
                                                             double* __restrict__ d_ell2, double* __restrict__ d_x2,
                                                             double* __restrict__ d_scale, int rows_per_cta) {
  constexpr int NRC = D + 1 + (DX1 ? D : 0);  // row-side components: W_d, S0, [XZ_d]
  constexpr int NW = kNT / 32;
  __shared__ double sx[kTI][D], sa[kTI][D], sl[kTI][D], sc[kTI], srs[kTI], srv[kTI];
  __shared__ double part[NW][kTI][NRC];
  __shared__ double red[32];
  __shared__ double panels[NW * RowReducer<NRC>::PANEL];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RowReducer<NRC> rr(panels, warp, lane);
  __shared__ double sexp[256];
  load_exp_table(sexp);
  const int jbase = blockIdx.y * (kNT * CPT) + threadIdx.x * CPT;
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(n1, row_begin + rows_per_cta);
  const double s = scale ? *scale : 1.0;

  double z[CPT][D], b[CPT][D], cj[CPT], cv[CPT];
  bool valid[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int j = jbase + c;
    valid[c] = j < n2;
    double prod = 1.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      z[c][d] = valid[c] ? x2[(long)j * D + d] : 0.0;
      const double lv = valid[c] ? ell2[(long)d * n2 + j] : 1.0;
      b[c][d] = lv * lv;
      prod *= 1.4142135623730951 * lv;
    }
    cj[c] = sqrt(prod);
    cv[c] = (g.colvec && valid[c]) ? g.colvec[j] : 0.0;
  }
  double cw[CPT][D], cs0[CPT], cxz[CPT][D];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    cs0[c] = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) cw[c][d] = cxz[c][d] = 0.0;
  }
  double acc_scale = 0.0;
  GPrefetch<CPT> gq(g, jbase, valid, vec_ok, row_begin, row_end);

  for (int i0 = row_begin; i0 < row_end; i0 += kTI) {
    __syncthreads();
    stage_rows_diag<D>(n1, i0, x1, ell1, sx, sa, sl, sc);
    for (int r = threadIdx.x; r < kTI; r += kNT) {
      const int i = i0 + r;
      srs[r] = (g.rowscale && i < n1) ? g.rowscale[i] : 1.0;
      srv[r] = (g.rowvec && i < n1) ? g.rowvec[i] : 0.0;
    }
    __syncthreads();
    const int nr = min(kTI, row_end - i0);
    for (int r = 0; r < nr; ++r) {
      double gv[CPT];
      gq.pop(i0 + r, gv);
#pragma unroll
      for (int c = 0; c < CPT; ++c) gv[c] = fma(srv[r], cv[c], gv[c] * srs[r]);
      double rw[D], rxz[D], rs0 = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) rw[d] = rxz[d] = 0.0;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        DiagPair<D> p;
        gibbs_diag_eval<D>(sx[r], sa[r], sc[r], z[c], b[c], cj[c], sexp, &p);
        const double gk0 = valid[c] ? gv[c] * p.k : 0.0;  // dL/ds contribution (unscaled kernel)
        acc_scale += gk0;
        const double gk = gk0 * s;
        rs0 += gk;
        cs0[c] += gk;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const double t = p.dl[d] * p.is[d];
          const double gis = gk * p.is[d];
          const double w = gis * fma(2.0 * p.dl[d], t, -1.0);
          rw[d] += w;
          cw[c][d] += w;
          if (DX1 || DX2) {
            const double xz = 2.0 * gk * t;
            if (DX1) rxz[d] += xz;
            if (DX2) cxz[c][d] += xz;
          }
        }
      }
      // row-side: reduce over the warp's 64 columns (batched shared-memory reduction), one partial per warp
#pragma unroll
      for (int d = 0; d < D; ++d) rr.put(r, d, rw[d]);
      rr.put(r, D, rs0);
      if (DX1) {
#pragma unroll
        for (int d = 0; d < D; ++d) rr.put(r, D + 1 + d, rxz[d]);
      }
      rr.flush_if_due(r, nr, [&](int row, int comp, double v) { part[warp][row][comp] = v; });
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nr * NRC; t += kNT) {
      const int r = t / NRC, comp = t % NRC;
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) v += part[w][r][comp];
      const int i = i0 + r;
      if (comp < D) {
        atomicAdd(&d_ell1[(long)comp * n1 + i], sl[r][comp] * v);
      } else if (comp == D) {
#pragma unroll
        for (int d = 0; d < D; ++d) atomicAdd(&d_ell1[(long)d * n1 + i], v / (2.0 * sl[r][d]));
      } else if (DX1) {
        atomicAdd(&d_x1[(long)i * D + (comp - D - 1)], -v);
      }
    }
  }
  // column-side: one atomic per (column, component) per CTA
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (!valid[c]) continue;
    const int j = jbase + c;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double lv = sqrt(b[c][d]);
      atomicAdd(&d_ell2[(long)d * n2 + j], fma(lv, cw[c][d], cs0[c] / (2.0 * lv)));
      if (DX2) atomicAdd(&d_x2[(long)j * D + d], cxz[c][d]);
    }
  }
  if (d_scale) {
    const double t = block_sum(acc_scale, red);
    if (threadIdx.x == 0) atomicAdd(d_scale, t);
  }
}


template <int D>
static int launch_fwd(int n1, int n2, const double* x1, const double* ell1, const double* x2, const double* ell2,
                      const double* scale, double* K, long ldk, const double* u, double* Ku, cudaStream_t st) {
  const int col_tiles = ceil_div(n2, kTJ);
  const int rpc = pick_rows_per_cta(n1, col_tiles, 16);
  dim3 grid(ceil_div(n1, rpc), col_tiles);
  const int vec_ok = (ldk % 2 == 0) && ((reinterpret_cast<uintptr_t>(K) & 15) == 0);
  if (u)
    gibbs_diag_fwd_kernel<D, true><<<grid, kNT, 0, st>>>(n1, n2, x1, ell1, x2, ell2, scale, K, ldk, vec_ok, u, Ku, rpc);
  else
    gibbs_diag_fwd_kernel<D, false><<<grid, kNT, 0, st>>>(n1, n2, x1, ell1, x2, ell2, scale, K, ldk, vec_ok, u, Ku, rpc);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

template <int D, int CPT>
static int launch_bwd_cpt(int n1, int n2, const double* x1, const double* ell1, const double* x2, const double* ell2,
                          const double* scale, GSpec g, double* d_ell1, double* d_x1, double* d_ell2, double* d_x2,
                          double* d_scale, cudaStream_t st) {
  const int col_tiles = ceil_div(n2, kNT * CPT);
  const int rpc = pick_rows_per_cta(n1, col_tiles, 8);
  dim3 grid(ceil_div(n1, rpc), col_tiles);
  const int vec_ok = g.Gm ? ((g.ldg % 2 == 0) && ((reinterpret_cast<uintptr_t>(g.Gm) & 15) == 0)) : 0;
#define NPGP_L(A, B)                                                                                             \
  gibbs_diag_bwd_kernel<D, A, B, CPT><<<grid, kNT, 0, st>>>(n1, n2, x1, ell1, x2, ell2, scale, g, vec_ok, d_ell1, \
                                                            d_x1, d_ell2, d_x2, d_scale, rpc)
  if (d_x1 && d_x2) NPGP_L(true, true);
  else if (d_x1) NPGP_L(true, false);
  else if (d_x2) NPGP_L(false, true);
  else NPGP_L(false, false);
#undef NPGP_L
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

template <int D>
static int launch_bwd(int n1, int n2, const double* x1, const double* ell1, const double* x2, const double* ell2,
                      const double* scale, GSpec g, double* d_ell1, double* d_x1, double* d_ell2, double* d_x2,
                      double* d_scale, cudaStream_t st) {
  // 2 columns per thread: 0.535 ms at B=65536, M=1024, D=3 against 0.61 ms for 1
  return launch_bwd_cpt<D, 2>(n1, n2, x1, ell1, x2, ell2, scale, g, d_ell1, d_x1, d_ell2, d_x2, d_scale, st);
}

}  // namespace npgp

using namespace npgp;

#define NPGP_DISPATCH_D(D, CALL)            \
  switch (D) {                              \
    case 1: { constexpr int DD = 1; CALL; } \
    case 2: { constexpr int DD = 2; CALL; } \
    case 3: { constexpr int DD = 3; CALL; } \
    case 4: { constexpr int DD = 4; CALL; } \
    case 5: { constexpr int DD = 5; CALL; } \
    case 6: { constexpr int DD = 6; CALL; } \
    default: return NPGP_EUNSUPPORTED;      \
  }

extern "C" int npgp_gibbs_diag_fwd(int D, int n1, int n2, const double* x1, const double* ell1, const double* x2,
                                   const double* ell2, const double* scale, double* K, long ldk, const double* u,
                                   double* Ku, cudaStream_t stream) {
  if (n1 < 0 || n2 < 0 || !K || ldk < n2 || (u && !Ku)) return NPGP_EINVAL;
  if (n1 == 0 || n2 == 0) return NPGP_OK;
  if (!x1 || !ell1 || !x2 || !ell2) return NPGP_EINVAL;
  NPGP_DISPATCH_D(D, return launch_fwd<DD>(n1, n2, x1, ell1, x2, ell2, scale, K, ldk, u, Ku, stream));
}

extern "C" int npgp_gibbs_diag_bwd(int D, int n1, int n2, const double* x1, const double* ell1, const double* x2,
                                   const double* ell2, const double* scale, const double* G, long ldg,
                                   const double* rowscale, const double* rowvec, const double* colvec, double* d_ell1,
                                   double* d_x1, double* d_ell2, double* d_x2, double* d_scale, cudaStream_t stream) {
  if (n1 < 0 || n2 < 0 || (!G && !rowvec) || (G && ldg < n2) || ((rowvec != nullptr) != (colvec != nullptr)))
    return NPGP_EINVAL;
  if (n1 == 0 || n2 == 0) return NPGP_OK;
  if (!x1 || !ell1 || !x2 || !ell2 || !d_ell1 || !d_ell2) return NPGP_EINVAL;
  GSpec g{G, ldg, rowscale, rowvec, colvec};
  NPGP_DISPATCH_D(D, return launch_bwd<DD>(n1, n2, x1, ell1, x2, ell2, scale, g, d_ell1, d_x1, d_ell2, d_x2, d_scale,
                                           stream));
}
