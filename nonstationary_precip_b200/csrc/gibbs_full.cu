// Fused full-matrix (Paciorek-Schervish) Gibbs cross-covariance tiles, d = 2 or 3, forward and analytic backward, plus
// the per-point Sigma(h) = softplus((h h^T)o(h h^T)) + DoD map and its backward.
// Replaces MultivariateGibbsKernel.forward / SparseMultivariateGibbsKernel.forward
// (reference models/multivariate_gibbs_kernel.py:98-150, models/sparse_multivariate_gibbs_kernel.py:103-154), which
// materialise (n1,n2,d,d) tensors and call batched LU det / inverse on them; here the d x d algebra is closed form
// (two adjugates per pair because the reference puts the 1e-5 jitter on the inverse but not on the determinants).
//
// Same tiling as gibbs_diag.cu: 128 threads x 2 columns, rows staged 32 at a time in shared memory.
// Sigma is passed packed symmetric, (n, d(d+1)/2) row-major: d=2 [00,01,11], d=3 [00,01,02,11,12,22].
// Gradients w.r.t. Sigma are returned in the same packing and hold the entries of the SYMMETRIC matrix dL/dSigma
// (so dL = sum_k G_kk dS_kk + 2 sum_{k<l} G_kl dS_kl).
#include <cstdlib>
#include "common.cuh"
#include "pairmath.cuh"

namespace npgp {


template <int d>
__device__ __forceinline__ void stage_rows_full(int n1, int i0, const double* __restrict__ x1,
                                                const double* __restrict__ S1, double (*sx)[d],
                                                double (*sS)[sym_size(d)], double* sq) {
  constexpr int P = sym_size(d);
  for (int r = threadIdx.x; r < kTI; r += kNT) {
    const int i = i0 + r;
    double S[P];
#pragma unroll
    for (int k = 0; k < d; ++k) sx[r][k] = (i < n1) ? x1[(long)i * d + k] : 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) S[p] = 0.0;
#pragma unroll
    for (int k = 0; k < d; ++k) S[sym_idx(d, k, k)] = 1.0;
    if (i < n1) {
#pragma unroll
      for (int p = 0; p < P; ++p) S[p] = S1[(long)i * P + p];
    }
#pragma unroll
    for (int p = 0; p < P; ++p) sS[r][p] = S[p];
    sq[r] = sqrt(sqrt(sym_det<d>(S)));
  }
}

template <int d>
__device__ __forceinline__ void load_col_full(int n2, int j, const double* __restrict__ x2,
                                              const double* __restrict__ S2, double* z, double* S, double& q,
                                              bool& valid) {
  constexpr int P = sym_size(d);
  valid = j < n2;
#pragma unroll
  for (int k = 0; k < d; ++k) z[k] = valid ? x2[(long)j * d + k] : 0.0;
#pragma unroll
  for (int p = 0; p < P; ++p) S[p] = 0.0;
#pragma unroll
  for (int k = 0; k < d; ++k) S[sym_idx(d, k, k)] = 1.0;
  if (valid) {
#pragma unroll
    for (int p = 0; p < P; ++p) S[p] = S2[(long)j * P + p];
  }
  q = full_col_factor<d>(sym_det<d>(S));  // includes the 2^(d/2) of det(A)^(-1/2)
}

template <int d, bool HAS_U>
__global__ void __launch_bounds__(kNT) gibbs_full_fwd_kernel(int n1, int n2, const double* __restrict__ x1,
                                                             const double* __restrict__ S1,
                                                             const double* __restrict__ x2,
                                                             const double* __restrict__ S2, double jit2,
                                                             const double* __restrict__ scale, double* __restrict__ K,
                                                             long ldk, int vec_ok, const double* __restrict__ u,
                                                             double* __restrict__ Ku, int rows_per_cta) {
  constexpr int P = sym_size(d);
  __shared__ double sx[kTI][d], sS[kTI][P], sq[kTI];
  __shared__ double sexp[256];
  load_exp_table(sexp);
  const int jbase = blockIdx.y * kTJ + threadIdx.x * kCPT;
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(n1, row_begin + rows_per_cta);
  const double s = scale ? *scale : 1.0;
  double z[kCPT][d], Sj[kCPT][P], qj[kCPT], uj[kCPT];
  bool valid[kCPT];
#pragma unroll
  for (int c = 0; c < kCPT; ++c) {
    load_col_full<d>(n2, jbase + c, x2, S2, z[c], Sj[c], qj[c], valid[c]);
    qj[c] *= s;
    uj[c] = (HAS_U && valid[c]) ? u[jbase + c] : 0.0;
  }
  for (int i0 = row_begin; i0 < row_end; i0 += kTI) {
    __syncthreads();
    stage_rows_full<d>(n1, i0, x1, S1, sx, sS, sq);
    __syncthreads();
    const int nr = min(kTI, row_end - i0);
#pragma unroll 2
    for (int r = 0; r < nr; ++r) {
      const double k0 = gibbs_full_eval<d>(sx[r], sS[r], sq[r], z[0], Sj[0], qj[0], jit2, sexp);
      const double k1 = gibbs_full_eval<d>(sx[r], sS[r], sq[r], z[1], Sj[1], qj[1], jit2, sexp);
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


#ifndef NPGP_FULL_BWD_MINB
#define NPGP_FULL_BWD_MINB 2
#endif
template <int d, bool DX1, bool DX2, int CPT>
__global__ void __launch_bounds__(kNT, NPGP_FULL_BWD_MINB) gibbs_full_bwd_kernel(int n1, int n2, const double* __restrict__ x1,
                                                             const double* __restrict__ S1,
                                                             const double* __restrict__ x2,
                                                             const double* __restrict__ S2, double jit2,
                                                             const double* __restrict__ scale, GSpec g, int vec_ok,
                                                             double* __restrict__ d_S1, double* __restrict__ d_x1,
                                                             double* __restrict__ d_S2, double* __restrict__ d_x2,
                                                             double* __restrict__ d_scale, int rows_per_cta) {
  constexpr int P = sym_size(d);
  constexpr int NRC = P + 1 + (DX1 ? d : 0);  // row-side components: W_p, S0, [XZ_k]
  constexpr int NW = kNT / 32;
  __shared__ double sx[kTI][d], sS[kTI][P], sq[kTI], srs[kTI], srv[kTI];
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
  double z[CPT][d], Sj[CPT][P], qj[CPT], cv[CPT];
  bool valid[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    load_col_full<d>(n2, jbase + c, x2, S2, z[c], Sj[c], qj[c], valid[c]);
    cv[c] = (g.colvec && valid[c]) ? g.colvec[jbase + c] : 0.0;
    qj[c] *= s;  // K, and with it every g K below, carries the outputscale
  }
  double cW[CPT][P], cs0[CPT], cxz[CPT][d];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    cs0[c] = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) cW[c][p] = 0.0;
#pragma unroll
    for (int k = 0; k < d; ++k) cxz[c][k] = 0.0;
  }
  GPrefetch<CPT> gq(g, jbase, valid, vec_ok, row_begin, row_end);

  for (int i0 = row_begin; i0 < row_end; i0 += kTI) {
    __syncthreads();
    stage_rows_full<d>(n1, i0, x1, S1, sx, sS, sq);
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
      double rW[P], rxz[d], rs0 = 0.0;
#pragma unroll
      for (int p = 0; p < P; ++p) rW[p] = 0.0;
#pragma unroll
      for (int k = 0; k < d; ++k) rxz[k] = 0.0;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        FullPair<d> pr;
        gibbs_full_eval<d>(sx[r], sS[r], sq[r], z[c], Sj[c], qj[c], jit2, sexp, &pr);  // qj carries the outputscale
        const double gk = valid[c] ? gv[c] * pr.k : 0.0;
        rs0 += gk;
        cs0[c] += gk;
        // dlogK/dSigma (pair part, same for both sides) = 0.5 w w^T - 0.25 A^-1 = 0.5 w w^T - hr2 adj(At)
        const double c2 = -gk * pr.hr2;
        double t[d];  // 0.5 gk w: also the input gradient, 2 gk w = 4 t (the factor is applied once, after the loops)
#pragma unroll
        for (int a = 0; a < d; ++a) t[a] = (0.5 * gk) * pr.w[a];
#pragma unroll
        for (int a = 0; a < d; ++a)
#pragma unroll
          for (int b = a; b < d; ++b) {
            const int p = sym_idx(d, a, b);
            const double w = fma(t[a], pr.w[b], c2 * pr.adjA[p]);
            rW[p] += w;
            cW[c][p] += w;
          }
        if (DX1 || DX2) {
#pragma unroll
          for (int k = 0; k < d; ++k) {
            if (DX1) rxz[k] += t[k];
            if (DX2) cxz[c][k] += t[k];
          }
        }
      }
#pragma unroll
      for (int p = 0; p < P; ++p) rr.put(r, p, rW[p]);
      rr.put(r, P, rs0);
      if (DX1) {
#pragma unroll
        for (int k = 0; k < d; ++k) rr.put(r, P + 1 + k, rxz[k]);
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
      if (comp < P) {
        atomicAdd(&d_S1[(long)i * P + comp], v);
      } else if (comp == P) {
        // per-point part: 0.25 * S0 * Sigma_i^-1
        double adj[P], det;
        sym_adj_det<d>(sS[r], adj, det);
        const double f = 0.25 * v / det;
#pragma unroll
        for (int p = 0; p < P; ++p) atomicAdd(&d_S1[(long)i * P + p], f * adj[p]);
      } else if (DX1) {
        atomicAdd(&d_x1[(long)i * d + (comp - P - 1)], -4.0 * v);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (!valid[c]) continue;
    const int j = jbase + c;
    double adj[P], det;
    sym_adj_det<d>(Sj[c], adj, det);
    const double f = 0.25 * cs0[c] / det;
#pragma unroll
    for (int p = 0; p < P; ++p) atomicAdd(&d_S2[(long)j * P + p], fma(f, adj[p], cW[c][p]));
    if (DX2) {
#pragma unroll
      for (int k = 0; k < d; ++k) atomicAdd(&d_x2[(long)j * d + k], 4.0 * cxz[c][k]);
    }
  }
  if (d_scale) {  // dL/ds = sum_ij G_ij K_ij / s: the column sums of g K (K carries s)
    double acc_scale = 0.0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc_scale += valid[c] ? cs0[c] : 0.0;
    const double t = block_sum(acc_scale / s, red);
    if (threadIdx.x == 0) atomicAdd(d_scale, t);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Sigma(h) and its backward (multivariate_gibbs_kernel.py:98)
// ---------------------------------------------------------------------------------------------------------------------
template <int d>
__global__ void sigma_from_h_fwd_kernel(int n, const double* __restrict__ H, const double* __restrict__ Dm,
                                        double* __restrict__ S) {
  constexpr int P = sym_size(d);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double h[d], Dl[d * d], out[P];
#pragma unroll
  for (int k = 0; k < d; ++k) h[k] = H[(long)i * d + k];
#pragma unroll
  for (int k = 0; k < d * d; ++k) Dl[k] = Dm[k];
  sigma_from_h_row<d>(h, Dl, out);
#pragma unroll
  for (int p = 0; p < P; ++p) S[(long)i * P + p] = out[p];
}

// dS holds the symmetric-matrix gradient (packed).  dH_m = 4 sum_l G_ml sigma'(u_ml^2) u_ml h_l;  dD_kl = 2 D_kl sum_n G_kl.
template <int d>
__global__ void sigma_from_h_bwd_kernel(int n, const double* __restrict__ H, const double* __restrict__ Dm,
                                        const double* __restrict__ dS, double* __restrict__ dH,
                                        double* __restrict__ dDm) {
  constexpr int P = sym_size(d);
  __shared__ double red[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double gD[P];
#pragma unroll
  for (int p = 0; p < P; ++p) gD[p] = 0.0;
  if (i < n) {
    double h[d], G[P], gh[d];
#pragma unroll
    for (int k = 0; k < d; ++k) {
      h[k] = H[(long)i * d + k];
      gh[k] = 0.0;
    }
#pragma unroll
    for (int p = 0; p < P; ++p) G[p] = dS[(long)i * P + p];
#pragma unroll
    for (int k = 0; k < d; ++k)
#pragma unroll
      for (int l = 0; l < d; ++l) {
        const int p = (k <= l) ? sym_idx(d, k, l) : sym_idx(d, l, k);
        const double u = h[k] * h[l];
        gh[k] += 4.0 * G[p] * sigmoid20(u * u) * u * h[l];
      }
#pragma unroll
    for (int k = 0; k < d; ++k) dH[(long)i * d + k] += gh[k];
#pragma unroll
    for (int p = 0; p < P; ++p) gD[p] = G[p];
  }
  if (dDm) {
#pragma unroll
    for (int k = 0; k < d; ++k)
#pragma unroll
      for (int l = k; l < d; ++l) {
        const double t = block_sum(gD[sym_idx(d, k, l)], red);
        if (threadIdx.x == 0) {
          atomicAdd(&dDm[k * d + l], 2.0 * Dm[k * d + l] * t);
          if (l != k) atomicAdd(&dDm[l * d + k], 2.0 * Dm[l * d + k] * t);
        }
      }
  }
}


template <int d>
static int launch_full_fwd(int n1, int n2, const double* x1, const double* S1, const double* x2, const double* S2,
                           double jitter, const double* scale, double* K, long ldk, const double* u, double* Ku,
                           cudaStream_t st) {
  const int col_tiles = ceil_div(n2, kTJ);
  const int rpc = pick_rows_per_cta(n1, col_tiles, 16);
  dim3 grid(ceil_div(n1, rpc), col_tiles);
  const int vec_ok = (ldk % 2 == 0) && ((reinterpret_cast<uintptr_t>(K) & 15) == 0);
  if (u)
    gibbs_full_fwd_kernel<d, true><<<grid, kNT, 0, st>>>(n1, n2, x1, S1, x2, S2, 2.0 * jitter, scale, K, ldk, vec_ok, u,
                                                         Ku, rpc);
  else
    gibbs_full_fwd_kernel<d, false><<<grid, kNT, 0, st>>>(n1, n2, x1, S1, x2, S2, 2.0 * jitter, scale, K, ldk, vec_ok,
                                                          u, Ku, rpc);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

template <int d, int CPT>
static int launch_full_bwd_cpt(int n1, int n2, const double* x1, const double* S1, const double* x2, const double* S2,
                               double jitter, const double* scale, GSpec g, double* d_S1, double* d_x1, double* d_S2,
                               double* d_x2, double* d_scale, cudaStream_t st) {
  const int col_tiles = ceil_div(n2, kNT * CPT);
  const int rpc = pick_rows_per_cta(n1, col_tiles, 8);
  dim3 grid(ceil_div(n1, rpc), col_tiles);
  const int vec_ok = g.Gm ? ((g.ldg % 2 == 0) && ((reinterpret_cast<uintptr_t>(g.Gm) & 15) == 0)) : 0;
#define NPGP_L(A, B)                                                                                               \
  gibbs_full_bwd_kernel<d, A, B, CPT><<<grid, kNT, 0, st>>>(n1, n2, x1, S1, x2, S2, 2.0 * jitter, scale, g, vec_ok, \
                                                            d_S1, d_x1, d_S2, d_x2, d_scale, rpc)
  if (d_x1 && d_x2) NPGP_L(true, true);
  else if (d_x1) NPGP_L(true, false);
  else if (d_x2) NPGP_L(false, true);
  else NPGP_L(false, false);
#undef NPGP_L
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

// Columns per thread of the backward kernel: 2 (230 registers, 2 CTAs/SM) measured 0.80 ms at B=65536, M=1024, d=3 against
// 0.97 ms for 1 column (124 registers, 4 CTAs/SM) and 0.85 ms for 2 columns forced to 3 CTAs/SM (spills).
template <int d>
static int launch_full_bwd(int n1, int n2, const double* x1, const double* S1, const double* x2, const double* S2,
                           double jitter, const double* scale, GSpec g, double* d_S1, double* d_x1, double* d_S2,
                           double* d_x2, double* d_scale, cudaStream_t st) {
  return launch_full_bwd_cpt<d, 2>(n1, n2, x1, S1, x2, S2, jitter, scale, g, d_S1, d_x1, d_S2, d_x2, d_scale, st);
}

}  // namespace npgp

using namespace npgp;

extern "C" int npgp_gibbs_full_fwd(int d, int n1, int n2, const double* x1, const double* S1, const double* x2,
                                   const double* S2, double jitter, const double* scale, double* K, long ldk,
                                   const double* u, double* Ku, cudaStream_t stream) {
  if (n1 < 0 || n2 < 0 || !K || ldk < n2 || (u && !Ku)) return NPGP_EINVAL;
  if (n1 == 0 || n2 == 0) return NPGP_OK;
  if (!x1 || !S1 || !x2 || !S2) return NPGP_EINVAL;
  if (d == 2) return launch_full_fwd<2>(n1, n2, x1, S1, x2, S2, jitter, scale, K, ldk, u, Ku, stream);
  if (d == 3) return launch_full_fwd<3>(n1, n2, x1, S1, x2, S2, jitter, scale, K, ldk, u, Ku, stream);
  return NPGP_EUNSUPPORTED;
}

extern "C" int npgp_gibbs_full_bwd(int d, int n1, int n2, const double* x1, const double* S1, const double* x2,
                                   const double* S2, double jitter, const double* scale, const double* G, long ldg,
                                   const double* rowscale, const double* rowvec, const double* colvec, double* d_S1,
                                   double* d_x1, double* d_S2, double* d_x2, double* d_scale, cudaStream_t stream) {
  if (n1 < 0 || n2 < 0 || (!G && !rowvec) || (G && ldg < n2) || ((rowvec != nullptr) != (colvec != nullptr)))
    return NPGP_EINVAL;
  if (n1 == 0 || n2 == 0) return NPGP_OK;
  if (!x1 || !S1 || !x2 || !S2 || !d_S1 || !d_S2) return NPGP_EINVAL;
  GSpec g{G, ldg, rowscale, rowvec, colvec};
  if (d == 2) return launch_full_bwd<2>(n1, n2, x1, S1, x2, S2, jitter, scale, g, d_S1, d_x1, d_S2, d_x2, d_scale, stream);
  if (d == 3) return launch_full_bwd<3>(n1, n2, x1, S1, x2, S2, jitter, scale, g, d_S1, d_x1, d_S2, d_x2, d_scale, stream);
  return NPGP_EUNSUPPORTED;
}

extern "C" int npgp_sigma_from_h_fwd(int d, int n, const double* H, const double* Dm, double* S, cudaStream_t stream) {
  if (n < 0 || !Dm) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  if (!H || !S) return NPGP_EINVAL;
  const int nb = ceil_div(n, 256);
  if (d == 2) sigma_from_h_fwd_kernel<2><<<nb, 256, 0, stream>>>(n, H, Dm, S);
  else if (d == 3) sigma_from_h_fwd_kernel<3><<<nb, 256, 0, stream>>>(n, H, Dm, S);
  else return NPGP_EUNSUPPORTED;
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_sigma_from_h_bwd(int d, int n, const double* H, const double* Dm, const double* dS, double* dH,
                                     double* dDm, cudaStream_t stream) {
  if (n < 0 || !Dm) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  if (!H || !dS || !dH) return NPGP_EINVAL;
  const int nb = ceil_div(n, 256);
  if (d == 2) sigma_from_h_bwd_kernel<2><<<nb, 256, 0, stream>>>(n, H, Dm, dS, dH, dDm);
  else if (d == 3) sigma_from_h_bwd_kernel<3><<<nb, 256, 0, stream>>>(n, H, Dm, dS, dH, dDm);
  else return NPGP_EUNSUPPORTED;
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
