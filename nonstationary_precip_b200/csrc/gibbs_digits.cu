// Gibbs cross-covariance tiles emitted directly as byte-digit planes for the int8 tensor-core contractions (oz8.cu):
// K(X,Z) = scale * Gibbs kernel is evaluated pair by pair in FP64 exactly as in gibbs_diag.cu / gibbs_full.cu (reference
// models/gibbs_kernels.py:154-162, models/multivariate_gibbs_kernel.py:101-150, models/sparse_multivariate_gibbs_kernel.py:
// 105-154), but instead of 8 bytes of FP64 per pair the kernel stores the 7 signed base-256 digits of y = rint(K_ij 2^(54-e)),
// e = exponent of the outputscale (0 <= K_ij <= scale), in the row layout of oz8.cuh.  The N x M matrix then never exists in FP64: the
// row-quadratic product T = K C reads the planes K-major, the SYRK K^T K reads the same planes MN-major, and the row dot
// rebuilds its K tile from them.  The fused matrix-vector product K u (predictive mean) is accumulated from the FP64 values.
//
// Tiling: one thread per ROW (its point and latent field value in registers), the CTA's columns staged once in shared
// memory and read by broadcast; a thread finishes 16 consecutive columns of its row, i.e. exactly one 16-byte vector per
// digit plane, and the 32 lanes of a warp (32 consecutive rows) write 512 contiguous bytes per plane.
#include <cstdint>

#include "common.cuh"
#include "oz8.cuh"
#include "pairmath.cuh"

namespace npgp {

constexpr int GD_ROWS = 128;      // rows per CTA = threads per CTA = rows of a digit-plane block
constexpr int GD_MAXCOLS = 256;   // columns staged per CTA

template <int d>
struct FullColumn {  // column-side data of the full-matrix kernel, padded to whole 16-byte vectors
  static constexpr int P = sym_size(d);
  static constexpr int N = ((d + P + 2) + 1) / 2 * 2;  // z[d], S[P], q*scale, u
};

// Ku_part[blockIdx.y * ku_stride + row] = sum over this CTA's columns of K_ij u_j (plain stores; the consumer adds the
// gridDim.y partials in index order).  n2 % 16 == 0; rows >= n1 (up to the 128-row block) get zero digits.
template <int d, bool HAS_U>
__global__ void __launch_bounds__(GD_ROWS) gibbs_full_fwd_digits_kernel(
    int n1, int n2, const double* __restrict__ x1, const double* __restrict__ S1, const double* __restrict__ x2,
    const double* __restrict__ S2, double jit2, const double* __restrict__ scale, int8_t* __restrict__ digits,
    const double* __restrict__ u, double* __restrict__ Ku_part, long ku_stride, int cols_per_cta) {
  constexpr int P = sym_size(d);
  constexpr int CN = FullColumn<d>::N;
  __shared__ __align__(16) double scol[GD_MAXCOLS][CN];
  __shared__ double sexp[256];
  load_exp_table(sexp);
  const int r = blockIdx.x * GD_ROWS + threadIdx.x;
  const int c_begin = blockIdx.y * cols_per_cta;
  const int nc = min(n2, c_begin + cols_per_cta) - c_begin;
  const double s = scale ? *scale : 1.0;
  const int e = o8_exponent_of_scale(s);
  const double sc = (e == O8_POISON) ? 0.0 : o8_pow2(O8_FRAC - e);
  for (int c = threadIdx.x; c < nc; c += GD_ROWS) {
    const int j = c_begin + c;
    double Sj[P];
#pragma unroll
    for (int k = 0; k < d; ++k) scol[c][k] = x2[(long)j * d + k];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      Sj[p] = S2[(long)j * P + p];
      scol[c][d + p] = Sj[p];
    }
    scol[c][d + P] = full_col_factor<d>(sym_det<d>(Sj)) * s;  // 2^(d/2) det(S_j)^(1/4) * outputscale
    scol[c][d + P + 1] = HAS_U ? u[j] : 0.0;
  }
  double xi[d], Si[P];
#pragma unroll
  for (int k = 0; k < d; ++k) xi[k] = (r < n1) ? x1[(long)r * d + k] : 0.0;
#pragma unroll
  for (int p = 0; p < P; ++p) Si[p] = 0.0;
#pragma unroll
  for (int k = 0; k < d; ++k) Si[sym_idx(d, k, k)] = 1.0;
  if (r < n1) {
#pragma unroll
    for (int p = 0; p < P; ++p) Si[p] = S1[(long)r * P + p];
  }
  const double qi = (r < n1) ? sqrt(sqrt(sym_det<d>(Si))) : 0.0;  // a padded row evaluates to K = 0: zero digits
  __syncthreads();
  const int nks = n2 / O8_KS;
  double acc = 0.0;
  for (int c0 = 0; c0 < nc; c0 += 16) {
    long long y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const double* cd = scol[c0 + j];
      double col[CN];
#pragma unroll
      for (int v = 0; v < CN; v += 2) {
        const double2 t = *reinterpret_cast<const double2*>(cd + v);
        col[v] = t.x;
        col[v + 1] = t.y;
      }
      const double k = gibbs_full_eval<d>(xi, Si, qi, col, col + d, col[d + P], jit2, sexp);
      y[j] = o8_quantise(k, sc);
      if (HAS_U) acc = fma(k, col[d + P + 1], acc);
    }
    const int jg = c_begin + c0;
    o8_store_digits(y, digits + o8_a_offset(r, jg, nks), (long)nks * O8_A_PLANE);
  }
  if (HAS_U && r < n1) Ku_part[(long)blockIdx.y * ku_stride + r] = acc;
}

// diagonal Gibbs kernel: ell1 (D,n1), ell2 (D,n2) dim-major as in gibbs_diag.cu
template <int D, bool HAS_U>
__global__ void __launch_bounds__(GD_ROWS) gibbs_diag_fwd_digits_kernel(
    int n1, int n2, const double* __restrict__ x1, const double* __restrict__ ell1, const double* __restrict__ x2,
    const double* __restrict__ ell2, const double* __restrict__ scale, int8_t* __restrict__ digits,
    const double* __restrict__ u, double* __restrict__ Ku_part, long ku_stride, int cols_per_cta) {
  constexpr int CN = (2 * D + 2 + 1) / 2 * 2;  // z[D], l^2[D], c*scale, u
  __shared__ __align__(16) double scol[GD_MAXCOLS][CN];
  __shared__ double sexp[256];
  load_exp_table(sexp);
  const int r = blockIdx.x * GD_ROWS + threadIdx.x;
  const int c_begin = blockIdx.y * cols_per_cta;
  const int nc = min(n2, c_begin + cols_per_cta) - c_begin;
  const double s = scale ? *scale : 1.0;
  const int e = o8_exponent_of_scale(s);
  const double sc = (e == O8_POISON) ? 0.0 : o8_pow2(O8_FRAC - e);
  for (int c = threadIdx.x; c < nc; c += GD_ROWS) {
    const int j = c_begin + c;
    double prod = 1.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double lv = ell2[(long)k * n2 + j];
      scol[c][k] = x2[(long)j * D + k];
      scol[c][D + k] = lv * lv;
      prod *= 1.4142135623730951 * lv;
    }
    scol[c][2 * D] = sqrt(prod) * s;
    scol[c][2 * D + 1] = HAS_U ? u[j] : 0.0;
  }
  double xi[D], ai[D], prod = 1.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    xi[k] = (r < n1) ? x1[(long)r * D + k] : 0.0;
    const double lv = (r < n1) ? ell1[(long)k * n1 + r] : 1.0;
    ai[k] = lv * lv;
    prod *= 1.4142135623730951 * lv;
  }
  const double ci = (r < n1) ? sqrt(prod) : 0.0;
  __syncthreads();
  const int nks = n2 / O8_KS;
  double acc = 0.0;
  for (int c0 = 0; c0 < nc; c0 += 16) {
    long long y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const double* cd = scol[c0 + j];
      double col[CN];
#pragma unroll
      for (int v = 0; v < CN; v += 2) {
        const double2 t = *reinterpret_cast<const double2*>(cd + v);
        col[v] = t.x;
        col[v + 1] = t.y;
      }
      const double k = gibbs_diag_eval<D>(xi, ai, ci, col, col + D, col[2 * D], sexp);
      y[j] = o8_quantise(k, sc);
      if (HAS_U) acc = fma(k, col[2 * D + 1], acc);
    }
    const int jg = c_begin + c0;
    o8_store_digits(y, digits + o8_a_offset(r, jg, nks), (long)nks * O8_A_PLANE);
  }
  if (HAS_U && r < n1) Ku_part[(long)blockIdx.y * ku_stride + r] = acc;
}

// columns per CTA: a multiple of 16 that gives >= ~4 CTAs per SM when the row count alone does not
static int gd_cols_per_cta(int n1, int n2) {
  const int row_ctas = ceil_div(n1, GD_ROWS);
  int splits = ceil_div(4L * kNumSMs, row_ctas);
  const int min_splits = ceil_div(n2, GD_MAXCOLS);
  if (splits < min_splits) splits = min_splits;
  int cols = ceil_div(ceil_div(n2, splits), 16) * 16;
  if (cols < 64) cols = 64;
  if (cols > GD_MAXCOLS) cols = GD_MAXCOLS;
  return cols;
}

}  // namespace npgp

using namespace npgp;

// number of column splits (= partial vectors of the fused K u) npgp_gibbs_*_fwd_digits uses for an n1 x n2 problem
extern "C" int npgp_gibbs_digits_splits(int n1, int n2) {
  if (n1 <= 0 || n2 <= 0) return 0;
  return ceil_div(n2, gd_cols_per_cta(n1, n2));
}

// Full-matrix Gibbs kernel as digit planes.  digits: npgp_o8_digits_bytes(n1, n2, 128) bytes; n2 % 32 == 0.
// u / Ku_part (optional, both or none): Ku_part (splits x ku_stride, ku_stride >= n1) receives the per-split partials of K u.
extern "C" int npgp_gibbs_full_fwd_digits(int d, int n1, int n2, const double* x1, const double* S1, const double* x2,
                                          const double* S2, double jitter, const double* scale, void* digits,
                                          const double* u, double* Ku_part, long ku_stride, cudaStream_t stream) {
  if (n1 < 0 || n2 < 0 || (u != nullptr) != (Ku_part != nullptr) || (u && ku_stride < n1)) return NPGP_EINVAL;
  if (n1 == 0 || n2 == 0) return NPGP_OK;
  if (!x1 || !S1 || !x2 || !S2 || !digits || !scale) return NPGP_EINVAL;
  if (n2 % O8_KS) return NPGP_EUNSUPPORTED;
  const int cols = gd_cols_per_cta(n1, n2);
  dim3 grid(ceil_div(n1, GD_ROWS), ceil_div(n2, cols));
#define NPGP_L(DD, UU)                                                                                                    \
  gibbs_full_fwd_digits_kernel<DD, UU><<<grid, GD_ROWS, 0, stream>>>(n1, n2, x1, S1, x2, S2, 2.0 * jitter, scale,          \
                                                                      (int8_t*)digits, u, Ku_part, ku_stride, cols)
  if (d == 2) { if (u) NPGP_L(2, true); else NPGP_L(2, false); }
  else if (d == 3) { if (u) NPGP_L(3, true); else NPGP_L(3, false); }
  else return NPGP_EUNSUPPORTED;
#undef NPGP_L
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_gibbs_diag_fwd_digits(int D, int n1, int n2, const double* x1, const double* ell1, const double* x2,
                                          const double* ell2, const double* scale, void* digits, const double* u,
                                          double* Ku_part, long ku_stride, cudaStream_t stream) {
  if (n1 < 0 || n2 < 0 || (u != nullptr) != (Ku_part != nullptr) || (u && ku_stride < n1)) return NPGP_EINVAL;
  if (n1 == 0 || n2 == 0) return NPGP_OK;
  if (!x1 || !ell1 || !x2 || !ell2 || !digits || !scale) return NPGP_EINVAL;
  if (n2 % O8_KS) return NPGP_EUNSUPPORTED;
  const int cols = gd_cols_per_cta(n1, n2);
  dim3 grid(ceil_div(n1, GD_ROWS), ceil_div(n2, cols));
#define NPGP_L(DD, UU)                                                                                                    \
  gibbs_diag_fwd_digits_kernel<DD, UU><<<grid, GD_ROWS, 0, stream>>>(n1, n2, x1, ell1, x2, ell2, scale, (int8_t*)digits, u, \
                                                                      Ku_part, ku_stride, cols)
#define NPGP_LU(DD) do { if (u) NPGP_L(DD, true); else NPGP_L(DD, false); } while (0)
  switch (D) {
    case 1: NPGP_LU(1); break;
    case 2: NPGP_LU(2); break;
    case 3: NPGP_LU(3); break;
    case 4: NPGP_LU(4); break;
    case 5: NPGP_LU(5); break;
    case 6: NPGP_LU(6); break;
    default: return NPGP_EUNSUPPORTED;
  }
#undef NPGP_LU
#undef NPGP_L
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
