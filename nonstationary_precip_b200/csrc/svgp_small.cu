// Small kernels of the SVGP-Gibbs ELBO step: weighted column sums K^T w, the Gaussian expected-log-likelihood
// reduction with its per-row gradient seeds, the Cholesky-backward mask, and a fused Adam update over the flat
// parameter buffer.  (VariationalELBO / GaussianLikelihood.expected_log_prob semantics: SURVEY.md Appendix B.4.)
#include "common.cuh"

namespace npgp {

// out[j] += sum_i w_i K[i,j]   (K is n x M row-major).  Block = 256 columns x row slice; coalesced along j.
__global__ void __launch_bounds__(256) colwsum_kernel(int n, int M, const double* __restrict__ K, long ldk,
                                                      const double* __restrict__ w, int rows_per_cta,
                                                      double* __restrict__ out) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
  if (j >= M) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int i = r0;
  for (; i + 3 < r1; i += 4) {
    a0 = fma(w ? w[i] : 1.0, K[(long)i * ldk + j], a0);
    a1 = fma(w ? w[i + 1] : 1.0, K[(long)(i + 1) * ldk + j], a1);
    a2 = fma(w ? w[i + 2] : 1.0, K[(long)(i + 2) * ldk + j], a2);
    a3 = fma(w ? w[i + 3] : 1.0, K[(long)(i + 3) * ldk + j], a3);
  }
  for (; i < r1; ++i) a0 = fma(w ? w[i] : 1.0, K[(long)i * ldk + j], a0);
  atomicAdd(&out[j], (a0 + a1) + (a2 + a3));
}

// out[i] = sum_j A[i,j] v[j]  (one warp per row)
__global__ void __launch_bounds__(256) gemv_n_kernel(int n, int M, const double* __restrict__ A, long lda,
                                                     const double* __restrict__ v, double* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  double a = 0.0;
  for (int j = lane; j < M; j += 32) a = fma(A[(long)row * lda + j], v[j], a);
  a = warp_sum(a);
  if (lane == 0) out[row] = a;
}

// Gaussian expected log-likelihood over rows and gradient seeds.
//   v_i = max(kdiag + jitter_xx + q_i, min_var);  ell_i = -0.5 [ ((y-mu)^2 + v)/s2 + log s2 + log 2pi ]
//   acc[0] += sum ell_i;  acc[1] += sum ((y-mu)^2 + v)   (for d/d s2);  acc[2] += #unclamped rows
//   gmu_i = wscale (y-mu)/s2;   gv_i = -0.5 wscale / s2 (0 where clamped)
__global__ void __launch_bounds__(256) gauss_ell_kernel(int n, const double* __restrict__ y,
                                                        const double* __restrict__ mu, const double* __restrict__ q,
                                                        const double* __restrict__ kdiag_p, double jitter_xx,
                                                        double min_var, const double* __restrict__ noise_p,
                                                        double wscale, double* __restrict__ var_out,
                                                        double* __restrict__ gmu, double* __restrict__ gv,
                                                        double* __restrict__ acc) {
  __shared__ double red[32];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const double s2 = *noise_p, kd = *kdiag_p;
  double e = 0.0, r2v = 0.0, cnt = 0.0;
  if (i < n) {
    double v = kd + jitter_xx + q[i];
    const bool clamped = v < min_var;
    if (clamped) v = min_var;
    const double r = y[i] - mu[i];
    e = -0.5 * ((r * r + v) / s2 + log(s2) + 1.8378770664093453);
    r2v = r * r + v;
    cnt = clamped ? 0.0 : 1.0;
    if (var_out) var_out[i] = v;
    gmu[i] = wscale * r / s2;
    gv[i] = clamped ? 0.0 : -0.5 * wscale / s2;
  }
  double t = block_sum(e, red);
  if (threadIdx.x == 0) atomicAdd(&acc[0], t);
  t = block_sum(r2v, red);
  if (threadIdx.x == 0) atomicAdd(&acc[1], t);
  t = block_sum(cnt, red);
  if (threadIdx.x == 0) atomicAdd(&acc[2], t);
}

// Phi(X): lower triangle with halved diagonal, in place (Cholesky backward, Murray 2016), scaled by alpha
__global__ void phi_mask_kernel(int M, double* X, long ldx, double alpha) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= M || c >= M) return;
  double* p = X + (long)r * ldx + c;
  *p = (c < r) ? alpha * *p : ((c == r) ? 0.5 * alpha * *p : 0.0);
}

// Adam (torch.optim.Adam semantics, no weight decay / amsgrad), maximize=false:  p -= lr * mhat / (sqrt(vhat) + eps)
__global__ void adam_kernel(long n, double* __restrict__ p, const double* __restrict__ g, double* __restrict__ m,
                            double* __restrict__ v, const double* __restrict__ mask, double lr, double b1, double b2,
                            double eps, double bc1, double bc2, double gscale) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && mask[i] == 0.0) return;
  const double gi = gscale * g[i];
  const double mi = b1 * m[i] + (1.0 - b1) * gi;
  const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= lr * (mi / bc1) / (sqrt(vi / bc2) + eps);
}

// Adam with the step counter on the device (so that a captured CUDA graph replays with the right bias correction):
// step_dev[0] holds the number of steps taken so far; bump_step_kernel increments it after the update.
__global__ void adam_dev_kernel(long n, double* __restrict__ p, const double* __restrict__ g, double* __restrict__ m,
                                double* __restrict__ v, const double* __restrict__ mask, double lr, double b1,
                                double b2, double eps, const double* __restrict__ step_dev, double gscale) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && mask[i] == 0.0) return;
  const double t = step_dev[0] + 1.0;
  const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  const double gi = gscale * g[i];
  const double mi = b1 * m[i] + (1.0 - b1) * gi;
  const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= lr * (mi / bc1) / (sqrt(vi / bc2) + eps);
}

__global__ void bump_step_kernel(double* step_dev) { step_dev[0] += 1.0; }

// ---- failure handling inside a captured step (psd_safe_cholesky, models/gibbs_kernels.py:201, raises on a failed
// factorisation; a replayed graph cannot raise, so the step records the failure on the device and protects the parameters)
// status (sticky): bit 0 = a Cholesky reported a bad pivot / timed out, bit 1 = the loss is not finite
__global__ void status_update_kernel(int* __restrict__ status, const int* __restrict__ info, const double* __restrict__ loss) {
  int s = 0;
  if (info && *info != 0) s |= 1;
  if (loss && !(fabs(*loss) < 1.7e308)) s |= 2;
  if (s) atomicOr(status, s);
}

// Adam as adam_dev_kernel, skipped entirely (parameters, moments and step counter untouched) while *status != 0
__global__ void adam_guarded_kernel(long n, double* __restrict__ p, const double* __restrict__ g, double* __restrict__ m,
                                    double* __restrict__ v, const double* __restrict__ mask, double lr, double b1, double b2,
                                    double eps, const double* __restrict__ step_dev, double gscale,
                                    const int* __restrict__ status) {
  if (*status != 0) return;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && mask[i] == 0.0) return;
  const double t = step_dev[0] + 1.0;
  const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  const double gi = gscale * g[i];
  const double mi = b1 * m[i] + (1.0 - b1) * gi;
  const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= lr * (mi / bc1) / (sqrt(vi / bc2) + eps);
}

__global__ void bump_step_guarded_kernel(double* step_dev, const int* __restrict__ status) {
  if (*status == 0) step_dev[0] += 1.0;
}

// ---- deterministic variants used with the digit-plane path (no FP64 atomics) -----------------------------------------
// mu_i = sum_s mu_part[s * stride + i] (partials of the fused K u, added in index order);  gmu_i = wscale (y_i - mu_i)/noise
__global__ void __launch_bounds__(256) mu_gmu_parts_kernel(int n, const double* __restrict__ y,
                                                           const double* __restrict__ mu_part, int nparts, long stride,
                                                           const double* __restrict__ noise_p, double wscale,
                                                           double* __restrict__ mu, double* __restrict__ gmu) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double m = 0.0;
  for (int s = 0; s < nparts; ++s) m += mu_part[(long)s * stride + i];
  mu[i] = m;
  gmu[i] = wscale * (y[i] - m) / *noise_p;
}

// As gauss_ell_kernel with q_i = sum_c q_part[c * q_stride + i]; per-block sums go to blk[3][gridDim.x] (plain stores) and
// gauss_ell_finish_kernel adds them in block order.  Clamped rows are appended to skip_rows (their weight in K^T W K is 0).
__global__ void __launch_bounds__(256) gauss_ell_parts_kernel(int n, const double* __restrict__ y, const double* __restrict__ mu,
                                                              const double* __restrict__ q_part, int nq, long q_stride,
                                                              const double* __restrict__ kdiag_p, double jitter_xx,
                                                              double min_var, const double* __restrict__ noise_p,
                                                              double wscale, double* __restrict__ var_out,
                                                              double* __restrict__ gmu, double* __restrict__ gv,
                                                              double* __restrict__ blk, int* __restrict__ skip_count,
                                                              int* __restrict__ skip_rows) {
  __shared__ double red[32];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const double s2 = *noise_p, kd = *kdiag_p;
  double e = 0.0, r2v = 0.0, cnt = 0.0;
  if (i < n) {
    double qi = 0.0;
    for (int c = 0; c < nq; ++c) qi += q_part[(long)c * q_stride + i];
    double v = kd + jitter_xx + qi;
    const bool clamped = v < min_var;
    if (clamped) v = min_var;
    const double r = y[i] - mu[i];
    e = -0.5 * ((r * r + v) / s2 + log(s2) + 1.8378770664093453);
    r2v = r * r + v;
    cnt = clamped ? 0.0 : 1.0;
    if (var_out) var_out[i] = v;
    gmu[i] = wscale * r / s2;
    gv[i] = clamped ? 0.0 : -0.5 * wscale / s2;
    if (clamped && skip_count) skip_rows[atomicAdd(skip_count, 1)] = i;
  }
  double t = block_sum(e, red);
  if (threadIdx.x == 0) blk[blockIdx.x] = t;
  t = block_sum(r2v, red);
  if (threadIdx.x == 0) blk[gridDim.x + blockIdx.x] = t;
  t = block_sum(cnt, red);
  if (threadIdx.x == 0) blk[2 * gridDim.x + blockIdx.x] = t;
}

// acc4 = [sum ell, sum ((y-mu)^2 + v), #unclamped rows, w0 = -0.5 wscale / noise (the weight of every unclamped row)]
__global__ void __launch_bounds__(256) gauss_ell_finish_kernel(int nblk, const double* __restrict__ blk,
                                                               const double* __restrict__ noise_p, double wscale,
                                                               double* __restrict__ acc4) {
  __shared__ double red[32];
  for (int k = 0; k < 3; ++k) {
    double a = 0.0;
    for (int b = threadIdx.x; b < nblk; b += 256) a += blk[(long)k * nblk + b];  // fixed assignment, fixed tree
    const double t = block_sum(a, red);
    if (threadIdx.x == 0) acc4[k] = t;
  }
  if (threadIdx.x == 0) acc4[3] = -0.5 * wscale / *noise_p;
}

}  // namespace npgp

using namespace npgp;

extern "C" int npgp_adam_step_dev(long n, double* p, const double* g, double* m, double* v, const double* mask,
                                  double lr, double beta1, double beta2, double eps, double* step_dev, double gscale,
                                  cudaStream_t stream) {
  if (n < 0 || !step_dev) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  if (!p || !g || !m || !v) return NPGP_EINVAL;
  adam_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, p, g, m, v, mask, lr, beta1, beta2, eps, step_dev,
                                                                   gscale);
  NPGP_LAUNCH_CHECK();
  bump_step_kernel<<<1, 1, 0, stream>>>(step_dev);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

// status (device int, sticky, zeroed by the caller once): |= 1 if *info != 0 (info may be NULL), |= 2 if *loss is not finite
// (loss may be NULL).  Lets a captured step record a failed factorisation / a NaN objective without a host round trip.
extern "C" int npgp_status_update(int* status, const int* info, const double* loss, cudaStream_t stream) {
  if (!status) return NPGP_EINVAL;
  status_update_kernel<<<1, 1, 0, stream>>>(status, info, loss);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

// npgp_adam_step_dev that leaves parameters, moments and the step counter untouched while *status != 0: a failed step of a
// replayed graph cannot corrupt the model; the host polls status and re-runs with more jitter (SVGPGibbs.recover).
extern "C" int npgp_adam_step_guarded(long n, double* p, const double* g, double* m, double* v, const double* mask, double lr,
                                      double beta1, double beta2, double eps, double* step_dev, double gscale,
                                      const int* status, cudaStream_t stream) {
  if (n < 0 || !step_dev || !status) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  if (!p || !g || !m || !v) return NPGP_EINVAL;
  adam_guarded_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, p, g, m, v, mask, lr, beta1, beta2, eps, step_dev,
                                                                       gscale, status);
  NPGP_LAUNCH_CHECK();
  bump_step_guarded_kernel<<<1, 1, 0, stream>>>(step_dev, status);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_colwsum(int n, int M, const double* K, long ldk, const double* w, double* out,
                            cudaStream_t stream) {
  if (n < 0 || M < 0) return NPGP_EINVAL;
  if (n == 0 || M == 0) return NPGP_OK;
  if (!K || !out) return NPGP_EINVAL;
  const int cb = ceil_div(M, 256);
  long row_ctas = (kNumSMs * 8 + cb - 1) / cb;
  long rpc = (n + row_ctas - 1) / row_ctas;
  if (rpc < 64) rpc = 64;
  dim3 grid(cb, ceil_div(n, rpc));
  colwsum_kernel<<<grid, 256, 0, stream>>>(n, M, K, ldk, w, (int)rpc, out);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_gemv_n(int n, int M, const double* A, long lda, const double* v, double* out, cudaStream_t stream) {
  if (n < 0 || M < 0) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  if (!A || !v || !out) return NPGP_EINVAL;
  gemv_n_kernel<<<ceil_div(n, 8), 256, 0, stream>>>(n, M, A, lda, v, out);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_gauss_ell(int n, const double* y, const double* mu, const double* q, const double* kdiag,
                              double jitter_xx, double min_var, const double* noise, double wscale, double* var_out,
                              double* gmu, double* gv, double* acc3, cudaStream_t stream) {
  if (n < 0) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  if (!y || !mu || !q || !kdiag || !noise || !gmu || !gv || !acc3) return NPGP_EINVAL;
  gauss_ell_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(n, y, mu, q, kdiag, jitter_xx, min_var, noise, wscale, var_out,
                                                         gmu, gv, acc3);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}


// mu = sum of `nparts` partial mean vectors (stride apart, index order) and gmu = wscale (y - mu) / *noise: the first
// gradient seed of the ELBO, available before the predictive variance is.
extern "C" int npgp_mu_gmu_parts(int n, const double* y, const double* mu_part, int nparts, long stride, const double* noise,
                                 double wscale, double* mu, double* gmu, cudaStream_t stream) {
  if (n < 0 || nparts < 1 || stride < n) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  if (!y || !mu_part || !noise || !mu || !gmu) return NPGP_EINVAL;
  mu_gmu_parts_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(n, y, mu_part, nparts, stride, noise, wscale, mu, gmu);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" long npgp_gauss_ell_parts_workspace_bytes(int n) { return 3L * ceil_div(n > 0 ? n : 1, 256) * (long)sizeof(double); }

// npgp_gauss_ell with the quadratic term given as `nq` partial vectors (q_i = sum_c q_part[c * q_stride + i], index order)
// and all reductions two-stage (bitwise reproducible).  acc4 (overwritten) = [sum ell, sum ((y-mu)^2 + v), #unclamped rows,
// w0 = -0.5 wscale / noise].  skip_count / skip_rows (optional; *skip_count zeroed by the caller): indices of clamped rows.
extern "C" int npgp_gauss_ell_parts(int n, const double* y, const double* mu, const double* q_part, int nq, long q_stride,
                                    const double* kdiag, double jitter_xx, double min_var, const double* noise, double wscale,
                                    double* var_out, double* gmu, double* gv, double* acc4, int* skip_count, int* skip_rows,
                                    void* work, long work_bytes, cudaStream_t stream) {
  if (n < 0 || nq < 0 || (nq > 0 && q_stride < n) || ((skip_count != nullptr) != (skip_rows != nullptr))) return NPGP_EINVAL;
  if (!acc4 || !noise) return NPGP_EINVAL;
  if (n > 0 && (!y || !mu || (nq > 0 && !q_part) || !kdiag || !gmu || !gv || !work)) return NPGP_EINVAL;
  if (n > 0 && work_bytes < npgp_gauss_ell_parts_workspace_bytes(n)) return NPGP_EWORKSPACE;
  const int nblk = ceil_div(n, 256);
  if (n > 0) {
    gauss_ell_parts_kernel<<<nblk, 256, 0, stream>>>(n, y, mu, q_part, nq, q_stride, kdiag, jitter_xx, min_var, noise, wscale,
                                                     var_out, gmu, gv, (double*)work, skip_count, skip_rows);
    NPGP_LAUNCH_CHECK();
  }
  gauss_ell_finish_kernel<<<1, 256, 0, stream>>>(nblk, (const double*)work, noise, wscale, acc4);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_phi_mask(int M, double* X, long ldx, double alpha, cudaStream_t stream) {
  if (M < 0 || (M > 0 && !X)) return NPGP_EINVAL;
  if (M == 0) return NPGP_OK;
  dim3 blk(32, 8), grd(ceil_div(M, 32), ceil_div(M, 8));
  phi_mask_kernel<<<grd, blk, 0, stream>>>(M, X, ldx, alpha);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_adam_step(long n, double* p, const double* g, double* m, double* v, const double* mask, double lr,
                              double beta1, double beta2, double eps, int step, double gscale, cudaStream_t stream) {
  if (n < 0 || step < 1) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  if (!p || !g || !m || !v) return NPGP_EINVAL;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, p, g, m, v, mask, lr, beta1, beta2, eps, bc1, bc2,
                                                               gscale);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
