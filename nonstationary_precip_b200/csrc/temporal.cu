// Temporal kernel of the spatio-temporal model: outputscale * RBF(t,t') * Periodic(t,t') on one input dimension,
//   k = s * exp(-0.5 tau^2 / l_r^2) * exp(-2 sin^2(pi |tau| / p) / l_p),   tau = t - t'
// (reference models/spatio_temporal_models.py:42: ScaleKernel(RBFKernel * PeriodicKernel, outputscale >= 7); GPyTorch
// <= 1.8 PeriodicKernel convention).  Forward writes K (n1 x n2); backward reduces an upstream
// gradient to the four hyper-parameter gradients (and, optionally, to the second input) with warp + block reductions.
#include "common.cuh"

namespace npgp {

// hyp (device): [lengthscale_rbf, lengthscale_per, period, outputscale]
__global__ void __launch_bounds__(256) rbfper_fwd_kernel(int n1, int n2, const double* __restrict__ t1,
                                                         const double* __restrict__ t2, const double* __restrict__ hyp,
                                                         double* __restrict__ K, long ldk) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int i0 = blockIdx.y * 32;
  if (j >= n2) return;
  const double lr = hyp[0], lp = hyp[1], per = hyp[2], os = hyp[3];
  const double a = -0.5 / (lr * lr), b = -2.0 / lp, ip = 1.0 / per;
  const double tj = t2[j];
  for (int i = i0; i < min(n1, i0 + 32); ++i) {
    const double tau = t1[i] - tj;
    const double sn = sinpi(fabs(tau) * ip);
    K[(long)i * ldk + j] = os * exp(fma(a * tau, tau, b * sn * sn));
  }
}

// out5: d/d lengthscale_rbf, d/d lengthscale_per, d/d period, d/d outputscale, (unused); dt2 (n2) optional
__global__ void __launch_bounds__(256) rbfper_bwd_kernel(int n1, int n2, const double* __restrict__ t1,
                                                         const double* __restrict__ t2, const double* __restrict__ hyp,
                                                         const double* __restrict__ G, long ldg, int rows_per_cta,
                                                         double* __restrict__ out4, double* __restrict__ dt2) {
  __shared__ double red[32];
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int i0 = blockIdx.y * rows_per_cta, i1 = min(n1, i0 + rows_per_cta);
  const double lr = hyp[0], lp = hyp[1], per = hyp[2], os = hyp[3];
  const double a = -0.5 / (lr * lr), b = -2.0 / lp, ip = 1.0 / per;
  double g_lr = 0.0, g_lp = 0.0, g_p = 0.0, g_os = 0.0, g_t = 0.0;
  if (j < n2) {
    const double tj = t2[j];
    for (int i = i0; i < i1; ++i) {
      const double tau = t1[i] - tj, at = fabs(tau);
      double sn, cs;
      sincospi(at * ip, &sn, &cs);
      const double k0 = exp(fma(a * tau, tau, b * sn * sn));  // without outputscale
      const double gk = G[(long)i * ldg + j] * k0;
      g_os += gk;
      const double gks = gk * os;
      g_lr += gks * tau * tau / (lr * lr * lr);
      g_lp += gks * 2.0 * sn * sn / (lp * lp);
      // d/dp of -2 sin^2(pi at / p)/lp = (4 sn cs / lp) * pi at / p^2
      const double dphase = 4.0 * sn * cs / lp * 3.14159265358979323846;
      g_p += gks * dphase * at * ip * ip;
      // d/dtau: -tau/lr^2 - (4 sn cs / lp) * pi sign(tau)/p ;  d/dt2 = -d/dtau
      g_t += gks * (tau / (lr * lr) + dphase * ip * (tau >= 0.0 ? 1.0 : -1.0));
    }
    if (dt2) atomicAdd(&dt2[j], g_t);
  }
  double t = block_sum(g_lr, red);
  if (threadIdx.x == 0) atomicAdd(&out4[0], t);
  t = block_sum(g_lp, red);
  if (threadIdx.x == 0) atomicAdd(&out4[1], t);
  t = block_sum(g_p, red);
  if (threadIdx.x == 0) atomicAdd(&out4[2], t);
  t = block_sum(g_os, red);
  if (threadIdx.x == 0) atomicAdd(&out4[3], t);
}

}  // namespace npgp

using namespace npgp;

extern "C" int npgp_rbfper_fwd(int n1, int n2, const double* t1, const double* t2, const double* hyp, double* K,
                               long ldk, cudaStream_t stream) {
  if (n1 < 0 || n2 < 0) return NPGP_EINVAL;
  if (n1 == 0 || n2 == 0) return NPGP_OK;
  if (!t1 || !t2 || !hyp || !K || ldk < n2) return NPGP_EINVAL;
  dim3 grid(ceil_div(n2, 256), ceil_div(n1, 32));
  rbfper_fwd_kernel<<<grid, 256, 0, stream>>>(n1, n2, t1, t2, hyp, K, ldk);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_rbfper_bwd(int n1, int n2, const double* t1, const double* t2, const double* hyp, const double* G,
                               long ldg, double* out4, double* dt2, cudaStream_t stream) {
  if (n1 < 0 || n2 < 0) return NPGP_EINVAL;
  if (n1 == 0 || n2 == 0) return NPGP_OK;
  if (!t1 || !t2 || !hyp || !G || !out4 || ldg < n2) return NPGP_EINVAL;
  const int cb = ceil_div(n2, 256);
  long row_ctas = (kNumSMs * 4 + cb - 1) / cb;
  long rpc = (n1 + row_ctas - 1) / row_ctas;
  if (rpc < 32) rpc = 32;
  dim3 grid(cb, ceil_div(n1, rpc));
  rbfper_bwd_kernel<<<grid, 256, 0, stream>>>(n1, n2, t1, t2, hyp, G, ldg, (int)rpc, out4, dt2);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
