// Doubly-stochastic variational inference (DSVI) kernels for the deep GP (reference models/dgps.py:53-111 via GPyTorch's
// DeepGPLayer.__call__ and DeepApproximateMLL(VariationalELBO)):
//
//   npgp_dsvi_sample[_bwd]   h = mu + sqrt(var) * eps  -- the marginal (diagonal) reparameterised sample that replaces
//                            Normal(mean, sqrt(variance)).rsample(); eps ~ N(0,1) from Philox4x32-10 keyed by
//                            (seed, global element index), so the draw does not depend on how samples / rows are sharded
//                            over ranks; eps can also be supplied (parity tests) or returned.
//   npgp_gauss_ell_batched   per-sample expected log-likelihood sums over rows, sum_i E_q log N(y_i | f_si, s2), with
//                            warp-shuffle + block reductions, and the gradient seeds d/dmu, d/dvar.
#include "common.cuh"

namespace npgp {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t* out) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// one standard normal per element index (Box-Muller on two 32+21-bit uniforms)
__device__ __forceinline__ double normal_from_index(unsigned long long seed, unsigned long long idx) {
  uint32_t r[4];
  philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double u1 = ((double)r[0] * 4294967296.0 + (double)r[1] + 0.5) * (1.0 / 18446744073709551616.0);  // (0,1)
  const double u2 = ((double)r[2] * 4294967296.0 + (double)r[3] + 0.5) * (1.0 / 18446744073709551616.0);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__global__ void dsvi_sample_kernel(long n, const double* __restrict__ mu, const double* __restrict__ var,
                                   const double* __restrict__ eps_in, unsigned long long seed,
                                   unsigned long long index_offset, double* __restrict__ h,
                                   double* __restrict__ eps_out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double e = eps_in ? eps_in[i] : normal_from_index(seed, index_offset + (unsigned long long)i);
  h[i] = fma(sqrt(var[i]), e, mu[i]);
  if (eps_out) eps_out[i] = e;
}

__global__ void dsvi_sample_bwd_kernel(long n, const double* __restrict__ var, const double* __restrict__ eps,
                                       const double* __restrict__ dh, double* __restrict__ dmu,
                                       double* __restrict__ dvar) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dmu[i] = dh[i];
  dvar[i] = 0.5 * dh[i] * eps[i] * rsqrt(var[i]);
}

// grid (ceil(n/256), S).  sums[s] += sum_i ell_si;  gmu, gvar (S,n) = wscale * d ell / d mu, d var
__global__ void __launch_bounds__(256) gauss_ell_batched_kernel(int n, const double* __restrict__ y,
                                                                const double* __restrict__ mu,
                                                                const double* __restrict__ var,
                                                                const double* __restrict__ noise_p, double wscale,
                                                                double* __restrict__ sums, double* __restrict__ sq,
                                                                double* __restrict__ gmu, double* __restrict__ gvar) {
  __shared__ double red[32];
  const int i = blockIdx.x * 256 + threadIdx.x, s = blockIdx.y;
  const double s2 = *noise_p;
  double e = 0.0, r2v = 0.0;
  if (i < n) {
    const long k = (long)s * n + i;
    const double r = y[i] - mu[k], v = var[k];
    r2v = r * r + v;
    e = -0.5 * (r2v / s2 + log(s2) + 1.8378770664093453);
    if (gmu) gmu[k] = wscale * r / s2;
    if (gvar) gvar[k] = -0.5 * wscale / s2;
  }
  double t = block_sum(e, red);
  if (threadIdx.x == 0) atomicAdd(&sums[s], t);
  if (sq) {
    t = block_sum(r2v, red);
    if (threadIdx.x == 0) atomicAdd(&sq[s], t);
  }
}

}  // namespace npgp

using namespace npgp;

extern "C" int npgp_dsvi_sample(long n, const double* mu, const double* var, const double* eps_in,
                                unsigned long long seed, unsigned long long index_offset, double* h, double* eps_out,
                                cudaStream_t stream) {
  if (n < 0) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  if (!mu || !var || !h) return NPGP_EINVAL;
  dsvi_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, mu, var, eps_in, seed, index_offset, h, eps_out);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_dsvi_sample_bwd(long n, const double* var, const double* eps, const double* dh, double* dmu,
                                    double* dvar, cudaStream_t stream) {
  if (n < 0) return NPGP_EINVAL;
  if (n == 0) return NPGP_OK;
  if (!var || !eps || !dh || !dmu || !dvar) return NPGP_EINVAL;
  dsvi_sample_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, var, eps, dh, dmu, dvar);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}

extern "C" int npgp_gauss_ell_batched(int S, int n, const double* y, const double* mu, const double* var,
                                      const double* noise, double wscale, double* sums, double* sq, double* gmu,
                                      double* gvar, cudaStream_t stream) {
  if (S < 0 || n < 0) return NPGP_EINVAL;
  if (S == 0 || n == 0) return NPGP_OK;
  if (!y || !mu || !var || !noise || !sums) return NPGP_EINVAL;
  dim3 grid(ceil_div(n, 256), S);
  gauss_ell_batched_kernel<<<grid, 256, 0, stream>>>(n, y, mu, var, noise, wscale, sums, sq, gmu, gvar);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
