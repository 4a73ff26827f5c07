// Cycle-level breakdown of the 64x64 diagonal-block factorisation used by potrf (csrc/chol.cu): one CTA, clock64 stamps
// around each phase of one 8-column iteration and around the doubling inverse.   Build + run (GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include \
//        -o tools/probes/chol_probe tools/probes/chol_probe.cu && tools/probes/chol_probe
#include <cstdio>
#include <vector>
#include "../../nonstationary_precip_b200/csrc/chol.cu"

long npgp_launch_counter = 0;  // normally defined in capi.cu
namespace npgp {  // normally dgemm.cu (the probe never takes the path that calls it)
int dgemm_impl(int, int, int, int, int, double, const double*, long, const double*, long, double, double*, long, int, int,
               int, cudaStream_t) { return -2; }
}

using namespace npgp;

__global__ void __launch_bounds__(CT) probe_kernel(const double* A, long long* stamps, double* out) {
  extern __shared__ double sm[];
  double *s = sm, *x = sm + TILE_SMEM, *tmp = sm + 2 * TILE_SMEM;
  double* sInv = tmp + 32 * 32;
  const int tid = threadIdx.x;
  load_tile(s, A, 64, 0, 0, 64, true);
  for (int e = tid; e < TB * TB; e += CT) x[(e >> 6) * TLD + (e & 63)] = 0.0;
  __syncthreads();
  int k = 0;
  auto stamp = [&]() {
    __syncthreads();
    if (tid == 0) stamps[k] = clock64();
    ++k;
  };
  int dummy = 0;
  stamp();
  if (tid == 0) chol8_serial(s, sInv, 0, 0, &dummy);
  stamp();  // 1: chol8
  if (tid < 8) inv8_column(s, sInv, x, 0, tid);
  stamp();  // 2: inv8
  {
    const int c0 = 0, r1 = 8, n = 56;
    if (tid < n) {
      const int r = r1 + tid;
      double l[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        double acc = s[r * TLD + c0 + c];
#pragma unroll
        for (int kk = 0; kk < c; ++kk) acc = fma(-l[kk], s[(c0 + c) * TLD + c0 + kk], acc);
        l[c] = acc * sInv[c0 + c];
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        s[r * TLD + c0 + c] = l[c];
        s[(c0 + c) * TLD + r] = 0.0;
      }
    }
    stamp();  // 3: panel
    for (int e = tid; e < n * n; e += CT) {
      const int i = r1 + e / n, kk = r1 + e % n;
      if (kk <= i) {
        double acc = s[i * TLD + kk];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc = fma(-s[i * TLD + c0 + c], s[kk * TLD + c0 + c], acc);
        s[i * TLD + kk] = acc;
      }
    }
    stamp();  // 4: trailing (all threads)
  }
  stamp();    // 5: empty (barrier + stamp overhead)
  // whole factorisation from scratch
  load_tile(s, A, 64, 0, 0, 64, true);
  stamp();    // 6: reload
  factor_invert_64(s, x, tmp, 0, &dummy);
  stamp();    // 7: factor_invert_64 total
  for (int e = tid; e < TB * TB; e += CT) out[e] = x[(e >> 6) * TLD + (e & 63)] + s[(e >> 6) * TLD + (e & 63)];
}

int main() {
  std::vector<double> h(64 * 64);
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) h[i * 64 + j] = (i == j ? 2.0 : 0.0) + 1.0 / (1.0 + (i > j ? i - j : j - i));
  double *A, *out;
  long long* st;
  cudaMalloc(&A, sizeof(double) * 4096);
  cudaMalloc(&out, sizeof(double) * 4096);
  cudaMalloc(&st, sizeof(long long) * 16);
  cudaMemcpy(A, h.data(), sizeof(double) * 4096, cudaMemcpyHostToDevice);
  const int smem = 3 * TILE_SMEM * (int)sizeof(double);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long hs[16];
  for (int rep = 0; rep < 3; ++rep) {
    probe_kernel<<<1, CT, smem>>>(A, st, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hs, st, sizeof(hs), cudaMemcpyDeviceToHost);
    const char* names[] = {"chol8_serial", "inv8_column", "panel(subst)", "trailing(256thr)", "empty", "reload", "factor_invert_64"};
    printf("rep %d:", rep);
    for (int k = 1; k <= 7; ++k) printf("  %s=%lld", names[k - 1], hs[k] - hs[k - 1]);
    printf("  (cycles)\n");
  }
  // back-to-back launches of one panel step (M = 1024): j = 0 (120 CTAs) and j = 14 (the diagonal CTA alone)
  {
    const int M = 1024, nblk = 16;
    double *Am, *Lm, *Pm;
    int* info;
    cudaMalloc(&Am, sizeof(double) * M * M);
    cudaMalloc(&Lm, sizeof(double) * M * M);
    cudaMalloc(&Pm, sizeof(double) * M * M);
    cudaMalloc(&info, sizeof(int));
    std::vector<double> hm((size_t)M * M);
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < M; ++j) hm[(size_t)i * M + j] = (i == j ? 4.0 : 0.0) + 1.0 / (1.0 + (i > j ? i - j : j - i));
    cudaFuncSetAttribute(potrf_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STEP_SMEM);
    cudaFuncSetAttribute(potrf_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FIRST_SMEM);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int js[2] = {0, 14};
    for (int v = 0; v < 2; ++v) {
      const int j = js[v], r = nblk - 1 - j, reps = 50;
      cudaMemcpy(Am, hm.data(), sizeof(double) * M * M, cudaMemcpyHostToDevice);
      cudaMemset(Pm, 0, sizeof(double) * M * M);
      cudaMemset(info, 0, sizeof(int));
      potrf_first_kernel<<<1, CT, FIRST_SMEM>>>(M, Am, M, Lm, M, Pm, M, info);
      for (int w = 0; w < 3; ++w) potrf_step_kernel<<<r * (r + 1) / 2, CT, STEP_SMEM>>>(M, nblk, j, Am, M, Lm, M, Pm, M, info);
      cudaEventRecord(e0);
      for (int w = 0; w < reps; ++w)
        potrf_step_kernel<<<r * (r + 1) / 2, CT, STEP_SMEM>>>(M, nblk, j, Am, M, Lm, M, Pm, M, info);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("potrf_step_kernel j=%d (%d CTAs): %.2f us per launch, back to back\n", j, r * (r + 1) / 2, 1e3 * ms / reps);
    }
    cudaEventRecord(e0);
    for (int w = 0; w < 50; ++w) potrf_first_kernel<<<1, CT, FIRST_SMEM>>>(M, Am, M, Lm, M, Pm, M, info);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("potrf_first_kernel: %.2f us per launch\n", 1e3 * ms / 50);
  }
  return 0;
}
