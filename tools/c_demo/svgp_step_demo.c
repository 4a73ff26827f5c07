/* A C caller of the whole SVGP-Gibbs ELBO step: no Python, no torch.  Reads a problem written by the test
 * (tests/test_c_caller_gpu.py) or generates one, runs `iters` calls of npgp_svgp_step on the default stream, writes the
 * flat gradient of the last step and the per-step losses.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include tools/c_demo/svgp_step_demo.c -o svgp_step_demo \
 *       -L nonstationary_precip_b200 -lnpgp -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/nonstationary_precip_b200
 *   ./svgp_step_demo problem.bin out.bin 3
 *
 * problem.bin: int32 header [variant, d, M, B, N_total, learn_z, include_prior, iters], then doubles:
 *   x (B*d), y (B), theta (n_pad), mask (n_pad), consts (full: row_os (1), row_lam (d); diag: prior_c (d), prior_os (d),
 *   prior_lam (d*d)).  out.bin: doubles loss[iters], grad (n_pad + 2), theta (n_pad). */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "npgp.h"

#define CK(call)                                                                      \
  do {                                                                                \
    int rc__ = (int)(call);                                                           \
    if (rc__ != 0) {                                                                  \
      fprintf(stderr, "%s:%d: %s failed with %d\n", __FILE__, __LINE__, #call, rc__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

static double* to_device(const double* h, long n) {
  double* d = NULL;
  CK(cudaMalloc((void**)&d, sizeof(double) * (n > 0 ? n : 1)));
  if (n > 0) CK(cudaMemcpy(d, h, sizeof(double) * n, cudaMemcpyHostToDevice));
  return d;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s problem.bin out.bin\n", argv[0]);
    return 2;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[8];
  if (fread(hdr, sizeof(int), 8, f) != 8) return 2;
  const int variant = hdr[0], d = hdr[1], M = hdr[2], B = hdr[3], iters = hdr[7];
  npgp_svgp_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.variant = variant, cfg.d = d, cfg.M = M, cfg.B_local = B, cfg.N_total = hdr[4], cfg.B_global = B, cfg.world_size = 1;
  cfg.jitter_zz = 1e-6, cfg.jitter_xx = 1e-4, cfg.kernel_jitter = 1e-5, cfg.min_var = 1e-6;
  cfg.learn_z = hdr[5], cfg.include_prior = hdr[6];
  const long n_pad = npgp_svgp_theta_size(&cfg);
  if (n_pad <= 0) return 3;
  const long n_const = variant == 1 ? 1 + d : 2 * d + d * d;
  const long n_in = (long)B * d + B + 2 * n_pad + n_const;
  double* h = (double*)malloc(sizeof(double) * n_in);
  if (fread(h, sizeof(double), n_in, f) != (size_t)n_in) return 2;
  fclose(f);
  const double *hx = h, *hy = hx + (long)B * d, *htheta = hy + B, *hmask = htheta + n_pad, *hconst = hmask + n_pad;
  double *x = to_device(hx, (long)B * d), *y = to_device(hy, B), *theta = to_device(htheta, n_pad), *mask = to_device(hmask, n_pad);
  double* consts = to_device(hconst, n_const);
  if (variant == 1) cfg.row_os = consts, cfg.row_lam = consts + 1;
  else cfg.prior_c = consts, cfg.prior_os = consts + d, cfg.prior_lam = consts + 2 * d;
  double *grad, *adam_m, *adam_v, *step;
  int* status;
  void* ws;
  const long ws_bytes = npgp_svgp_workspace_bytes(&cfg);
  CK(cudaMalloc((void**)&grad, sizeof(double) * (n_pad + 2)));
  CK(cudaMalloc((void**)&adam_m, sizeof(double) * n_pad));
  CK(cudaMalloc((void**)&adam_v, sizeof(double) * n_pad));
  CK(cudaMalloc((void**)&step, sizeof(double)));
  CK(cudaMalloc((void**)&status, sizeof(int)));
  CK(cudaMalloc(&ws, ws_bytes));
  CK(cudaMemset(adam_m, 0, sizeof(double) * n_pad));
  CK(cudaMemset(adam_v, 0, sizeof(double) * n_pad));
  CK(cudaMemset(step, 0, sizeof(double)));
  CK(cudaMemset(status, 0, sizeof(int)));
  npgp_svgp_plan* plan = NULL;
  CK(npgp_svgp_plan_create(&plan, &cfg, ws, ws_bytes));
  double* losses = (double*)malloc(sizeof(double) * iters);
  for (int it = 0; it < iters; ++it) {
    CK(npgp_svgp_step(plan, x, y, theta, grad, adam_m, adam_v, mask, step, status, 0.01, 0.9, 0.999, 1e-8, NULL, 0));
    CK(cudaMemcpy(&losses[it], grad + n_pad, sizeof(double), cudaMemcpyDeviceToHost));
  }
  int hstatus = -1;
  CK(cudaMemcpy(&hstatus, status, sizeof(int), cudaMemcpyDeviceToHost));
  double* hout = (double*)malloc(sizeof(double) * (2 * n_pad + 2));
  CK(cudaMemcpy(hout, grad, sizeof(double) * (n_pad + 2), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hout + n_pad + 2, theta, sizeof(double) * n_pad, cudaMemcpyDeviceToHost));
  f = fopen(argv[2], "wb");
  fwrite(losses, sizeof(double), iters, f);
  fwrite(hout, sizeof(double), 2 * n_pad + 2, f);
  fclose(f);
  CK(npgp_svgp_plan_destroy(plan));
  printf("npgp %d: %d steps, status %d, loss %.12f -> %.12f\n", npgp_version(), iters, hstatus, losses[0], losses[iters - 1]);
  return hstatus == 0 ? 0 : 4;
}
