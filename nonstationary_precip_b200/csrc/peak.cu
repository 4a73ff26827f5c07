// FP64 peak probes (MEASURED_PEAKS.json has HBM and bf16 only): a register-resident DFMA loop and a DMMA.8x8x4 loop.
// Used by bench.py / tools to report the measured FP64 ceiling next to the nominal one.
#include "common.cuh"

namespace npgp {

__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double seed, double* out) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;  // keep the loop alive
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double seed, double* out) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = seed + threadIdx.x * 1e-3, b = seed - threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

}  // namespace npgp

// mode 0: DFMA, mode 1: DMMA.  Launches `blocks` CTAs of 256 threads running `iters` iterations of 16 independent ops.
// FLOPs launched: mode 0: blocks*256*iters*16*2;  mode 1: blocks*8(warps)*iters*16*(8*8*4*2).
extern "C" int npgp_fp64_peak_probe(int mode, int blocks, int iters, double* out, cudaStream_t stream) {
  if (!out || blocks <= 0 || iters <= 0) return NPGP_EINVAL;
  if (mode == 0) npgp::dfma_peak_kernel<<<blocks, 256, 0, stream>>>(iters, 1.0, out);
  else npgp::dmma_peak_kernel<<<blocks, 256, 0, stream>>>(iters, 1.0, out);
  NPGP_LAUNCH_CHECK();
  return NPGP_OK;
}
