// Library identity for the C ABI (include/npgp.h).
#include "common.cuh"

extern "C" int npgp_version(void) { return 100; }  // 0.1.0

long npgp_launch_counter = 0;
extern "C" long npgp_launch_count(void) { return npgp_launch_counter; }

// Measurement aid: one thread writes the GPU's nanosecond timer into buf[slot] when the stream reaches this point.  Unlike a
// CUDA event it can be captured in a graph, so the real timeline of a replayed multi-stream step can be read back.
__global__ void npgp_timestamp_kernel(long long* buf, int slot) {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  buf[slot] = t;
}
extern "C" int npgp_timestamp(long long* buf, int slot, cudaStream_t stream) {
  if (!buf || slot < 0) return NPGP_EINVAL;
  npgp_timestamp_kernel<<<1, 1, 0, stream>>>(buf, slot);
  cudaError_t e__ = cudaGetLastError();  // (not counted in npgp_launch_count: diagnostics only)
  return e__ == cudaSuccess ? NPGP_OK : (int)e__;
}
