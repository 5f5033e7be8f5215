// probes.cu — FP64 roofline probes.  MEASURED_PEAKS.json has no fp64 entry, so the FP64 denominators
// (DFMA vector pipe and DMMA.8x8x4 tensor pipe) are measured with these two kernels on the same GPU,
// in the same run, as the numbers they normalise (tools/fp64_peak.py, bench.py).
#include "common.cuh"
#include "kernels.h"

namespace nd4b {

// 16 independent FMA chains per thread; 2*16*iters flop per thread.
__global__ void probe_dfma_kernel(double* out, int iters) {
  double x[16];
  const double a = 1.0000001, b = 1e-9 * (threadIdx.x + 1);
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = (double)(i + threadIdx.x);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

// 8 independent accumulator tiles per warp; 512 flop per DMMA.8x8x4.
__global__ void probe_dmma_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x + 1);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

cudaError_t launch_probe_dfma(cudaStream_t s, double* out, int iters, int blocks, int threads) {
  probe_dfma_kernel<<<blocks, threads, 0, s>>>(out, iters);
  return cudaGetLastError();
}
cudaError_t launch_probe_dmma(cudaStream_t s, double* out, int iters, int blocks, int threads) {
  probe_dmma_kernel<<<blocks, threads, 0, s>>>(out, iters);
  return cudaGetLastError();
}

}  // namespace nd4b
