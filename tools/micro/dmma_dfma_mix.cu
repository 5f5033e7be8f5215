// Are the DFMA (vector FP64) and DMMA.8x8x4 (tensor FP64) pipes independent on B200?  Runs DFMA-only, DMMA-only and a
// 1:1 flop mix; if the mix sustains more than either alone, the two can overlap.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NF, int NM>
__global__ void k(double* out, int iters) {
  double x[16], c[8][2];
  for (int i = 0; i < 16; i++) x[i] = i + threadIdx.x;
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0;
  const double a = 1.0000001, b = 1e-9 * (threadIdx.x + 1);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < NF; r++)
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
#pragma unroll
    for (int r = 0; r < NM; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
  for (int i = 0; i < 16; i++) s += x[i];
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}
template <int NF, int NM> void run(const char* name, double* out) {
  const int iters = 4096, blocks = 148 * 4, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NF, NM><<<blocks, threads>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<NF, NM><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fl_f = 2.0 * 16 * NF * (double)iters * blocks * threads, fl_m = 512.0 * 8 * NM * (double)iters * blocks * (threads / 32);
  printf("%-22s %.3f ms  dfma %.1f TF + dmma %.1f TF = %.1f TF\n", name, ms, fl_f / ms / 1e9, fl_m / ms / 1e9, (fl_f + fl_m) / ms / 1e9);
}
int main() {
  double* out; cudaMalloc(&out, 8);
  run<4, 0>("DFMA only", out);
  run<0, 1>("DMMA only", out);
  run<4, 1>("mix (1:1 flop)", out);   // 4*16*2*32 = 4096 flop/warp vs 8*512 = 4096 flop/warp per iteration
  run<8, 1>("mix (2:1 flop)", out);
  return 0;
}
