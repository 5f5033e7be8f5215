// Dependent DFMA latency of a timed warp while other warps of the SM use the FP64 pipe (sm_100a).
// mode 0: only timed warps.  mode 1: partner warps stream independent DMMA.8x8x4.  mode 2: partner warps stream independent DFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_contend lat_contend.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n, int timed_warps, int mode) {
  const int warp = threadIdx.x >> 5;
  double x = out[threadIdx.x & 31], y = out[32 + (threadIdx.x & 31)], z = 1.0000001;
  if (warp < timed_warps) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
      for (int u = 0; u < 16; u++) x = fma(x, z, y);
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
  } else if (mode == 1) {
    double c[8][2];
    for (int j = 0; j < 8; j++) { c[j][0] = x; c[j][1] = y; }
#pragma unroll 1
    for (int i = 0; i < 4 * n; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(z), "d"(y));
    }
    for (int j = 0; j < 8; j++) x += c[j][0] + c[j][1];
  } else if (mode == 2) {
    double c[8];
    for (int j = 0; j < 8; j++) c[j] = x + j;
#pragma unroll 1
    for (int i = 0; i < 8 * n; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++) c[j] = fma(c[j], z, y);
    }
    for (int j = 0; j < 8; j++) x += c[j];
  }
  out[threadIdx.x & 63] = x;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 16 * 8);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + i * 1e-3;
  cudaMemcpy(out, h, sizeof h, cudaMemcpyHostToDevice);
  const int n = 2000;
  struct { int warps, timed, mode; const char* what; } cfg[] = {
    {1, 1, 0, "1 warp alone"}, {4, 4, 0, "4 timed warps (one per SMSP)"}, {8, 8, 0, "8 timed warps (two per SMSP)"},
    {8, 4, 1, "4 timed + 4 DMMA streams (one per SMSP)"}, {12, 4, 1, "4 timed + 8 DMMA streams"},
    {8, 4, 2, "4 timed + 4 DFMA streams"}, {5, 4, 1, "4 timed + 1 DMMA stream (SMSP 0 only)"}, {12, 8, 1, "8 timed + 4 DMMA streams"}};
  for (auto& c : cfg) {
    k<<<1, 32 * c.warps>>>(out, cyc, n, c.timed, c.mode);
    long long v[16]; cudaMemcpy(v, cyc, sizeof v, cudaMemcpyDeviceToHost);
    printf("%-45s", c.what);
    for (int w = 0; w < c.timed; w++) printf(" %.1f", (double)v[w] / (16.0 * n));
    printf("  cycles per dependent DFMA\n");
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
