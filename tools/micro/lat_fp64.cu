// Dependent-issue latencies on sm_100a (one warp, one block): DFMA, DADD, DMUL, DMMA chain, SHFL(64-bit), LDS.128, MUFU.RSQ64H.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_fp64 lat_fp64.cu && ./lat_fp64
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n) {
  __shared__ double sm[64];
  double x = out[threadIdx.x], y = out[32 + threadIdx.x], z = 1.0000001;
  sm[threadIdx.x] = x; sm[threadIdx.x + 32] = y;
  long long t0, t1;
  // DFMA
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) x = fma(x, z, y);
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) x = x + y;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) x = x * z;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // DMMA dependent on accumulator
  double c0 = x, c1 = y;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(z), "d"(y));
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // DMMA dependent through the A operand (accumulator -> A fragment)
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      double d0 = 0, d1 = 0;
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(c0), "d"(y));
      c0 = d0;
    }
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // 64-bit shuffle
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) x = __shfl_xor_sync(0xffffffffu, x, 1);
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // LDS dependent (address from data)
  int idx = threadIdx.x & 31;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) idx = (int)sm[idx & 63] & 63;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // rsqrt (full) and approx
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) y = rsqrt(y);
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(y));
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) y = 1.0 / y;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
  // two independent DFMA chains (issue rate of one warp)
  double x2 = y;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) { x = fma(x, z, y); x2 = fma(x2, z, y); }
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[10] = t1 - t0;
  double x3 = z, x4 = c1;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) { x = fma(x, z, y); x2 = fma(x2, z, y); x3 = fma(x3, z, y); x4 = fma(x4, z, y); }
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[11] = t1 - t0;
  out[threadIdx.x] = x + y + c0 + c1 + idx + x2 + x3 + x4;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 16 * 8);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + i * 1e-3;
  cudaMemcpy(out, h, sizeof h, cudaMemcpyHostToDevice);
  const int n = 1000;
  k<<<1, 32>>>(out, cyc, n); k<<<1, 32>>>(out, cyc, n);
  long long c[16]; cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
  const char* name[] = {"DFMA dep", "DADD dep", "DMUL dep", "DMMA dep(acc)", "DMMA dep(A)", "SHFL64 dep", "LDS dep (+cvt)", "rsqrt()", "rsqrt.approx", "1.0/y", "DFMA 2 chains (per pair)", "DFMA 4 chains (per quad)"};
  for (int i = 0; i < 12; i++) printf("%-28s %.2f cycles\n", name[i], (double)c[i] / (16.0 * n));
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
