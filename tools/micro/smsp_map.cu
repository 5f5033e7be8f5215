// Which warps of a CTA share an SM sub-partition (scheduler + FP64 pipe)?  Warp 0 times a dependent DFMA chain while exactly
// one other warp k streams DMMAs: the chain slows from 8.4 to ~41 cycles per instruction only if k sits on warp 0's sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smsp_map smsp_map.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n, int partner) {
  const int warp = threadIdx.x >> 5;
  double x = out[threadIdx.x & 31], y = out[32 + (threadIdx.x & 31)], z = 1.0000001;
  if (warp == 0) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
      for (int u = 0; u < 16; u++) x = fma(x, z, y);
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[0] = t1 - t0;
  } else if (warp == partner) {
    double c[8][2];
    for (int j = 0; j < 8; j++) { c[j][0] = x; c[j][1] = y; }
#pragma unroll 1
    for (int i = 0; i < 4 * n; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(z), "d"(y));
    }
    for (int j = 0; j < 8; j++) x += c[j][0] + c[j][1];
  }
  out[threadIdx.x & 63] = x;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 16 * 8);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + i * 1e-3;
  cudaMemcpy(out, h, sizeof h, cudaMemcpyHostToDevice);
  const int n = 2000;
  for (int warps : {8, 12, 16}) {
    printf("CTA of %2d warps, DMMA partner k -> cycles per dependent DFMA of warp 0:", warps);
    for (int p = 1; p < warps; p++) {
      k<<<1, 32 * warps>>>(out, cyc, n, p);
      long long v; cudaMemcpy(&v, cyc, sizeof v, cudaMemcpyDeviceToHost);
      printf(" %d:%.0f", p, (double)v / (16.0 * n));
    }
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
