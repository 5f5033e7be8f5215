// Microbenchmark: cost of warp-uniform (broadcast) shared-memory loads vs distinct-address loads, and of SHFL.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bcast lds_bcast.cu && ./lds_bcast
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_uniform128(double* out, int iters) {
  __shared__ double2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  double a = 0, b = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 32; j++) { double2 v = sm[(it + j) & 1023]; a += v.x; b += v.y; }
  }
  if (a + b == 1.2345) out[0] = a;
}
__global__ void k_distinct128(double* out, int iters) {
  __shared__ double2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  double a = 0, b = 0;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 32; j++) { double2 v = sm[(it + 32 * j + lane) & 1023]; a += v.x; b += v.y; }
  }
  if (a + b == 1.2345) out[0] = a;
}
__global__ void k_uniform64(double* out, int iters) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double a = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 32; j++) a += sm[(it + j) & 1023];
  }
  if (a == 1.2345) out[0] = a;
}
__global__ void k_shfl(double* out, int iters) {
  double a = threadIdx.x, s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 32; j++) s += __shfl_sync(0xffffffffu, a + j, (it + j) & 31);
  }
  if (s == 1.2345) out[0] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 8);
  const int iters = 2000, blocks = 148 * 4, threads = 256;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  auto report = [&](const char* name, float ms) {
    // instructions per SM: blocks/148 * warps * iters*32 ; cycles = ms*clk
    double instr_per_sm = (double)blocks / 148 * (threads / 32) * iters * 32.0;
    printf("%-14s %.3f ms  -> %.2f clk per warp-instruction per SM (at %d kHz nominal)\n", name, ms, ms * 1e-3 * clk * 1e3 / instr_per_sm, clk);
  };
  report("uniform LDS128", timeit([&] { k_uniform128<<<blocks, threads>>>(out, iters); }));
  report("distinct LDS128", timeit([&] { k_distinct128<<<blocks, threads>>>(out, iters); }));
  report("uniform LDS64", timeit([&] { k_uniform64<<<blocks, threads>>>(out, iters); }));
  report("SHFL64", timeit([&] { k_shfl<<<blocks, threads>>>(out, iters); }));
  return 0;
}
