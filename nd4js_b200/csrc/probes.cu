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

// Self-check of the branch-free division / square-root helpers (common.cuh) against the compiler's own `/` and sqrt():
// every thread draws `per_thread` doubles from a counter-based generator that covers all exponents (and, every 16th
// draw, values next to the fast-path range limits), and counts results that differ in any bit from the IEEE operation
// although the helper reported its fast path as valid.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__global__ void selfcheck_kernel(long long per_thread, unsigned long long seed, unsigned long long* out) {
  const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long bad_sqrt = 0, bad_div = 0, fast_sqrt = 0, fast_div = 0;
  for (long long i = 0; i < per_thread; i++) {
    const unsigned long long ctr = (tid * (unsigned long long)per_thread + i) * 2 + seed;
    unsigned long long xa = mix64(ctr), xb = mix64(ctr + 1);
    if ((i & 15) == 15) {  // hug the range limits of the fast paths
      xa = (xa & 0x000fffffffffffffull) | ((0x035ull + (xa >> 61)) << 52);          // 2^-970 .. 2^-963
      xb = (xb & 0x800fffffffffffffull) | ((0x7fcull + ((xb >> 61) & 3)) << 52);    // 2^1021 .. Inf/NaN
    }
    if ((i & 31) == 7) xa &= 0x8000000000000000ull;                                   // +-0 numerators
    const double a = __longlong_as_double((long long)xa), b = __longlong_as_double((long long)xb);
    {
      bool ok = true;
      const double r = sqrt_fast(fabs(a), ok);
      const double w = sqrt(fabs(a));
      if (ok) { fast_sqrt++; if (__double_as_longlong(r) != __double_as_longlong(w)) bad_sqrt++; }
    }
    {
      bool ok = true;
      const ColRecip rc = col_recip(b);
      const double q = div_col(a, rc, ok);
      const double w = a / b;
      if (ok) { fast_div++; if (__double_as_longlong(q) != __double_as_longlong(w) && !(isnan(q) && isnan(w))) bad_div++; }
    }
  }
  atomicAdd(out + 0, bad_sqrt);
  atomicAdd(out + 1, bad_div);
  atomicAdd(out + 2, fast_sqrt);
  atomicAdd(out + 3, fast_div);
}

cudaError_t launch_selfcheck(cudaStream_t s, long long per_thread, unsigned long long seed, unsigned long long* out,
                             int blocks, int threads) {
  selfcheck_kernel<<<blocks, threads, 0, s>>>(per_thread, seed, out);
  return cudaGetLastError();
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
