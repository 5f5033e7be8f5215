// Per-phase clock64 profile of the blocked QR kernel: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
//   -I../../nd4js_b200/csrc -o qr_phase_prof qr_phase_prof.cu ; ./qr_phase_prof [variant] [batch]
#define QB_PROFILE
#include "../../nd4js_b200/csrc/qr_blocked.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const long long batch = argc > 2 ? atoll(argv[2]) : 65536;
  std::vector<double> h(batch * 2048);
  unsigned long long z = 88172645463325252ull;
  for (auto& x : h) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; x = (double)(z >> 11) / 9007199254740992.0 * 2 - 1; }
  double *A, *Q, *R;
  cudaMalloc(&A, h.size() * 8); cudaMalloc(&Q, h.size() * 8); cudaMalloc(&R, batch * 1024 * 8);
  cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  nd4b::launch_qr64x32_blocked(0, A, Q, R, batch, variant);
  cudaDeviceSynchronize();
  unsigned long long zero[16] = {0};
  cudaMemcpyToSymbol(qb_prof, zero, sizeof zero);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  nd4b::launch_qr64x32_blocked(0, A, Q, R, batch, variant);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long p[16];
  cudaMemcpyFromSymbol(p, qb_prof, sizeof p);
  const char* name[] = {"load+scale", "(unused)", "panel steps", "R store + V publish", "T (G + nilpotent product)",
                        "trailing update + R store", "Q phase", "Q store"};
  unsigned long long tot = 0;
  for (int i = 0; i < 8; i++) tot += p[i];
  printf("variant %d batch %lld: %.4f ms (with profiling hooks), %s\n", variant, batch, ms, cudaGetErrorString(cudaGetLastError()));
  for (int i = 0; i < 8; i++) printf("  %-28s %9.0f cycles per matrix  %5.1f%%\n", name[i], (double)p[i] / batch, 100.0 * p[i] / tot);
  printf("  %-28s %9.0f cycles per matrix\n", "total", (double)tot / batch);
  return 0;
}
