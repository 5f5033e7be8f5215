// solve.cu — batched triangular solves with broadcast: tril_solve / triu_solve (nd4js src/la/tri.js:45-98,156-293)
// and cholesky_solve (src/la/cholesky.js:75-144 = _tril_solve followed by _tril_t_solve, tri.js:100-125).
//
// Every right-hand-side column is independent, and for one entry x_ij the reference performs a fixed sequence:
// start from y_ij, subtract t_ik * x_kj for k ascending (forward) or descending (backward), each product and
// difference rounded separately, then divide by the diagonal.  One thread per (matrix, rhs column) replays exactly
// that sequence (__dmul_rn/__dsub_rn, IEEE division), so X is bit-identical to the reference's.
// Threads of a warp take consecutive rhs columns of one matrix: X/Y accesses are coalesced, T is a broadcast read.
#include "common.cuh"
#include "kernels.h"

namespace nd4b {

__device__ __forceinline__ void decode_batch2(const BatchMap& map, int64_t m_local, int64_t& ao, int64_t& bo) {
  ao = m_local * map.a_lin;
  bo = m_local * map.b_lin;
  if (map.a_lin >= 0 && map.b_lin >= 0) return;
  int64_t m = m_local + map.base, oa = 0, ob = 0;
#pragma unroll 1
  for (int d = map.nd - 1; d >= 0; d--) {
    const int64_t q = m / map.size[d];
    const int64_t idx = m - q * map.size[d];
    oa += idx * map.a_str[d];
    ob += idx * map.b_str[d];
    m = q;
  }
  if (map.a_lin < 0) ao = oa;
  if (map.b_lin < 0) bo = ob;
}

constexpr int kSolveThreads = 128;

// op 0: x = L^-1 y (forward), 1: x = U^-1 y (backward), 2: x = L^-T L^-1 y
__global__ void __launch_bounds__(kSolveThreads)
tri_solve_kernel(int op, const double* __restrict__ T, const double* __restrict__ Y, double* __restrict__ X,
                 int64_t batch, int M, int J, BatchMap map) {
  const int64_t e = (int64_t)blockIdx.x * kSolveThreads + threadIdx.x;
  if (e >= batch * J) return;
  const int64_t m = e / J;
  const int j = (int)(e - m * J);
  int64_t to, yo;
  decode_batch2(map, m, to, yo);
  const double* t = T + to;
  const double* y = Y + yo + j;
  double* x = X + m * (int64_t)M * J + j;

  if (op == 1) {
    for (int i = M; i-- > 0;) {
      double s = y[(int64_t)i * J];
      for (int k = M; --k > i;) s = sub_rn(s, mul_rn(t[(int64_t)i * M + k], x[(int64_t)k * J]));
      x[(int64_t)i * J] = s / t[(int64_t)i * M + i];
    }
    return;
  }
  for (int i = 0; i < M; i++) {
    double s = y[(int64_t)i * J];
    for (int k = 0; k < i; k++) s = sub_rn(s, mul_rn(t[(int64_t)i * M + k], x[(int64_t)k * J]));
    x[(int64_t)i * J] = s / t[(int64_t)i * M + i];
  }
  if (op == 2) {
    // _tril_t_solve: entry i receives x_i -= L[k][i] * x_k for k = M-1 .. i+1 (descending), then /= L[i][i]
    for (int i = M; i-- > 0;) {
      double s = x[(int64_t)i * J];
      for (int k = M; --k > i;) s = sub_rn(s, mul_rn(t[(int64_t)k * M + i], x[(int64_t)k * J]));
      x[(int64_t)i * J] = s / t[(int64_t)i * M + i];
    }
  }
}

cudaError_t launch_tri_solve(cudaStream_t s, int op, const double* T, const double* Y, double* X,
                             int64_t batch, int M, int J, const BatchMap& map) {
  if (batch <= 0) return cudaSuccess;
  const int64_t threads = batch * J;
  const int64_t grid = (threads + kSolveThreads - 1) / kSolveThreads;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  tri_solve_kernel<<<(unsigned)grid, kSolveThreads, 0, s>>>(op, T, Y, X, batch, M, J, map);
  return cudaGetLastError();
}

}  // namespace nd4b
