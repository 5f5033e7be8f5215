// svd_solve.cu — nd.la.svd_rank / svd_lstsq / svd_solve (nd4js src/la/svd.js:31-58, 61-100, 103-226) on the device.
//
//  * svd_lstsq_kernel : x = V^T diag(1/sv[:R]) U^T y with the reference's rank cut R (first r with |sv_r| <= sqrt(eps) |sv_0|).
//                       One CTA per result matrix; U[N,M], sv[M], V[M,I], y[N,J] broadcast independently over the leading
//                       dims (svd.js:128-147, 201-218).  Every entry keeps the reference's own operation sequence
//                       (svd.js:177-193): tmp_ij = sum_k U_ki y_kj, k ascending from 0, product and sum rounded
//                       separately; tmp_ij / sv_i (IEEE); x_ij = sum_{k<R} V_ki tmp_kj likewise — the result is
//                       bit-identical with the reference.  Threads own (i, j) entries with i fastest, so that the reads of
//                       U and V rows are coalesced; tmp lives in shared memory.
//  * svd_rank_kernel  : R per sv vector (svd.js:31-58) with the reference's finiteness check (only entries before the cut).
#include "common.cuh"
#include "kernels.h"

namespace nd4b {

constexpr double kSqrtEps = 1.4901161193847656e-08;  // Math.sqrt(Number.EPSILON) = 2^-26, exact

__device__ __forceinline__ int64_t decode1(const BatchMap4& map, int o, int64_t m_local) {
  if (map.lin[o] >= 0) return m_local * map.lin[o];
  int64_t rem = map.base + m_local, off = 0;
  for (int d = map.nd - 1; d >= 0; d--) {
    const int64_t q = rem / map.size[d], r = rem - q * map.size[d];
    off += r * map.str[o][d];
    rem = q;
  }
  return off;
}

// rank cut of one sv vector; *bad is set when a non-finite entry is met before the cut (the reference throws there)
__device__ __forceinline__ int svd_rank_of(const double* sv, int M, bool* bad) {
  const double T = mul_rn(kSqrtEps, fabs(sv[0]));
  for (int r = 0; r < M; r++) {
    const double a = fabs(sv[r]);
    if (!(a <= 1.7976931348623157e308)) { *bad = true; return r; }   // NaN or Infinity
    if (a <= T) return r;
  }
  return M;
}

template <int T>
__global__ void __launch_bounds__(T)
svd_lstsq_kernel(const double* __restrict__ U, const double* __restrict__ SV, const double* __restrict__ V,
                 const double* __restrict__ Y, double* __restrict__ X, int64_t batch, int N, int M, int I, int J,
                 BatchMap4 map, int* fail) {
  extern __shared__ double sl_tmp[];   // [M][J]
  __shared__ int s_rank, s_bad;
  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const double* u = U + decode1(map, 0, m);
  const double* sv = SV + decode1(map, 1, m);
  const double* v = V + decode1(map, 2, m);
  const double* y = Y + decode1(map, 3, m);
  double* x = X + m * (int64_t)I * J;
  const int tid = threadIdx.x;
  // The rank cut (svd_rank_of: the first r with |sv_r| <= sqrt(eps) |sv_0|, a non-finite entry before it is an error) found by
  // all threads at once: one thread walking sv left the rest of the CTA waiting behind 64 dependent global loads.
  if (tid == 0) { s_rank = M; s_bad = M; }
  __syncthreads();
  {
    const double T0 = mul_rn(kSqrtEps, fabs(sv[0]));
    for (int r = tid; r < M; r += T) {
      const double a = fabs(sv[r]);
      if (!(a <= 1.7976931348623157e308)) atomicMin(&s_bad, r);   // NaN or Infinity
      else if (a <= T0) atomicMin(&s_rank, r);
    }
  }
  __syncthreads();
  if (s_bad < s_rank) {   // svd.js:168-169: the scan meets a non-finite value before the cut: the call throws; nothing of this matrix is defined
    if (tid == 0 && fail) atomicExch(fail, 1);
    return;
  }
  const int R = s_rank;
  // tmp = U^T y (rows < R), then / sv
  for (int e = tid; e < R * J; e += T) {
    const int j = e / R, i = e - j * R;   // i fastest: consecutive threads read consecutive U entries of a row
    double acc = 0.0;
#pragma unroll 8
    for (int k = 0; k < N; k++) acc = add_rn(acc, mul_rn(u[(int64_t)k * M + i], y[(int64_t)k * J + j]));
    sl_tmp[i * J + j] = acc / sv[i];
  }
  __syncthreads();
  for (int e = tid; e < I * J; e += T) {
    const int j = e / I, i = e - j * I;
    double acc = 0.0;   // x_dat starts zeroed (svd.js:151) and accumulates k ascending
#pragma unroll 8
    for (int k = 0; k < R; k++) acc = add_rn(acc, mul_rn(v[(int64_t)k * I + i], sl_tmp[k * J + j]));
    x[(int64_t)i * J + j] = acc;
  }
}

__global__ void svd_rank_kernel(const double* __restrict__ SV, int* __restrict__ rank, int64_t batch, int M, int* fail) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= batch) return;
  bool bad = false;
  rank[m] = svd_rank_of(SV + m * M, M, &bad);
  if (bad && fail) atomicExch(fail, 1);
}

cudaError_t launch_svd_lstsq(cudaStream_t s, const double* U, const double* SV, const double* V, const double* Y, double* X,
                             int64_t batch, int N, int M, int I, int J, const BatchMap4& map, int* fail) {
  const size_t smem = sizeof(double) * (size_t)M * J;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  const int work = (M > I ? M : I) * J;
  if (work <= 64) {
    cudaFuncSetAttribute(svd_lstsq_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    svd_lstsq_kernel<64><<<(unsigned)batch, 64, smem, s>>>(U, SV, V, Y, X, batch, N, M, I, J, map, fail);
  } else if (work <= 128) {
    cudaFuncSetAttribute(svd_lstsq_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    svd_lstsq_kernel<128><<<(unsigned)batch, 128, smem, s>>>(U, SV, V, Y, X, batch, N, M, I, J, map, fail);
  } else {
    cudaFuncSetAttribute(svd_lstsq_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    svd_lstsq_kernel<256><<<(unsigned)batch, 256, smem, s>>>(U, SV, V, Y, X, batch, N, M, I, J, map, fail);
  }
  return cudaGetLastError();
}

cudaError_t launch_svd_rank(cudaStream_t s, const double* SV, int* rank, int64_t batch, int M, int* fail) {
  svd_rank_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, s>>>(SV, rank, batch, M, fail);
  return cudaGetLastError();
}

}  // namespace nd4b
