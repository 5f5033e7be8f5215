// solve.cu — batched triangular solves with broadcast: tril_solve / triu_solve (nd4js src/la/tri.js:45-98,156-293)
// and cholesky_solve (src/la/cholesky.js:75-144 = _tril_solve followed by _tril_t_solve, tri.js:100-125).
//
// Every right-hand-side column is independent, and for one entry x_ij the reference performs a fixed sequence:
// start from y_ij, subtract t_ik * x_kj for k ascending (forward) or descending (backward), each product and
// difference rounded separately, then divide by the diagonal.  One thread per (matrix, rhs column) replays exactly
// that sequence (__dmul_rn/__dsub_rn, IEEE division), so X is bit-identical to the reference's.
// Threads of a warp take consecutive rhs columns of one matrix: X/Y accesses are coalesced, T is a broadcast read.
// That generic kernel serves M < 16 and M > 64; trisolve16_kernel (M = 16, the solves that follow C3) and
// trisolve_warp_kernel (16 < M <= 64) keep the triangle in shared memory, and qr_lstsq32_kernel fuses Q^T y with
// the back substitution — all with the same per-entry sequences, i.e. bit-identical results.
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

// ------------------------------------------------------------------------------------------------
// M = 16 (the solves that follow C3): 4 threads per matrix, 8 matrices per warp, like chol16_kernel.  The referenced
// triangles of the warp's 8 matrices are staged in shared memory in packed form (136 entries + the 16 hoisted
// reciprocals of the diagonal; matrix stride 156 doubles: the rows a quad reads at once and the matrices of a warp fall
// into distinct banks) — 1.2 KB per matrix, so that 20 warps = 160 matrices are in flight per SM.  Thread t owns rows
// t, t+4, t+8, t+12 of the right-hand side.  The substitution runs right-looking: as soon as x_k is known every row
// subtracts t_ik * x_k, so the rows proceed in parallel while each entry still sees the reference's own sequence —
// k ascending (forward) or descending (backward), product and difference rounded separately, IEEE division (hoisted
// reciprocal of div.rn's fast path, common.cuh) — and X stays bit-identical.  JB right-hand sides are carried at once.
// ------------------------------------------------------------------------------------------------
constexpr int kTs16Warps = 4;
constexpr int kTsMS = 156;   // doubles per matrix: 136 packed triangle entries, 16 reciprocals, 4 pad (156 = 12 mod 16)
constexpr size_t kTs16Smem = sizeof(double) * kTs16Warps * 8 * kTsMS;

// packed position of T[i][k]: lower triangle row-major (k <= i), or upper triangle row-major (k >= i)
template <bool UPPER>
__device__ __forceinline__ constexpr int ts_idx(int i, int k) {
  return UPPER ? 16 * i - i * (i - 1) / 2 + (k - i) : i * (i + 1) / 2 + k;
}

__device__ __noinline__ double ts_ieee_div(double a, double b) { return a / b; }  // rare path, out of line

__device__ __forceinline__ bool recip_range_ok(double b) {   // ColRecip::bnorm
  return (((unsigned)__double2hiint(b) & 0x7fffffffu) - 0x01700000u) < 0x7d000000u;
}

// x_k = s_k / t_kk for JB right-hand sides: the owner's partial sums are fetched from its lane, the quotient is formed from
// the precomputed reciprocal (bit-identical with `/` inside the fast-path ranges; `ok` is cleared otherwise and the
// caller redoes the matrix with ts16_slow_column).  No branch: the substitution is one basic block.
template <int JB>
__device__ __forceinline__ void ts_divide(double (&xk)[JB], const double (&sk)[JB], int src_lane, double tkk, double y2, bool& ok) {
  ColRecip rc;
  rc.b = tkk;
  rc.y = y2;
  rc.bhi = __int_as_float(__double2hiint(tkk));
  rc.bnorm = recip_range_ok(tkk);
#pragma unroll
  for (int jb = 0; jb < JB; jb++) xk[jb] = div_col(shfl(sk[jb], src_lane), rc, ok);
}

// Rare path of trisolve16_kernel: one right-hand-side column of one matrix, sequentially, in the reference's own order with
// the plain IEEE division (the sequences of tri_solve_kernel above), from the packed triangle in shared memory.
template <int OP>
__device__ __noinline__ void ts16_slow_column(const double* mine, const double* y, double* x, int J) {
  constexpr int N = 16;
  constexpr bool UPPER = (OP == 1);
  double xs[N];
  if (OP != 1) {
    for (int i = 0; i < N; i++) {
      double s = y[(int64_t)i * J];
      for (int k = 0; k < i; k++) s = sub_rn(s, mul_rn(mine[ts_idx<false>(i, k)], xs[k]));
      xs[i] = s / mine[ts_idx<false>(i, i)];
    }
  }
  if (OP != 0) {
    for (int i = N; i-- > 0;) {
      double s = (OP == 1) ? y[(int64_t)i * J] : xs[i];
      for (int k = N; --k > i;) s = sub_rn(s, mul_rn(UPPER ? mine[ts_idx<true>(i, k)] : mine[ts_idx<false>(k, i)], xs[k]));
      xs[i] = s / mine[ts_idx<UPPER>(i, i)];
    }
  }
  for (int i = 0; i < N; i++) x[(int64_t)i * J] = xs[i];
}

template <int OP, int JB>
__global__ void __launch_bounds__(kTs16Warps * 32, 4)
trisolve16_kernel(const double* __restrict__ T, const double* __restrict__ Y, double* __restrict__ X,
                  int64_t batch, int J, BatchMap map) {
  constexpr int N = 16;
  constexpr bool UPPER = (OP == 1);
  extern __shared__ __align__(16) double ts_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = lane & 3, q = lane >> 2, qbase = lane & ~3;
  double* tile = ts_smem + warp * 8 * kTsMS;
  const int64_t m0 = ((int64_t)blockIdx.x * kTs16Warps + warp) * 8;
  if (m0 >= batch) return;  // warp-uniform
  const int nmat = (int)min((int64_t)8, batch - m0);
  const bool valid = q < nmat;
  const int64_t m = m0 + (valid ? q : 0);
  int64_t to, yo;
  decode_batch2(map, m, to, yo);

  // stage the referenced triangle of the 8 matrices, packed (8-byte async copies: the rows of a triangle are short)
  {
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
#pragma unroll 1
    for (int mm = 0; mm < nmat; mm++) {
      const int64_t tmm = __shfl_sync(kFull, to, 4 * mm);
      const double* src = T + tmm;
#pragma unroll
      for (int e = lane; e < N * N; e += 32) {
        const int i = e >> 4, k = e & 15;
        if (UPPER ? (k >= i) : (k <= i))
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tile_s + (uint32_t)(mm * kTsMS + ts_idx<UPPER>(i, k)) * 8u), "l"(src + e) : "memory");
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
  }
  double* mine = tile + (valid ? q : 0) * kTsMS;
  // reciprocals of the diagonal, off the substitution chain
  if (valid) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int i = t + 4 * s;
      mine[136 + i] = col_recip(mine[ts_idx<UPPER>(i, i)]).y;
    }
  }
  __syncwarp();
  const double* y = Y + yo;
  double* x = X + m * (int64_t)(N * J);

#pragma unroll 1
  for (int j0 = 0; j0 < J; j0 += JB) {
    double sv[4][JB];
    bool ok = true;
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
      for (int jb = 0; jb < JB; jb++) sv[s][jb] = (j0 + jb < J) ? y[(int64_t)(t + 4 * s) * J + j0 + jb] : 0.0;

    if (OP != 1) {  // forward substitution with the lower triangle: k ascending
#pragma unroll
      for (int k = 0; k < N; k++) {
        double xk[JB];
        ts_divide<JB>(xk, sv[k >> 2], qbase | (k & 3), mine[ts_idx<false>(k, k)], mine[136 + k], ok);
#pragma unroll
        for (int s = k >> 2; s < 4; s++) {
          const int i = t + 4 * s;
          const double tik = mine[ts_idx<false>(max(i, k), k)];
#pragma unroll
          for (int jb = 0; jb < JB; jb++) {
            if (i > k) sv[s][jb] = sub_rn(sv[s][jb], mul_rn(tik, xk[jb]));
            else if (i == k) sv[s][jb] = xk[jb];
          }
        }
      }
    }
    if (OP != 0) {  // backward substitution, k descending: op 1 with the upper triangle U_ik, op 2 with L^T (t_ik = L_ki)
#pragma unroll
      for (int k = N - 1; k >= 0; k--) {
        double xk[JB];
        ts_divide<JB>(xk, sv[k >> 2], qbase | (k & 3), mine[ts_idx<UPPER>(k, k)], mine[136 + k], ok);
#pragma unroll
        for (int s = 0; s <= (k >> 2); s++) {
          const int i = t + 4 * s;
          const double tik = UPPER ? mine[ts_idx<true>(min(i, k), k)] : mine[ts_idx<false>(k, min(i, k))];
#pragma unroll
          for (int jb = 0; jb < JB; jb++) {
            if (i < k) sv[s][jb] = sub_rn(sv[s][jb], mul_rn(tik, xk[jb]));
            else if (i == k) sv[s][jb] = xk[jb];
          }
        }
      }
    }
    int bad = ok ? 0 : 1;   // quad-uniform already (the numerators are broadcasts), made explicit
    bad |= __shfl_xor_sync(kFull, bad, 1);
    bad |= __shfl_xor_sync(kFull, bad, 2);
    if (valid && !bad) {
#pragma unroll
      for (int s = 0; s < 4; s++)
#pragma unroll
        for (int jb = 0; jb < JB; jb++)
          if (j0 + jb < J) x[(int64_t)(t + 4 * s) * J + j0 + jb] = sv[s][jb];
    } else if (valid && t == 0) {
      for (int jb = 0; jb < JB && j0 + jb < J; jb++) ts16_slow_column<OP>(mine, y + j0 + jb, x + j0 + jb, J);
    }
  }
}

template <int OP, int JB>
static cudaError_t launch_trisolve16_op(cudaStream_t s, const double* T, const double* Y, double* X,
                                        int64_t batch, int J, const BatchMap& map) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(trisolve16_kernel<OP, JB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTs16Smem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int64_t grid = (batch + kTs16Warps * 8 - 1) / (kTs16Warps * 8);
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  trisolve16_kernel<OP, JB><<<(unsigned)grid, kTs16Warps * 32, kTs16Smem, s>>>(T, Y, X, batch, J, map);
  return cudaGetLastError();
}

template <int JB>
static cudaError_t launch_trisolve16(cudaStream_t s, int op, const double* T, const double* Y, double* X,
                                     int64_t batch, int J, const BatchMap& map) {
  if (op == 0) return launch_trisolve16_op<0, JB>(s, T, Y, X, batch, J, map);
  if (op == 1) return launch_trisolve16_op<1, JB>(s, T, Y, X, batch, J, map);
  return launch_trisolve16_op<2, JB>(s, T, Y, X, batch, J, map);
}

// ------------------------------------------------------------------------------------------------
// qr_lstsq (nd4js src/la/qr.js:186-273) for thin factors, fused: x = R[:L,:L]^-1 (Q[:, :L]^T y), L = min(M, I) <= 32.
// One warp per matrix.  Lane i accumulates (Q^T y)_i over the rows of Q in the reference's order (k ascending, product and
// sum rounded separately, starting from the zero the reference's result array is initialised with): a row of Q is one
// coalesced 256-byte read.  The upper triangle of R's leading block is staged in shared memory packed by columns (4.2 KB per
// matrix, a column is contiguous: conflict-free; 32 warps per SM) and the back substitution runs right-looking with k descending, as _triu_solve does per entry
// (tri.js:73-98).  Bit-identical with the reference; Q, R, y are each read once, nothing but x is written.
// ------------------------------------------------------------------------------------------------
constexpr int kLsWarps = 8;
constexpr int kLsTri = 32 * 33 / 2;   // R's leading block, upper triangle packed by columns: (i, k), i <= k, at k(k+1)/2 + i
constexpr size_t kLsSmem = sizeof(double) * kLsWarps * kLsTri;

template <int JB>
__global__ void __launch_bounds__(kLsWarps * 32, JB == 1 ? 4 : 3)
qr_lstsq32_kernel(const double* __restrict__ Q, const double* __restrict__ R, const double* __restrict__ Y,
                  double* __restrict__ X, int64_t batch, int N, int M, int I, int J) {
  extern __shared__ __align__(16) double ls_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * kLsWarps + warp;
  if (m >= batch) return;  // warp-uniform
  const int L = M < I ? M : I;
  double* rt = ls_smem + warp * kLsTri;
  const double* q = Q + m * (int64_t)N * M;
  const double* r = R + m * (int64_t)M * I;
  const double* y = Y + m * (int64_t)N * J;
  double* x = X + m * (int64_t)I * J;

  {  // rows of R's leading block, coalesced, asynchronously: they are not needed before the first back substitution
    const uint32_t rt_s = (uint32_t)__cvta_generic_to_shared(rt);
    for (int i = 0; i < L; i++)
      if (lane < L && lane >= i)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(rt_s + (uint32_t)(lane * (lane + 1) / 2 + i) * 8u), "l"(r + (int64_t)i * I + lane) : "memory");
  }
  double dii = 1.0, yii = 1.0;

#pragma unroll 1
  for (int j0 = 0; j0 < J; j0 += JB) {
    double s[JB];
#pragma unroll
    for (int jb = 0; jb < JB; jb++) s[jb] = 0.0;
    // Q^T y, k ascending (the unrolled body keeps 8 row reads of 256 B in flight per warp: the kernel is HBM bound)
#pragma unroll 8
    for (int k = 0; k < N; k++) {
      const double qk = (lane < L) ? ldg1_stream(q + (int64_t)k * M + lane) : 0.0;
#pragma unroll
      for (int jb = 0; jb < JB; jb++) {
        const double yk = (j0 + jb < J) ? __ldg(y + (int64_t)k * J + j0 + jb) : 0.0;
        s[jb] = add_rn(s[jb], mul_rn(qk, yk));
      }
    }
    if (j0 == 0) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
      // reciprocal of my diagonal entry (hoisted part of the IEEE division, common.cuh)
      dii = (lane < L) ? rt[lane * (lane + 1) / 2 + lane] : 1.0;
      yii = col_recip(dii).y;
    }
    for (int k = L - 1; k >= 0; k--) {  // back substitution, k descending
      ColRecip rc;
      rc.b = shfl(dii, k);
      rc.y = shfl(yii, k);
      rc.bhi = __int_as_float(__double2hiint(rc.b));
      rc.bnorm = recip_range_ok(rc.b);
      const double rik = rt[k * (k + 1) / 2 + (lane < k ? lane : k)];
      double num[JB], xk[JB];
      bool ok = true;
#pragma unroll
      for (int jb = 0; jb < JB; jb++) {
        num[jb] = shfl(s[jb], k);
        xk[jb] = div_col(num[jb], rc, ok);
      }
      if (!ok) {
#pragma unroll
        for (int jb = 0; jb < JB; jb++) xk[jb] = ts_ieee_div(num[jb], rc.b);
      }
#pragma unroll
      for (int jb = 0; jb < JB; jb++) {
        if (lane < k) s[jb] = sub_rn(s[jb], mul_rn(rik, xk[jb]));
        else if (lane == k) s[jb] = xk[jb];
      }
    }
#pragma unroll
    for (int jb = 0; jb < JB; jb++)
      if (j0 + jb < J && lane < I) x[(int64_t)lane * J + j0 + jb] = (lane < L) ? s[jb] : 0.0;
  }
}

cudaError_t launch_qr_lstsq(cudaStream_t s, const double* Q, const double* R, const double* Y, double* X,
                            int64_t batch, int N, int M, int I, int J) {
  if (batch <= 0) return cudaSuccess;
  if (M > 32 || I > 32) return cudaErrorInvalidValue;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(qr_lstsq32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLsSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qr_lstsq32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLsSmem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int64_t grid = (batch + kLsWarps - 1) / kLsWarps;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  if (J == 1) qr_lstsq32_kernel<1><<<(unsigned)grid, kLsWarps * 32, kLsSmem, s>>>(Q, R, Y, X, batch, N, M, I, J);
  else qr_lstsq32_kernel<4><<<(unsigned)grid, kLsWarps * 32, kLsSmem, s>>>(Q, R, Y, X, batch, N, M, I, J);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// 16 < M <= 64 (and M < 16 with few right-hand sides per matrix): one warp per matrix, the referenced triangle staged in
// shared memory (odd row stride: the column walk of the right-looking update is conflict-free, a row is contiguous),
// lane l owns rows l and l + 32 of the right-hand side.  As soon as x_k is known (owner's partial sum broadcast by a
// shuffle, IEEE quotient through the hoisted reciprocal of div.rn's fast path, the full division outside its ranges —
// a warp-uniform branch), every row subtracts t_ik * x_k: each entry sees the reference's own sequence (k ascending
// forward, descending backward; product and difference rounded separately), X is bit-identical.
// ------------------------------------------------------------------------------------------------
constexpr int kTsWarpWarps = 2;

template <int OP, int R, int JB>
__global__ void __launch_bounds__(kTsWarpWarps * 32)
trisolve_warp_kernel(const double* __restrict__ T, const double* __restrict__ Y, double* __restrict__ X,
                     int64_t batch, int M, int J, BatchMap map) {
  constexpr bool UPPER = (OP == 1);
  extern __shared__ __align__(16) double tsw_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * kTsWarpWarps + warp;
  if (m >= batch) return;  // warp-uniform
  const int ld = M | 1;
  double* tile = tsw_smem + (size_t)warp * (M * ld + M);
  double* rcp = tile + M * ld;
  int64_t to, yo;
  decode_batch2(map, m, to, yo);
  {
    const double* src = T + to;
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    const float inv = 1.0f / (float)M;
    for (int e = lane; e < M * M; e += 32) {
      const int i = (int)(((float)e + 0.5f) * inv), k = e - i * M;    // exact: e < 4096
      if (UPPER ? (k >= i) : (k <= i))
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tile_s + (uint32_t)(i * ld + k) * 8u), "l"(src + e) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    for (int i = lane; i < M; i += 32) rcp[i] = col_recip(tile[i * ld + i]).y;
    __syncwarp();
  }
  const double* y = Y + yo;
  double* x = X + m * (int64_t)M * J;
  for (int j0 = 0; j0 < J; j0 += JB) {   // JB right-hand sides share the chain of a step
    double sv[R][JB];
#pragma unroll
    for (int rr = 0; rr < R; rr++)
#pragma unroll
      for (int jb = 0; jb < JB; jb++) {
        const int i = lane + 32 * rr;
        sv[rr][jb] = (i < M && j0 + jb < J) ? y[(int64_t)i * J + j0 + jb] : 0.0;
      }
    auto quotient = [&](int k, double (&xk)[JB]) {
      ColRecip rc;
      rc.b = tile[k * ld + k];
      rc.y = rcp[k];
      rc.bhi = __int_as_float(__double2hiint(rc.b));
      rc.bnorm = recip_range_ok(rc.b);
      double num[JB];
      bool ok = true;
#pragma unroll
      for (int jb = 0; jb < JB; jb++) {
        const double mine = (R > 1 && k >= 32) ? sv[R - 1][jb] : sv[0][jb];
        num[jb] = shfl(mine, k & 31);
        xk[jb] = div_col(num[jb], rc, ok);
      }
      if (!ok) {   // warp-uniform
#pragma unroll
        for (int jb = 0; jb < JB; jb++) xk[jb] = ts_ieee_div(num[jb], rc.b);
      }
    };
    if (OP != 1) {
      for (int k = 0; k < M; k++) {
        double xk[JB];
        quotient(k, xk);
#pragma unroll
        for (int rr = 0; rr < R; rr++) {
          const int i = lane + 32 * rr;
          const double tik = (i > k && i < M) ? tile[i * ld + k] : 0.0;
#pragma unroll
          for (int jb = 0; jb < JB; jb++) {
            if (i > k && i < M) sv[rr][jb] = sub_rn(sv[rr][jb], mul_rn(tik, xk[jb]));
            else if (i == k) sv[rr][jb] = xk[jb];
          }
        }
      }
    }
    if (OP != 0) {
      for (int k = M - 1; k >= 0; k--) {
        double xk[JB];
        quotient(k, xk);
#pragma unroll
        for (int rr = 0; rr < R; rr++) {
          const int i = lane + 32 * rr;
          const double tik = (i < k) ? (UPPER ? tile[i * ld + k] : tile[k * ld + i]) : 0.0;
#pragma unroll
          for (int jb = 0; jb < JB; jb++) {
            if (i < k) sv[rr][jb] = sub_rn(sv[rr][jb], mul_rn(tik, xk[jb]));
            else if (i == k) sv[rr][jb] = xk[jb];
          }
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < R; rr++)
#pragma unroll
      for (int jb = 0; jb < JB; jb++) {
        const int i = lane + 32 * rr;
        if (i < M && j0 + jb < J) x[(int64_t)i * J + j0 + jb] = sv[rr][jb];
      }
  }
}

template <int OP, int R, int JB>
static cudaError_t launch_trisolve_warp_op(cudaStream_t s, const double* T, const double* Y, double* X,
                                           int64_t batch, int M, int J, const BatchMap& map) {
  const size_t smem = sizeof(double) * kTsWarpWarps * ((size_t)M * (M | 1) + M);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(trisolve_warp_kernel<OP, R, JB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(double) * kTsWarpWarps * (64 * 65 + 64)));
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int64_t grid = (batch + kTsWarpWarps - 1) / kTsWarpWarps;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  trisolve_warp_kernel<OP, R, JB><<<(unsigned)grid, kTsWarpWarps * 32, smem, s>>>(T, Y, X, batch, M, J, map);
  return cudaGetLastError();
}

template <int R, int JB>
static cudaError_t launch_trisolve_warp(cudaStream_t s, int op, const double* T, const double* Y, double* X,
                                        int64_t batch, int M, int J, const BatchMap& map) {
  if (op == 0) return launch_trisolve_warp_op<0, R, JB>(s, T, Y, X, batch, M, J, map);
  if (op == 1) return launch_trisolve_warp_op<1, R, JB>(s, T, Y, X, batch, M, J, map);
  return launch_trisolve_warp_op<2, R, JB>(s, T, Y, X, batch, M, J, map);
}

cudaError_t launch_tri_solve(cudaStream_t s, int op, const double* T, const double* Y, double* X,
                             int64_t batch, int M, int J, const BatchMap& map) {
  if (batch <= 0) return cudaSuccess;
  if (M == 16) {
    if (J == 1) return launch_trisolve16<1>(s, op, T, Y, X, batch, J, map);
    if (J == 2) return launch_trisolve16<2>(s, op, T, Y, X, batch, J, map);
    return launch_trisolve16<4>(s, op, T, Y, X, batch, J, map);
  }
  // one warp per matrix with the triangle in shared memory; tiny systems with many right-hand sides keep the
  // thread-per-column kernel (its X / Y accesses are coalesced over the columns)
  if (M > 16 && M <= 32) return J == 1 ? launch_trisolve_warp<1, 1>(s, op, T, Y, X, batch, M, J, map) : launch_trisolve_warp<1, 4>(s, op, T, Y, X, batch, M, J, map);
  if (M > 32 && M <= 64) return J == 1 ? launch_trisolve_warp<2, 1>(s, op, T, Y, X, batch, M, J, map) : launch_trisolve_warp<2, 4>(s, op, T, Y, X, batch, M, J, map);
  const int64_t threads = batch * J;
  const int64_t grid = (threads + kSolveThreads - 1) / kSolveThreads;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  tri_solve_kernel<<<(unsigned)grid, kSolveThreads, 0, s>>>(op, T, Y, X, batch, M, J, map);
  return cudaGetLastError();
}

}  // namespace nd4b
