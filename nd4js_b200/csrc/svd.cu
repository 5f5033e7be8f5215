// svd.cu — batched one-sided (Hestenes) Jacobi SVD: the new nd.la.svd_jac_1sided.
//
// Contract = the svd_jac_* family of nd4js (src/la/svd_jac_2sided.js:30-144, post-processing rules
// src/la/_svd_jac_utils.js:123-188, shapes src/help.js:2321-2337):  A[rows,cols] -> U[rows,L], sv[L], V[L,cols],
// L = min(rows,cols), A = U diag(sv) V, sv >= +0 sorted descending (stable), U^T U = I, V V^T = I also for
// rank-deficient input (zero columns are completed to an orthonormal basis; diagonal input is reproduced
// exactly, as the reference's shared suite demands of every svd_jac*: _generic_test_svd_decomp.js:180-216).
//
// Algorithm: G = A (tall) or A^T (wide); rotate column pairs (p,q) of G until all pairs satisfy
// (g_p.g_q)^2 <= (m*eps)^2 |g_p|^2 |g_q|^2 (the analogue of the reference's stopping rule, svd_jac_2sided.js:57,112);
// pairs are visited in a round-robin (tournament) order: n/2 disjoint pairs per step, n-1 steps per sweep;
// the same rotations accumulate V.  Then sv_j = |g_j|, U = G diag(1/sv).
//
//  * svd64_kernel      : 64x64; one CTA (256 threads) per matrix; G and V column-major in shared memory;
//                        8 lanes per column pair, 32 pairs (one full tournament step) in flight per CTA.
//  * svd_generic_kernel: any shape; one CTA per matrix, G^T and V^T in global (L2-resident) scratch.
#include "common.cuh"
#include "kernels.h"
#include <float.h>

namespace nd4b {

constexpr int kMaxSweeps = 30;
constexpr double kEps = 2.220446049250313e-16;

struct Rot { double c, s; };

// Rotation that orthogonalises columns with |g_p|^2=a, |g_q|^2=b, g_p.g_q=d (d != 0):
//   g_p' = c g_p - s g_q,  g_q' = s g_p + c g_q.
__device__ __forceinline__ Rot make_rotation(double a, double b, double d) {
  const double zeta = (b - a) / (2.0 * d);
  const double az = fabs(zeta);
  double t;
  if (az > 1e150) t = 0.5 / zeta;  // sqrt(1+zeta^2) would overflow
  else t = copysign(1.0, zeta) / (az + sqrt(fma(zeta, zeta, 1.0)));
  Rot r;
  r.c = 1.0 / sqrt(fma(t, t, 1.0));
  r.s = r.c * t;
  return r;
}

// Tournament pairing of n2 (even) players, step s in [0,n2-1), slot P in [0,n2/2).
__device__ __forceinline__ void rr_pair(int n2, int s, int P, int& p, int& q) {
  const int m1 = n2 - 1;
  if (P == 0) { p = m1; q = s; }
  else { p = (s + P) % m1; q = (s - P + m1) % m1; }
  if (p > q) { const int x = p; p = q; q = x; }
}

// Completes the columns flagged in `zero` (bit per sorted position is not needed: flags are per column)
// to an orthonormal set.  One warp.  Columns are contiguous vectors of length m at base + col*ld.
// zero_flag[col] != 0 marks a column to be replaced.
__device__ void complete_basis_warp(double* base, int ld, int m, int n, const int* zero_flag, int* done_flag) {
  const int lane = threadIdx.x & 31;
  for (int z = 0; z < n; z++) {
    if (!zero_flag[z]) continue;
    // row with the smallest squared norm over the columns that already hold unit vectors
    double best = DBL_MAX;
    int best_i = 0x7fffffff;
    for (int i = lane; i < m; i += 32) {
      double rn = 0.0;
      for (int a = 0; a < n; a++)
        if (!zero_flag[a] || done_flag[a]) { const double x = base[(int64_t)a * ld + i]; rn = fma(x, x, rn); }
      if (rn < best) { best = rn; best_i = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(kFull, best, o);
      const int oi = __shfl_xor_sync(kFull, best_i, o);
      if (ob < best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    double* u = base + (int64_t)z * ld;
    for (int i = lane; i < m; i += 32) u[i] = (i == best_i) ? 1.0 : 0.0;
    __syncwarp();
    for (int pass = 0; pass < 2; pass++)
      for (int a = 0; a < n; a++) {
        if (a == z || (zero_flag[a] && !done_flag[a])) continue;
        const double* ua = base + (int64_t)a * ld;
        double dot = 0.0;
        for (int i = lane; i < m; i += 32) dot = fma(ua[i], u[i], dot);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(kFull, dot, o);
        if (dot != 0.0)
          for (int i = lane; i < m; i += 32) u[i] = fma(-dot, ua[i], u[i]);
        __syncwarp();
      }
    double nn = 0.0;
    for (int i = lane; i < m; i += 32) nn = fma(u[i], u[i], nn);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(kFull, nn, o);
    const double inv = 1.0 / sqrt(nn);
    if (nn != 1.0)
      for (int i = lane; i < m; i += 32) u[i] *= inv;
    __syncwarp();
    if (lane == 0) done_flag[z] = 1;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// 64x64
// ------------------------------------------------------------------------------------------------
constexpr int kSvd64Threads = 256;
constexpr int kSvd64LD = 66;  // column stride in doubles: 528 B = 33*16 B keeps 16-byte alignment
constexpr size_t kSvd64Smem = sizeof(double) * (2 * 64 * kSvd64LD + 64) + sizeof(int) * (64 * 3 + 4);

__global__ void __launch_bounds__(kSvd64Threads, 3)
svd64_kernel(const double* __restrict__ A, double* __restrict__ U, double* __restrict__ SV, double* __restrict__ V,
             int64_t batch, int* sweeps_out, int* fail_out) {
  constexpr int N = 64, LD = kSvd64LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Gs = reinterpret_cast<double*>(smem_raw);
  double* Vs = Gs + N * LD;
  double* sq = Vs + N * LD;                 // squared column norms, then sigma
  int* perm = reinterpret_cast<int*>(sq + N);  // perm[l] = column holding the l-th largest sigma
  int* zero_flag = perm + N;
  int* done_flag = zero_flag + N;

  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7, P = warp * 4 + (lane >> 3);
  const double* a_in = A + m * (N * N);

  for (int e = tid; e < N * N; e += kSvd64Threads) {
    const int i = e >> 6, j = e & 63;
    Gs[j * LD + i] = ldg1_stream(a_in + e);
    Vs[j * LD + i] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();

  const double tol2 = (N * kEps) * (N * kEps);
  int sweeps = 0;
  bool converged = false;
  for (; sweeps < kMaxSweeps && !converged;) {
    int rotated = 0;
    sweeps++;
    for (int s = 0; s < N - 1; s++) {
      int p, q;
      rr_pair(N, s, P, p, q);
      double2* gp = reinterpret_cast<double2*>(Gs + p * LD + 2 * sub);
      double2* gq = reinterpret_cast<double2*>(Gs + q * LD + 2 * sub);
      double2 xp[4], xq[4];
      double a = 0.0, b = 0.0, d = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        xp[k] = gp[8 * k];  // rows 16k + 2sub + {0,1}
        xq[k] = gq[8 * k];
        a = fma(xp[k].x, xp[k].x, a); a = fma(xp[k].y, xp[k].y, a);
        b = fma(xq[k].x, xq[k].x, b); b = fma(xq[k].y, xq[k].y, b);
        d = fma(xp[k].x, xq[k].x, d); d = fma(xp[k].y, xq[k].y, d);
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        a += shfl_xor(a, o);
        b += shfl_xor(b, o);
        d += shfl_xor(d, o);
      }
      if (d * d > tol2 * a * b) {  // uniform inside the 8-lane group
        rotated = 1;
        const Rot r = make_rotation(a, b, d);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          double2 np, nq;
          np.x = fma(r.c, xp[k].x, -r.s * xq[k].x); np.y = fma(r.c, xp[k].y, -r.s * xq[k].y);
          nq.x = fma(r.s, xp[k].x, r.c * xq[k].x);  nq.y = fma(r.s, xp[k].y, r.c * xq[k].y);
          gp[8 * k] = np;
          gq[8 * k] = nq;
        }
        double2* vp = reinterpret_cast<double2*>(Vs + p * LD + 2 * sub);
        double2* vq = reinterpret_cast<double2*>(Vs + q * LD + 2 * sub);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const double2 yp = vp[8 * k], yq = vq[8 * k];
          double2 np, nq;
          np.x = fma(r.c, yp.x, -r.s * yq.x); np.y = fma(r.c, yp.y, -r.s * yq.y);
          nq.x = fma(r.s, yp.x, r.c * yq.x);  nq.y = fma(r.s, yp.y, r.c * yq.y);
          vp[8 * k] = np;
          vq[8 * k] = nq;
        }
      }
      __syncthreads();
    }
    converged = !__syncthreads_or(rotated);
  }
  if (tid == 0) {
    if (sweeps_out) atomicMax(sweeps_out, sweeps);
    if (!converged && fail_out) atomicExch(fail_out, 1);
  }

  // singular values: column norms (each 8-lane group handles columns P and P+32)
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int col = P + 32 * h;
    const double2* g = reinterpret_cast<const double2*>(Gs + col * LD + 2 * sub);
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const double2 x = g[8 * k]; a = fma(x.x, x.x, a); a = fma(x.y, x.y, a); }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) a += shfl_xor(a, o);
    if (sub == 0) sq[col] = sqrt(a);
  }
  __syncthreads();
  // stable descending rank + zero detection
  if (tid < N) {
    const double sj = sq[tid];
    int rank = 0;
    double smax = 0.0;
    for (int k = 0; k < N; k++) {
      const double sk = sq[k];
      rank += (sk > sj) || (sk == sj && k < tid);
      smax = fmax(smax, sk);
    }
    perm[rank] = tid;
    const int z = !(sj > smax * 1e-290) || !(sj >= DBL_MIN);
    zero_flag[tid] = z;
    done_flag[tid] = 0;
  }
  __syncthreads();
  // normalise the columns of G in place -> U columns
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int col = P + 32 * h;
    if (!zero_flag[col]) {
      const double sj = sq[col];
      double2* g = reinterpret_cast<double2*>(Gs + col * LD + 2 * sub);
#pragma unroll
      for (int k = 0; k < 4; k++) { double2 x = g[8 * k]; x.x /= sj; x.y /= sj; g[8 * k] = x; }
    }
  }
  __syncthreads();
  int any_zero = 0;
  for (int k = 0; k < N; k++) any_zero |= zero_flag[k];
  if (any_zero) {
    if (warp == 0) complete_basis_warp(Gs, LD, N, N, zero_flag, done_flag);
    __syncthreads();
  }

  double* u_out = U + m * (N * N);
  double* v_out = V + m * (N * N);
  for (int e = tid; e < N * N; e += kSvd64Threads) {
    const int i = e >> 6, l = e & 63;
    u_out[e] = Gs[perm[l] * LD + i];   // U[i][l]
    v_out[e] = Vs[perm[i] * LD + l];   // V[l'][j] with l' = i, j = l
  }
  if (tid < N) SV[m * N + tid] = zero_flag[perm[tid]] ? 0.0 : sq[perm[tid]];
}

// ------------------------------------------------------------------------------------------------
// Generic shape.  mm = max(rows,cols) (vector length), n = min(rows,cols) (number of columns of G).
// scratch per matrix: Gt[n*mm], Vt[n*n], sig[n], then ints perm[n], zero[n], done[n] (packed in doubles)
// ------------------------------------------------------------------------------------------------
constexpr int kSvdGenThreads = 256;

__host__ __device__ inline size_t svd_gen_scratch_doubles(int rows, int cols) {
  const size_t mm = rows > cols ? rows : cols, n = rows < cols ? rows : cols;
  return n * mm + n * n + n + (3 * n + 1) / 2 + 1;
}

__global__ void __launch_bounds__(kSvdGenThreads)
svd_generic_kernel(const double* __restrict__ A, double* __restrict__ U, double* __restrict__ SV, double* __restrict__ V,
                   int64_t batch, int rows, int cols, int* sweeps_out, int* fail_out, double* __restrict__ work) {
  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const bool wide = rows < cols;
  const int mm = wide ? cols : rows, n = wide ? rows : cols;
  double* Gt = work + m * svd_gen_scratch_doubles(rows, cols);
  double* Vt = Gt + (size_t)n * mm;
  double* sig = Vt + (size_t)n * n;
  int* perm = reinterpret_cast<int*>(sig + n);
  int* zero_flag = perm + n;
  int* done_flag = zero_flag + n;
  const double* a_in = A + m * (int64_t)rows * cols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kSvdGenThreads / 32;

  // G = A (tall: Gt[j][i] = A[i][j]) or A^T (wide: Gt[j][i] = A[j][i])
  for (int64_t e = tid; e < (int64_t)rows * cols; e += kSvdGenThreads) {
    if (wide) Gt[e] = a_in[e];
    else { const int i = (int)(e / cols), j = (int)(e % cols); Gt[(size_t)j * mm + i] = a_in[e]; }
  }
  for (int64_t e = tid; e < (int64_t)n * n; e += kSvdGenThreads) Vt[e] = (e / n == e % n) ? 1.0 : 0.0;
  __syncthreads();

  const double tol2 = ((double)mm * kEps) * ((double)mm * kEps);
  const int n2 = n + (n & 1);
  int sweeps = 0;
  bool converged = (n < 2);
  for (; sweeps < kMaxSweeps && !converged;) {
    int rotated = 0;
    sweeps++;
    for (int s = 0; s < n2 - 1; s++) {
      for (int P = warp; P < n2 / 2; P += NW) {
        int p, q;
        rr_pair(n2, s, P, p, q);
        if (q >= n) continue;  // bye (odd n)
        double* gp = Gt + (size_t)p * mm;
        double* gq = Gt + (size_t)q * mm;
        double a = 0.0, b = 0.0, d = 0.0;
        for (int i = lane; i < mm; i += 32) {
          const double x = gp[i], y = gq[i];
          a = fma(x, x, a); b = fma(y, y, b); d = fma(x, y, d);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += shfl_xor(a, o); b += shfl_xor(b, o); d += shfl_xor(d, o); }
        if (d * d > tol2 * a * b) {
          rotated = 1;
          const Rot r = make_rotation(a, b, d);
          for (int i = lane; i < mm; i += 32) {
            const double x = gp[i], y = gq[i];
            gp[i] = fma(r.c, x, -r.s * y);
            gq[i] = fma(r.s, x, r.c * y);
          }
          double* vp = Vt + (size_t)p * n;
          double* vq = Vt + (size_t)q * n;
          for (int i = lane; i < n; i += 32) {
            const double x = vp[i], y = vq[i];
            vp[i] = fma(r.c, x, -r.s * y);
            vq[i] = fma(r.s, x, r.c * y);
          }
        }
      }
      __syncthreads();
    }
    converged = !__syncthreads_or(rotated);
  }
  if (tid == 0) {
    if (sweeps_out) atomicMax(sweeps_out, sweeps);
    if (!converged && fail_out) atomicExch(fail_out, 1);
  }

  for (int j = warp; j < n; j += NW) {
    const double* g = Gt + (size_t)j * mm;
    double a = 0.0;
    for (int i = lane; i < mm; i += 32) a = fma(g[i], g[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += shfl_xor(a, o);
    if (lane == 0) sig[j] = sqrt(a);
  }
  __syncthreads();
  for (int j = tid; j < n; j += kSvdGenThreads) {
    const double sj = sig[j];
    int rank = 0;
    double smax = 0.0;
    for (int k = 0; k < n; k++) {
      const double sk = sig[k];
      rank += (sk > sj) || (sk == sj && k < j);
      smax = fmax(smax, sk);
    }
    perm[rank] = j;
    zero_flag[j] = !(sj > smax * 1e-290) || !(sj >= DBL_MIN);
    done_flag[j] = 0;
  }
  __syncthreads();
  for (int j = warp; j < n; j += NW) {
    if (zero_flag[j]) continue;
    double* g = Gt + (size_t)j * mm;
    const double sj = sig[j];
    for (int i = lane; i < mm; i += 32) g[i] /= sj;
  }
  __syncthreads();
  int any_zero = 0;
  for (int k = 0; k < n; k++) any_zero |= zero_flag[k];
  if (any_zero) {
    if (warp == 0) complete_basis_warp(Gt, mm, mm, n, zero_flag, done_flag);
    __syncthreads();
  }

  // outputs: L = n.  tall: U[i][l] = Gt[perm l][i], V[l][j] = Vt[perm l][j]
  //                  wide: U[i][l] = Vt[perm l][i], V[l][j] = Gt[perm l][j]
  double* u_out = U + m * (int64_t)rows * n;
  double* v_out = V + m * (int64_t)n * cols;
  const double* usrc = wide ? Vt : Gt;
  const int uld = wide ? n : mm;
  const double* vsrc = wide ? Gt : Vt;
  const int vld = wide ? mm : n;
  for (int64_t e = tid; e < (int64_t)rows * n; e += kSvdGenThreads) {
    const int i = (int)(e / n), l = (int)(e % n);
    u_out[e] = usrc[(size_t)perm[l] * uld + i];
  }
  for (int64_t e = tid; e < (int64_t)n * cols; e += kSvdGenThreads) {
    const int l = (int)(e / cols), j = (int)(e % cols);
    v_out[e] = vsrc[(size_t)perm[l] * vld + j];
  }
  for (int l = tid; l < n; l += kSvdGenThreads) SV[m * n + l] = zero_flag[perm[l]] ? 0.0 : sig[perm[l]];
}

size_t svd_workspace_bytes(int64_t batch, int rows, int cols) {
  if (rows == 64 && cols == 64) return 0;
  return sizeof(double) * (size_t)batch * svd_gen_scratch_doubles(rows, cols);
}

cudaError_t launch_svd_jac1(cudaStream_t s, const double* A, double* U, double* sv, double* V,
                            int64_t batch, int rows, int cols, int* sweeps, int* fail,
                            double* work, size_t work_bytes) {
  if (batch <= 0) return cudaSuccess;
  if (batch > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  if (rows == 64 && cols == 64) {
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(svd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSvd64Smem);
      if (e != cudaSuccess) return e;
      attr_set[dev] = true;
    }
    svd64_kernel<<<(unsigned)batch, kSvd64Threads, kSvd64Smem, s>>>(A, U, sv, V, batch, sweeps, fail);
    return cudaGetLastError();
  }
  const size_t need = svd_workspace_bytes(batch, rows, cols);
  if (work == nullptr || work_bytes < need) return cudaErrorInvalidValue;
  svd_generic_kernel<<<(unsigned)batch, kSvdGenThreads, 0, s>>>(A, U, sv, V, batch, rows, cols, sweeps, fail, work);
  return cudaGetLastError();
}

}  // namespace nd4b
