// qr.cu — batched thin/full QR by Householder reflections (replaces the Givens QR of nd4js
// src/la/qr.js:27-145; same shapes: Q[rows,L], R[L,cols], L=min(rows,cols)).
//
// Sign convention: every reflector is chosen so that R_kk = +||x|| >= 0 (LAPACK dlarfgp style,
// cancellation-free), i.e. the unique QR factor with non-negative diagonal.  The reference's tall
// branch leaves arbitrary signs on diag(R) (qr.js:111-115), so parity is checked after normalising
// the reference with sign(R_ii) (tests/), plus reconstruction and orthogonality residuals.
//
//  * qr64x32_kernel : rows=64, cols=32; one warp per matrix, the whole matrix in registers
//                     (lane (r,c) = (lane>>2, lane&3) owns rows r+8i and columns 8j+2c+{0,1}: 64 doubles),
//                     column norms and v^T A by warp shuffles, reflectors kept in place below the diagonal,
//                     Q formed in place by backward accumulation (dorg2r).  40 960 B and 218 453 flop per matrix.
//  * qr_generic_kernel : any shape; one CTA per matrix working in a global (L2-resident) scratch copy.
#include "common.cuh"
#include "kernels.h"

namespace nd4b {

// Householder scalars for x = (x0, rest) with sigma = |rest|^2:  H = I - tau * v v^T, v = (1, rest/v0),
// H x = (beta, 0), beta = ||x|| >= 0.
struct Reflector { double beta, tau, inv_v0; };

__device__ __forceinline__ Reflector make_reflector(double x0, double sigma) {
  Reflector h;
  if (sigma == 0.0) {  // nothing below the diagonal: identity, or a sign flip to keep R_kk >= 0
    h.beta = fabs(x0);
    h.tau = (x0 < 0.0) ? 2.0 : 0.0;
    h.inv_v0 = 0.0;  // rest is all zero anyway
    return h;
  }
  // scale-free norm: ||x|| = max * sqrt((x0/max)^2 + sigma/max^2) is not needed for |x| in a sane range;
  // guard only against overflow/underflow of x0^2 + sigma
  double nrm = sqrt(fma(x0, x0, sigma));
  const double v0 = (x0 <= 0.0) ? (x0 - nrm) : (-sigma / (x0 + nrm));
  const double v0sq = v0 * v0;
  h.beta = nrm;
  h.tau = 2.0 * v0sq / (sigma + v0sq);
  h.inv_v0 = 1.0 / v0;
  return h;
}

constexpr int kQrWarps = 4;

__global__ void __launch_bounds__(kQrWarps * 32)
qr64x32_kernel(const double* __restrict__ A, double* __restrict__ Q, double* __restrict__ R, int64_t batch) {
  constexpr int ROWS = 64, COLS = 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, c = lane & 3;
  const int64_t m = (int64_t)blockIdx.x * kQrWarps + warp;
  if (m >= batch) return;  // warp-uniform
  const double* a_in = A + m * (ROWS * COLS);

  // a[i][jj]: row r+8i, column 8*(jj>>1) + 2c + (jj&1)
  double a[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int jp = 0; jp < 4; jp++) {
      const double2 v = ldg2_stream(a_in + (r + 8 * i) * COLS + 8 * jp + 2 * c);
      a[i][2 * jp] = v.x;
      a[i][2 * jp + 1] = v.y;
    }

  double mytau = 0.0;  // lane k keeps tau_k

  // ---------------- R phase ----------------
#pragma unroll
  for (int k = 0; k < COLS; k++) {
    const int i0 = k >> 3, r0 = k & 7;          // row k lives in slot i0 of lanes with r == r0
    const int ck = (k & 7) >> 1;                // column k lives in lanes with c == ck ...
    const int jk = 2 * (k >> 3) + (k & 1);      // ... in register column jk
    const int src_c = (lane & ~3) | ck;         // same row group, owning column group

    double ss = 0.0;
#pragma unroll
    for (int i = i0; i < 8; i++) {
      const double x = a[i][jk];
      if (i > i0 || r > r0) ss = fma(x, x, ss);
    }
    ss += shfl_xor(ss, 4);
    ss += shfl_xor(ss, 8);
    ss += shfl_xor(ss, 16);
    const double sigma = shfl(ss, ck);
    const double x0 = shfl(a[i0][jk], 4 * r0 + ck);
    const Reflector h = make_reflector(x0, sigma);
    if (lane == k) mytau = h.tau;

    double vv[8];
#pragma unroll
    for (int i = i0; i < 8; i++) {
      const double x = shfl(a[i][jk], src_c);
      vv[i] = x * h.inv_v0;
    }
    vv[i0] = (r > r0) ? vv[i0] : (r == r0 ? 1.0 : 0.0);
    if (c == ck) {
#pragma unroll
      for (int i = i0; i < 8; i++)
        if (i > i0 || r > r0) a[i][jk] = vv[i];
      if (r == r0) a[i0][jk] = h.beta;
    }

    // trailing columns: slots jj >= 2*(k>>3) can hold columns > k
#pragma unroll
    for (int jj = 2 * (k >> 3); jj < 8; jj++) {
      const int col = 8 * (jj >> 1) + 2 * c + (jj & 1);
      double w = 0.0;
#pragma unroll
      for (int i = i0; i < 8; i++) w = fma(vv[i], a[i][jj], w);
      w += shfl_xor(w, 4);
      w += shfl_xor(w, 8);
      w += shfl_xor(w, 16);
      const double f = (col > k) ? h.tau * w : 0.0;
#pragma unroll
      for (int i = i0; i < 8; i++) a[i][jj] = fma(-f, vv[i], a[i][jj]);
    }
  }

  // ---------------- store R (rows 0..31 are slots i < 4) ----------------
  {
    double* r_out = R + m * (COLS * COLS);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int row = r + 8 * i;
#pragma unroll
      for (int jp = 0; jp < 4; jp++) {
        const int col = 8 * jp + 2 * c;
        const double x = (col >= row) ? a[i][2 * jp] : 0.0;
        const double y = (col + 1 >= row) ? a[i][2 * jp + 1] : 0.0;
        stg2_stream(r_out + row * COLS + col, x, y);
      }
    }
  }

  // ---------------- Q phase: Q = H_0 ... H_31 [I;0], formed in place over the reflectors ----------------
#pragma unroll
  for (int k = COLS - 1; k >= 0; k--) {
    const int i0 = k >> 3, r0 = k & 7;
    const int ck = (k & 7) >> 1;
    const int jk = 2 * (k >> 3) + (k & 1);
    const int src_c = (lane & ~3) | ck;
    const double tau = shfl(mytau, k);

    double vv[8];
#pragma unroll
    for (int i = i0; i < 8; i++) vv[i] = shfl(a[i][jk], src_c);
    vv[i0] = (r > r0) ? vv[i0] : (r == r0 ? 1.0 : 0.0);

#pragma unroll
    for (int jj = 2 * (k >> 3); jj < 8; jj++) {
      const int col = 8 * (jj >> 1) + 2 * c + (jj & 1);
      double w = 0.0;
#pragma unroll
      for (int i = i0; i < 8; i++) w = fma(vv[i], a[i][jj], w);
      w += shfl_xor(w, 4);
      w += shfl_xor(w, 8);
      w += shfl_xor(w, 16);
      const double f = (col > k) ? tau * w : 0.0;
#pragma unroll
      for (int i = i0; i < 8; i++) a[i][jj] = fma(-f, vv[i], a[i][jj]);
    }
    // column k of Q: e_k - tau * v_k ; rows < k are zero
    if (c == ck) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int row = r + 8 * i;
        double q = 0.0;
        if (i >= i0) q = (row > k) ? -tau * vv[i] : (row == k ? 1.0 - tau : 0.0);
        a[i][jk] = q;
      }
    }
  }

  double* q_out = Q + m * (ROWS * COLS);
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int jp = 0; jp < 4; jp++)
      stg2_stream(q_out + (r + 8 * i) * COLS + 8 * jp + 2 * c, a[i][2 * jp], a[i][2 * jp + 1]);
}

// ------------------------------------------------------------------------------------------------
// Generic shape: one CTA per matrix.  W (rows x cols scratch) holds A, then R above / reflectors
// below the diagonal; tau in scratch; Q is accumulated directly in the output.
// ------------------------------------------------------------------------------------------------
constexpr int kQrGenThreads = 128;

__device__ __forceinline__ double block_sum(double v, double* red) {
  // red: shared array of kQrGenThreads/32 + 1 doubles
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  __syncthreads();  // protect red from the previous use
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kQrGenThreads / 32; w++) s += red[w];
  return s;
}

__global__ void __launch_bounds__(kQrGenThreads)
qr_generic_kernel(const double* __restrict__ A, double* __restrict__ Q, double* __restrict__ R,
                  int64_t batch, int rows, int cols, double* __restrict__ work) {
  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const int L = rows < cols ? rows : cols;
  const int64_t rc = (int64_t)rows * cols;
  double* W = work + m * (rc + L);
  double* tau = W + rc;
  const double* a_in = A + m * rc;
  double* q = Q + m * (int64_t)rows * L;
  double* rr = R + m * (int64_t)L * cols;
  __shared__ double red[kQrGenThreads / 32 + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kQrGenThreads / 32;

  for (int64_t e = tid; e < rc; e += kQrGenThreads) W[e] = a_in[e];
  __syncthreads();

  for (int k = 0; k < L; k++) {
    double part = 0.0;
    for (int i = k + 1 + tid; i < rows; i += kQrGenThreads) {
      const double x = W[(int64_t)i * cols + k];
      part = fma(x, x, part);
    }
    const double sigma = block_sum(part, red);
    const double x0 = W[(int64_t)k * cols + k];
    const Reflector h = make_reflector(x0, sigma);
    __syncthreads();  // everyone has read x0
    for (int i = k + 1 + tid; i < rows; i += kQrGenThreads) W[(int64_t)i * cols + k] *= h.inv_v0;
    if (tid == 0) { W[(int64_t)k * cols + k] = h.beta; tau[k] = h.tau; }
    __syncthreads();
    // trailing columns: one warp per column
    for (int j = k + 1 + warp; j < cols; j += NW) {
      double w = (lane == 0) ? W[(int64_t)k * cols + j] : 0.0;  // v_k = 1
      for (int i = k + 1 + lane; i < rows; i += 32) w = fma(W[(int64_t)i * cols + k], W[(int64_t)i * cols + j], w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
      const double f = h.tau * w;
      if (lane == 0) W[(int64_t)k * cols + j] -= f;
      for (int i = k + 1 + lane; i < rows; i += 32)
        W[(int64_t)i * cols + j] = fma(-f, W[(int64_t)i * cols + k], W[(int64_t)i * cols + j]);
    }
    __syncthreads();
  }

  // R
  for (int64_t e = tid; e < (int64_t)L * cols; e += kQrGenThreads) {
    const int i = (int)(e / cols), j = (int)(e % cols);
    rr[e] = (j >= i) ? W[(int64_t)i * cols + j] : 0.0;
  }
  // Q = H_0 ... H_{L-1} [I_L; 0]  (rows x L), backward accumulation
  for (int64_t e = tid; e < (int64_t)rows * L; e += kQrGenThreads) {
    const int i = (int)(e / L), j = (int)(e % L);
    q[e] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int k = L - 1; k >= 0; k--) {
    const double tk = tau[k];
    for (int j = k + warp; j < L; j += NW) {
      double w = (lane == 0) ? q[(int64_t)k * L + j] : 0.0;
      for (int i = k + 1 + lane; i < rows; i += 32) w = fma(W[(int64_t)i * cols + k], q[(int64_t)i * L + j], w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
      const double f = tk * w;
      if (lane == 0) q[(int64_t)k * L + j] -= f;
      for (int i = k + 1 + lane; i < rows; i += 32)
        q[(int64_t)i * L + j] = fma(-f, W[(int64_t)i * cols + k], q[(int64_t)i * L + j]);
    }
    __syncthreads();
  }
}

size_t qr_workspace_bytes(int64_t batch, int rows, int cols) {
  if (rows == 64 && cols == 32) return 0;
  const int L = rows < cols ? rows : cols;
  return sizeof(double) * (size_t)batch * ((size_t)rows * cols + L);
}

cudaError_t launch_qr(cudaStream_t s, const double* A, double* Q, double* R, int64_t batch, int rows, int cols,
                      double* work, size_t work_bytes) {
  if (batch <= 0) return cudaSuccess;
  const bool aligned = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Q) |
                         reinterpret_cast<uintptr_t>(R)) & 15) == 0;
  if (rows == 64 && cols == 32 && aligned) {
    const int64_t grid = (batch + kQrWarps - 1) / kQrWarps;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    qr64x32_kernel<<<(unsigned)grid, kQrWarps * 32, 0, s>>>(A, Q, R, batch);
    return cudaGetLastError();
  }
  const int L = rows < cols ? rows : cols;
  const size_t need = sizeof(double) * (size_t)batch * ((size_t)rows * cols + L);
  if (work == nullptr || work_bytes < need || batch > 0x7fffffffLL) return cudaErrorInvalidValue;
  qr_generic_kernel<<<(unsigned)batch, kQrGenThreads, 0, s>>>(A, Q, R, batch, rows, cols, work);
  return cudaGetLastError();
}

}  // namespace nd4b
