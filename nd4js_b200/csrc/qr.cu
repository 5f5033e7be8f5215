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
#include <cstdlib>

namespace nd4b {

// Householder scalars for x = (x0, rest) with sigma = |rest|^2:  H = I - tau * v v^T, v = (1, rest/v0),
// H x = (beta, 0), beta = ||x|| >= 0 (cancellation-free, dlarfgp style).  Used by the generic kernel.
struct Reflector { double beta, tau, inv_v0; };

__device__ __forceinline__ Reflector make_reflector(double x0, double sigma) {
  Reflector h;
  // Nothing below the diagonal — exactly, or to working precision (|rest| <= 2^-53 |x0|: then ||x|| == |x0| in double and
  // the reflector is the identity to 2^-106), or a column that is numerically zero against the prescaled matrix (largest
  // entry in [1, 2)).  Without the last two cases an exactly rank-deficient matrix drives this into the subnormal range: the
  // residue left under the diagonal shrinks by 2^-53 per eliminated column, after about ten columns v0 = -sigma / (x0 + ||x||)
  // is ~1e-160, v0^2 is subnormal with a few bits left, and tau = 2 v0^2 / (sigma + v0^2) makes H non-orthogonal at the 1e-7
  // level (found by the reference's own suite: qr_test.js "random matrices with zero rows", 48 x 52 of rank 4).
  if (sigma == 0.0 || sigma <= x0 * x0 * 0x1p-106 || fma(x0, x0, sigma) < 0x1p-900) {  // identity, or a sign flip to keep R_kk >= 0
    h.beta = fabs(x0);
    h.tau = (x0 < 0.0) ? 2.0 : 0.0;
    h.inv_v0 = 0.0;  // rest is all zero anyway
    return h;
  }
  double nrm = sqrt(fma(x0, x0, sigma));
  const double v0 = (x0 <= 0.0) ? (x0 - nrm) : (-sigma / (x0 + nrm));
  const double v0sq = v0 * v0;
  h.beta = nrm;
  h.tau = 2.0 * v0sq / (sigma + v0sq);
  h.inv_v0 = 1.0 / v0;
  return h;
}

// The register kernel uses the classic (dlarfg) sign choice beta = -sign(x0)*||x||, which needs one rsqrt and
// one reciprocal on the critical path instead of a sqrt and three divisions:
//   v0 = x0 - beta = x0 + sign(x0)||x||,   tau = (beta - x0)/beta = 1 + |x0|/||x||.
// Rows of R / columns of Q with beta < 0 are negated when they are stored, so that diag(R) >= 0 as documented.
struct ReflectorS { double beta, tau, inv_v0; };

__device__ __forceinline__ ReflectorS make_reflector_signed(double x0, double sigma) {
  ReflectorS h;
  const double s = fma(x0, x0, sigma);
  if (s < 0x1p-900) { h.beta = 0.0; h.tau = 0.0; h.inv_v0 = 0.0; return h; }   // numerically zero against the prescaled matrix
  const double rn = rsqrt(s);
  const double nrm = s * rn;
  h.beta = -copysign(nrm, x0);
  h.tau = fma(fabs(x0), rn, 1.0);
  h.inv_v0 = 1.0 / (x0 - h.beta);
  return h;
}

constexpr int kQrWarps = 4;

// One Householder step on the register tile.  KB = k>>3 and E = k&1 are compile-time (they select registers),
// ck = (k&7)>>1 is a run-time value: 8 code variants per phase instead of 32 keeps the kernel inside the
// instruction cache.  a[i][jj]: row r+8i, column 8*(jj>>1) + 2c + (jj&1).
template <int KB, int E, bool QPHASE>
__device__ __forceinline__ void qr_step(double (&a)[8][8], int ck, int lane, int r, int c, double& mytau, double& mysgn) {
  constexpr int i0 = KB, jk = 2 * KB + E;
  const int k = 8 * KB + 2 * ck + E;
  const int r0 = 2 * ck + E;               // row k lives in slot i0 of lanes with r == r0
  const int src_c = (lane & ~3) | ck;      // same row group, owning column group
  double tau, inv_v0 = 1.0;

  if (!QPHASE) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int i = i0; i < 8; i++) {
      const double x = a[i][jk];
      const double xx = (i > i0 || r > r0) ? x : 0.0;
      if ((i - i0) & 1) s1 = fma(xx, xx, s1); else s0 = fma(xx, xx, s0);
    }
    double ss = s0 + s1;
    ss += shfl_xor(ss, 4);
    ss += shfl_xor(ss, 8);
    ss += shfl_xor(ss, 16);
    const double sigma = shfl(ss, ck);
    const double x0 = shfl(a[i0][jk], 4 * r0 + ck);
    const ReflectorS h = make_reflector_signed(x0, sigma);
    tau = h.tau;
    inv_v0 = h.inv_v0;
    if (lane == k) { mytau = h.tau; mysgn = (h.beta < 0.0) ? -1.0 : 1.0; }
    if (c == ck && r == r0) a[i0][jk] = h.beta;
  } else {
    tau = shfl(mytau, k);
  }

  double vv[8];
#pragma unroll
  for (int i = i0; i < 8; i++) {
    const double x = shfl(a[i][jk], src_c);
    vv[i] = QPHASE ? x : x * inv_v0;
  }
  vv[i0] = (r > r0) ? vv[i0] : (r == r0 ? 1.0 : 0.0);
  if (!QPHASE && c == ck) {
#pragma unroll
    for (int i = i0; i < 8; i++)
      if (i > i0 || r > r0) a[i][jk] = vv[i];
  }

  // w_j = v^T A_j for the column slots that can hold columns > k, all slots interleaved (independent chains)
  constexpr int J0 = 2 * KB, NJ = 8 - J0;
  double w[NJ];
  {
    // two partial chains per column slot (even / odd row slots) halve the dependent-FMA depth
    double w1[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) { w[j] = vv[i0] * a[i0][J0 + j]; w1[j] = (i0 + 1 < 8) ? vv[(i0 + 1) & 7] * a[(i0 + 1) & 7][J0 + j] : 0.0; }
#pragma unroll
    for (int i = i0 + 2; i < 8; i++)
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        if ((i - i0) & 1) w1[j] = fma(vv[i], a[i][J0 + j], w1[j]);
        else w[j] = fma(vv[i], a[i][J0 + j], w[j]);
      }
#pragma unroll
    for (int j = 0; j < NJ; j++) w[j] += w1[j];
  }
#pragma unroll
  for (int j = 0; j < NJ; j++) w[j] += shfl_xor(w[j], 4);
#pragma unroll
  for (int j = 0; j < NJ; j++) w[j] += shfl_xor(w[j], 8);
#pragma unroll
  for (int j = 0; j < NJ; j++) w[j] += shfl_xor(w[j], 16);
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const int jj = J0 + j;
    const int col = 8 * (jj >> 1) + 2 * c + (jj & 1);
    const double f = (col > k) ? -tau * w[j] : 0.0;
#pragma unroll
    for (int i = i0; i < 8; i++) a[i][jj] = fma(f, vv[i], a[i][jj]);
  }

  if (QPHASE && c == ck) {  // column k of Q: e_k - tau * v_k ; rows < k are zero
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int row = r + 8 * i;
      double q = 0.0;
      if (i >= i0) q = (row > k) ? -tau * vv[i] : (row == k ? 1.0 - tau : 0.0);
      a[i][jk] = q;
    }
  }
}

template <int KB, bool QPHASE>
__device__ __forceinline__ void qr_block(double (&a)[8][8], int lane, int r, int c, double& mytau, double& mysgn) {
  if (!QPHASE) {
#pragma unroll 1
    for (int ck = 0; ck < 4; ck++) {
      qr_step<KB, 0, false>(a, ck, lane, r, c, mytau, mysgn);
      qr_step<KB, 1, false>(a, ck, lane, r, c, mytau, mysgn);
    }
  } else {
#pragma unroll 1
    for (int ck = 3; ck >= 0; ck--) {
      qr_step<KB, 1, true>(a, ck, lane, r, c, mytau, mysgn);
      qr_step<KB, 0, true>(a, ck, lane, r, c, mytau, mysgn);
    }
  }
}

template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
qr64x32_kernel(const double* __restrict__ A, double* __restrict__ Q, double* __restrict__ R, int64_t batch) {
  constexpr int ROWS = 64, COLS = 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, c = lane & 3;
  const int64_t m = (int64_t)blockIdx.x * WARPS + warp;
  if (m >= batch) return;  // warp-uniform
  const double* a_in = A + m * (ROWS * COLS);

  double a[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int jp = 0; jp < 4; jp++) {
      const double2 v = ldg2_stream(a_in + (r + 8 * i) * COLS + 8 * jp + 2 * c);
      a[i][2 * jp] = v.x;
      a[i][2 * jp + 1] = v.y;
    }

  // scale guard (see pow2_prescale): one max-reduction per matrix, a multiplication only for extreme magnitudes
  double amax = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) amax = fmax(amax, fabs(a[i][j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, shfl_xor(amax, o));
  const double pre = pow2_prescale(amax);
  if (pre != 1.0) {  // warp-uniform
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) a[i][j] *= pre;
  }
  const double post = 1.0 / pre;  // exact (power of two)

  double mytau = 0.0, mysgn = 1.0;  // lane k keeps tau_k and the sign of beta_k

  qr_block<0, false>(a, lane, r, c, mytau, mysgn);
  qr_block<1, false>(a, lane, r, c, mytau, mysgn);
  qr_block<2, false>(a, lane, r, c, mytau, mysgn);
  qr_block<3, false>(a, lane, r, c, mytau, mysgn);

  // ---------------- store R (rows 0..31 are slots i < 4), rows with beta < 0 negated ----------------
  {
    double* r_out = R + m * (COLS * COLS);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int row = r + 8 * i;
      const double sg = shfl(mysgn, row) * post;
#pragma unroll
      for (int jp = 0; jp < 4; jp++) {
        const int col = 8 * jp + 2 * c;
        const double x = (col >= row) ? sg * a[i][2 * jp] : 0.0;
        const double y = (col + 1 >= row) ? sg * a[i][2 * jp + 1] : 0.0;
        stg2_stream(r_out + row * COLS + col, x, y);
      }
    }
  }

  // ---------------- Q = H_0 ... H_31 [I;0], formed in place over the reflectors ----------------
  qr_block<3, true>(a, lane, r, c, mytau, mysgn);
  qr_block<2, true>(a, lane, r, c, mytau, mysgn);
  qr_block<1, true>(a, lane, r, c, mytau, mysgn);
  qr_block<0, true>(a, lane, r, c, mytau, mysgn);

  double* q_out = Q + m * (ROWS * COLS);
  double sg[8];
#pragma unroll
  for (int jj = 0; jj < 8; jj++) sg[jj] = shfl(mysgn, 8 * (jj >> 1) + 2 * c + (jj & 1));
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int jp = 0; jp < 4; jp++)
      stg2_stream(q_out + (r + 8 * i) * COLS + 8 * jp + 2 * c, sg[2 * jp] * a[i][2 * jp], sg[2 * jp + 1] * a[i][2 * jp + 1]);
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

  double amax = 0.0;
  for (int64_t e = tid; e < rc; e += kQrGenThreads) { const double x = a_in[e]; W[e] = x; amax = fmax(amax, fabs(x)); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(kFull, amax, o));
  if (lane == 0) red[warp] = amax;
  __syncthreads();
  amax = 0.0;
#pragma unroll
  for (int w = 0; w < NW; w++) amax = fmax(amax, red[w]);
  const double pre = pow2_prescale(amax), post = 1.0 / pre;
  __syncthreads();
  if (pre != 1.0) {
    for (int64_t e = tid; e < rc; e += kQrGenThreads) W[e] *= pre;
    __syncthreads();
  }

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
    rr[e] = (j >= i) ? W[(int64_t)i * cols + j] * post : 0.0;
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

// ------------------------------------------------------------------------------------------------
// Any shape that fits in shared memory: one CTA per matrix, W = [A | Y] (rows x (cols + ycols), odd leading dimension so
// that a column walk is bank-conflict free) and Q (rows x L) in shared memory — the same Householder steps as the global
// kernels above at shared-memory instead of L2 latency.  FORM_Q: qr_decomp (Q rows x L, R L x cols); otherwise
// _qr_decomp_inplace (R rows x cols, Q^T Y rows x ycols).  Shapes that do not fit keep the global-memory kernels.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int qr_smem_ld(int n) { return n | 1; }
inline size_t qr_smem_bytes(int rows, int cols, int ycols, bool form_q) {
  const int L = rows < cols ? rows : cols;
  return sizeof(double) * ((size_t)rows * qr_smem_ld(cols + ycols) + (form_q ? (size_t)rows * qr_smem_ld(L) : 0) + L);
}
constexpr size_t kQrSmemLimit = 200 * 1024;

template <bool FORM_Q>
__global__ void __launch_bounds__(kQrGenThreads)
qr_smem_kernel(const double* __restrict__ A, const double* __restrict__ Y, double* __restrict__ Q, double* __restrict__ R,
               double* __restrict__ QtY, int64_t batch, int rows, int cols, int ycols) {
  extern __shared__ __align__(16) double qr_sm[];
  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const int L = rows < cols ? rows : cols;
  const int wc = cols + ycols, ldw = qr_smem_ld(wc), ldq = qr_smem_ld(L);
  double* W = qr_sm;
  double* q = W + (size_t)rows * ldw;
  double* tau = q + (FORM_Q ? (size_t)rows * ldq : 0);
  __shared__ double red[kQrGenThreads / 32 + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kQrGenThreads / 32;
  const double* a_in = A + m * (int64_t)rows * cols;

  double amax = 0.0;
  for (int e = tid; e < rows * cols; e += kQrGenThreads) {
    const double x = a_in[e];
    W[(e / cols) * ldw + e % cols] = x;
    amax = fmax(amax, fabs(x));
  }
  if (ycols > 0) {
    const double* y_in = Y + m * (int64_t)rows * ycols;
    for (int e = tid; e < rows * ycols; e += kQrGenThreads) W[(e / ycols) * ldw + cols + e % ycols] = y_in[e];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(kFull, amax, o));
  if (lane == 0) red[warp] = amax;
  __syncthreads();
  amax = 0.0;
#pragma unroll
  for (int w = 0; w < NW; w++) amax = fmax(amax, red[w]);
  const double pre = pow2_prescale(amax), post = 1.0 / pre;
  __syncthreads();
  if (pre != 1.0) {  // A only: Q^T Y does not depend on the scale of A
    for (int e = tid; e < rows * cols; e += kQrGenThreads) W[(e / cols) * ldw + e % cols] *= pre;
    __syncthreads();
  }

  for (int k = 0; k < L; k++) {
    double part = 0.0;
    for (int i = k + 1 + tid; i < rows; i += kQrGenThreads) {
      const double x = W[i * ldw + k];
      part = fma(x, x, part);
    }
    const double sigma = block_sum(part, red);
    const double x0 = W[k * ldw + k];
    const Reflector h = make_reflector(x0, sigma);
    __syncthreads();  // everyone has read x0
    for (int i = k + 1 + tid; i < rows; i += kQrGenThreads) W[i * ldw + k] *= h.inv_v0;
    if (tid == 0) { W[k * ldw + k] = h.beta; tau[k] = h.tau; }
    __syncthreads();
    for (int j = k + 1 + warp; j < wc; j += NW) {  // trailing columns of A and all columns of Y: one warp per column
      double w = (lane == 0) ? W[k * ldw + j] : 0.0;  // v_k = 1
      for (int i = k + 1 + lane; i < rows; i += 32) w = fma(W[i * ldw + k], W[i * ldw + j], w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
      const double f = h.tau * w;
      if (lane == 0) W[k * ldw + j] -= f;
      for (int i = k + 1 + lane; i < rows; i += 32) W[i * ldw + j] = fma(-f, W[i * ldw + k], W[i * ldw + j]);
    }
    __syncthreads();
  }

  if (!FORM_Q) {
    double* r_out = R + m * (int64_t)rows * cols;
    for (int e = tid; e < rows * cols; e += kQrGenThreads) {
      const int i = e / cols, j = e % cols;
      r_out[e] = (j >= i) ? W[i * ldw + j] * post : 0.0;
    }
    double* y_out = QtY + m * (int64_t)rows * ycols;
    for (int e = tid; e < rows * ycols; e += kQrGenThreads) y_out[e] = W[(e / ycols) * ldw + cols + e % ycols];
    return;
  }
  double* r_out = R + m * (int64_t)L * cols;
  for (int e = tid; e < L * cols; e += kQrGenThreads) {
    const int i = e / cols, j = e % cols;
    r_out[e] = (j >= i) ? W[i * ldw + j] * post : 0.0;
  }
  // Q = H_0 ... H_{L-1} [I_L; 0], backward accumulation in shared memory
  for (int e = tid; e < rows * L; e += kQrGenThreads) q[(e / L) * ldq + e % L] = (e / L == e % L) ? 1.0 : 0.0;
  __syncthreads();
  for (int k = L - 1; k >= 0; k--) {
    const double tk = tau[k];
    for (int j = k + warp; j < L; j += NW) {
      double w = (lane == 0) ? q[k * ldq + j] : 0.0;
      for (int i = k + 1 + lane; i < rows; i += 32) w = fma(W[i * ldw + k], q[i * ldq + j], w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
      const double f = tk * w;
      if (lane == 0) q[k * ldq + j] -= f;
      for (int i = k + 1 + lane; i < rows; i += 32) q[i * ldq + j] = fma(-f, W[i * ldw + k], q[i * ldq + j]);
    }
    __syncthreads();
  }
  double* q_out = Q + m * (int64_t)rows * L;
  for (int e = tid; e < rows * L; e += kQrGenThreads) q_out[e] = q[(e / L) * ldq + e % L];
}

template <bool FORM_Q>
static cudaError_t launch_qr_smem(cudaStream_t s, const double* A, const double* Y, double* Q, double* R, double* QtY,
                                  int64_t batch, int rows, int cols, int ycols) {
  const size_t bytes = qr_smem_bytes(rows, cols, ycols, FORM_Q);
  static size_t attr_bytes[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && attr_bytes[dev] < bytes) {
    cudaError_t e = cudaFuncSetAttribute(qr_smem_kernel<FORM_Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQrSmemLimit);
    if (e != cudaSuccess) return e;
    attr_bytes[dev] = kQrSmemLimit;
  }
  qr_smem_kernel<FORM_Q><<<(unsigned)batch, kQrGenThreads, bytes, s>>>(A, Y, Q, R, QtY, batch, rows, cols, ycols);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// _qr_decomp_inplace (nd4js src/la/qr.js:147-183): A[M,N] -> R (same shape, zero below the diagonal), Y[M,L] -> Q^T Y,
// Q never formed.  One CTA per matrix working directly in the two output buffers (L2-resident); the reflectors are
// applied to the trailing columns of A and to all columns of Y alike (one warp per column).  Same sign convention as the
// other QR kernels: diag(R) >= 0 (the reference's Givens sequence leaves arbitrary signs; rows of R and of Q^T Y flip together).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kQrGenThreads)
qr_inplace_kernel(const double* __restrict__ A, const double* __restrict__ Y, double* __restrict__ R, double* __restrict__ QtY,
                  int64_t batch, int M, int N, int L) {
  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const int64_t mn = (int64_t)M * N, ml = (int64_t)M * L;
  const double* a_in = A + m * mn;
  const double* y_in = Y + m * ml;
  double* W = R + m * mn;
  double* Z = QtY + m * ml;
  __shared__ double red[kQrGenThreads / 32 + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kQrGenThreads / 32;

  double amax = 0.0;
  for (int64_t e = tid; e < mn; e += kQrGenThreads) { const double x = a_in[e]; W[e] = x; amax = fmax(amax, fabs(x)); }
  for (int64_t e = tid; e < ml; e += kQrGenThreads) Z[e] = y_in[e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(kFull, amax, o));
  if (lane == 0) red[warp] = amax;
  __syncthreads();
  amax = 0.0;
#pragma unroll
  for (int w = 0; w < NW; w++) amax = fmax(amax, red[w]);
  const double pre = pow2_prescale(amax), post = 1.0 / pre;
  __syncthreads();
  if (pre != 1.0) {
    for (int64_t e = tid; e < mn; e += kQrGenThreads) W[e] *= pre;
    __syncthreads();
  }

  const int K = M < N ? M : N;
  for (int k = 0; k < K; k++) {
    double part = 0.0;
    for (int i = k + 1 + tid; i < M; i += kQrGenThreads) {
      const double x = W[(int64_t)i * N + k];
      part = fma(x, x, part);
    }
    const double sigma = block_sum(part, red);
    const double x0 = W[(int64_t)k * N + k];
    const Reflector h = make_reflector(x0, sigma);
    __syncthreads();  // everyone has read x0
    for (int i = k + 1 + tid; i < M; i += kQrGenThreads) W[(int64_t)i * N + k] *= h.inv_v0;
    if (tid == 0) W[(int64_t)k * N + k] = h.beta;
    __syncthreads();
    // columns k+1..N-1 of A, then the L columns of Y: one warp per column
    const int ncol = (N - 1 - k) + L;
    for (int jj = warp; jj < ncol; jj += NW) {
      const bool in_a = jj < N - 1 - k;
      double* col = in_a ? W + (k + 1 + jj) : Z + (jj - (N - 1 - k));
      const int ld = in_a ? N : L;
      double w = (lane == 0) ? col[(int64_t)k * ld] : 0.0;  // v_k = 1
      for (int i = k + 1 + lane; i < M; i += 32) w = fma(W[(int64_t)i * N + k], col[(int64_t)i * ld], w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
      const double f = h.tau * w;
      if (lane == 0) col[(int64_t)k * ld] -= f;
      for (int i = k + 1 + lane; i < M; i += 32) col[(int64_t)i * ld] = fma(-f, W[(int64_t)i * N + k], col[(int64_t)i * ld]);
    }
    __syncthreads();
  }
  for (int64_t e = tid; e < mn; e += kQrGenThreads) {
    const int i = (int)(e / N), j = (int)(e % N);
    W[e] = (j >= i) ? W[e] * post : 0.0;
  }
}

cudaError_t launch_qr_inplace(cudaStream_t s, const double* A, const double* Y, double* R, double* QtY,
                              int64_t batch, int M, int N, int L) {
  if (batch <= 0) return cudaSuccess;
  if (batch > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  // M <= 64, N <= 32, at most 8 right-hand sides: the register kernel's R phases plus its block reflectors applied to Y
  // (ND4B_QR_PADDED=0 keeps the shared-memory kernel for A/B timing)
  static int padded = -1;
  if (padded < 0) {
    const char* ev = getenv("ND4B_QR_PADDED");
    padded = ev ? atoi(ev) : 1;
  }
  if (padded && M <= 64 && N <= 32 && L >= 1 && L <= 8)
    return launch_qr_inplace_blocked(s, A, Y, R, QtY, batch, M, N, L);
  if (qr_smem_bytes(M, N, L, false) <= kQrSmemLimit) return launch_qr_smem<false>(s, A, Y, nullptr, R, QtY, batch, M, N, L);
  qr_inplace_kernel<<<(unsigned)batch, kQrGenThreads, 0, s>>>(A, Y, R, QtY, batch, M, N, L);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Tiny matrices (rows, cols <= 8): one lane per matrix, 32 matrices per warp.  The warp's matrices (one contiguous block)
// are staged with 8-byte cp.async into odd-stride shared-memory slots (conflict-free), every lane runs an unblocked
// Householder QR on its own slot (reflectors stored below the diagonal, Q accumulated backwards from [I; 0], rows of R /
// columns of Q with a negative diagonal entry negated, power-of-two prescaling against overflow of the squared norms)
// and the results leave through the slots as coalesced stores.  A 3x3 factorisation moves 216 B: HBM bound.
// ------------------------------------------------------------------------------------------------
constexpr int kQrTinyWarps = 4;

__global__ void __launch_bounds__(kQrTinyWarps * 32)
qr_tiny_kernel(const double* __restrict__ A, double* __restrict__ Q, double* __restrict__ R, int64_t batch, int rows, int cols) {
  extern __shared__ __align__(16) double qrt_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = rows < cols ? rows : cols;
  const int na = rows * cols, nq = rows * L, nr = L * cols;
  const int sa = na | 1, sq = nq | 1;
  double* as = qrt_smem + (size_t)warp * 32 * (sa + sq + 8);
  double* qs = as + 32 * sa;
  double* taus = qs + 32 * sq;
  const int64_t m0 = ((int64_t)blockIdx.x * kQrTinyWarps + warp) * 32;
  if (m0 >= batch) return;  // warp-uniform
  const int nmat = (int)min((int64_t)32, batch - m0);
  {
    const double* src = A + m0 * na;
    const uint32_t as_s = (uint32_t)__cvta_generic_to_shared(as);
    const float inv = 1.0f / (float)na;
    for (int g = lane; g < nmat * na; g += 32) {
      const int q = (int)(((float)g + 0.5f) * inv), e = g - q * na;   // exact: g < 2048
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(as_s + (uint32_t)(q * sa + e) * 8u), "l"(src + g) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
  }
  if (lane < nmat) {
    double* a = as + lane * sa;
    double* q = qs + lane * sq;
    double* tau = taus + lane * 8;
    double amax = 0.0;
    for (int e = 0; e < na; e++) amax = fmax(amax, fabs(a[e]));
    const double pre = pow2_prescale(amax);
    if (pre != 1.0)
      for (int e = 0; e < na; e++) a[e] *= pre;
    for (int j = 0; j < L; j++) {
      double sigma = 0.0;
      for (int i = j + 1; i < rows; i++) sigma = fma(a[i * cols + j], a[i * cols + j], sigma);
      const double alpha = a[j * cols + j];
      double t = 0.0;
      if (sigma > 0x1p-900) {   // below: a numerically zero column against the prescaled matrix (subnormal squares lose their bits)
        const double nrm = sqrt(fma(alpha, alpha, sigma));
        const double beta = alpha >= 0.0 ? -nrm : nrm;
        t = (beta - alpha) / beta;
        const double sc = 1.0 / (alpha - beta);
        for (int i = j + 1; i < rows; i++) a[i * cols + j] *= sc;
        a[j * cols + j] = beta;
        for (int c = j + 1; c < cols; c++) {
          double w = a[j * cols + c];
          for (int i = j + 1; i < rows; i++) w = fma(a[i * cols + j], a[i * cols + c], w);
          w *= t;
          a[j * cols + c] -= w;
          for (int i = j + 1; i < rows; i++) a[i * cols + c] = fma(-w, a[i * cols + j], a[i * cols + c]);
        }
      }
      tau[j] = t;
    }
    // Q = H_0 .. H_{L-1} [I; 0], accumulated backwards
    for (int e = 0; e < nq; e++) q[e] = 0.0;
    for (int j = 0; j < L; j++) q[j * L + j] = 1.0;
    for (int j = L - 1; j >= 0; j--) {
      const double t = tau[j];
      if (t != 0.0)
        for (int c = j; c < L; c++) {
          double w = q[j * L + c];
          for (int i = j + 1; i < rows; i++) w = fma(a[i * cols + j], q[i * L + c], w);
          w *= t;
          q[j * L + c] -= w;
          for (int i = j + 1; i < rows; i++) q[i * L + c] = fma(-w, a[i * cols + j], q[i * L + c]);
        }
    }
    // R: exact zeros below the diagonal, diag >= 0 (row j of R and column j of Q negated together), scale undone
    const double post = 1.0 / pre;
    for (int j = 0; j < L; j++) {
      const bool neg = a[j * cols + j] < 0.0;
      for (int c = 0; c < cols; c++) {
        double v = (c < j) ? 0.0 : a[j * cols + c];
        if (neg && c >= j) v = -v;
        a[j * cols + c] = (post != 1.0) ? v * post : v;
      }
      if (neg)
        for (int i = 0; i < rows; i++) q[i * L + j] = -q[i * L + j];
    }
  }
  __syncwarp();
  {
    double* qd = Q + m0 * nq;
    const float invq = 1.0f / (float)nq;
    for (int g = lane; g < nmat * nq; g += 32) {
      const int q = (int)(((float)g + 0.5f) * invq), e = g - q * nq;
      qd[g] = qs[q * sq + e];
    }
    double* rd = R + m0 * nr;
    const float invr = 1.0f / (float)nr;
    for (int g = lane; g < nmat * nr; g += 32) {
      const int q = (int)(((float)g + 0.5f) * invr), e = g - q * nr;   // R = the first L rows of the slot (row stride cols)
      rd[g] = as[q * sa + e];
    }
  }
}

static cudaError_t launch_qr_tiny(cudaStream_t s, const double* A, double* Q, double* R, int64_t batch, int rows, int cols) {
  const int L = rows < cols ? rows : cols;
  const size_t smem = sizeof(double) * kQrTinyWarps * 32 * (size_t)((rows * cols | 1) + (rows * L | 1) + 8);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(qr_tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int64_t grid = (batch + kQrTinyWarps * 32 - 1) / (kQrTinyWarps * 32);
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  qr_tiny_kernel<<<(unsigned)grid, kQrTinyWarps * 32, smem, s>>>(A, Q, R, batch, rows, cols);
  return cudaGetLastError();
}

size_t qr_workspace_bytes(int64_t batch, int rows, int cols) {
  if (rows <= 64 && cols <= 32) return 0;
  if (qr_smem_bytes(rows, cols, 0, true) <= kQrSmemLimit) return 0;  // shared-memory kernel
  const int L = rows < cols ? rows : cols;
  return sizeof(double) * (size_t)batch * ((size_t)rows * cols + L);
}

cudaError_t launch_qr(cudaStream_t s, const double* A, double* Q, double* R, int64_t batch, int rows, int cols,
                      double* work, size_t work_bytes) {
  if (batch <= 0) return cudaSuccess;
  const bool aligned = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Q) |
                         reinterpret_cast<uintptr_t>(R)) & 15) == 0;
  if (rows == 64 && cols == 32) {
    if (batch > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    static int variant = -1;
    if (variant < 0) {
      const char* ev = getenv("ND4B_QR_VARIANT");
      variant = ev ? atoi(ev) : 0;  // 1 = the unblocked register kernel, 2 = the blocked one at two CTAs per SM (A/B profiling)
    }
    if (variant != 1 || !aligned) return launch_qr64x32_blocked(s, A, Q, R, batch, variant);
    // 222 registers -> 2 CTAs of 4 warps per SM.  Capping registers for 3 CTAs (168) or using 1-/2-warp CTAs was
    // measured slower (spills): 2.05 / 2.27 / 2.20 ms vs 1.86 ms on C4.
    qr64x32_kernel<kQrWarps, 2><<<(unsigned)((batch + kQrWarps - 1) / kQrWarps), kQrWarps * 32, 0, s>>>(A, Q, R, batch);
    return cudaGetLastError();
  }
  // rows <= 64, cols <= 32: zero-padded into the tuned 64 x 32 register kernel (19 ns per matrix whatever the shape);
  // tiny matrices stay with the shared-memory kernel (ND4B_QR_PADDED=0 switches the padded route off for A/B timing)
  static int padded = -1;
  if (padded < 0) {
    const char* ev = getenv("ND4B_QR_PADDED");
    padded = ev ? atoi(ev) : 1;
  }
  if (padded && rows <= 8 && cols <= 8 && batch >= 64) return launch_qr_tiny(s, A, Q, R, batch, rows, cols);
  if (padded && rows <= 64 && cols <= 32 && batch <= 0x7fffffffLL)
    return launch_qr_padded_blocked(s, A, Q, R, batch, rows, cols);
  if (batch <= 0x7fffffffLL && qr_smem_bytes(rows, cols, 0, true) <= kQrSmemLimit)
    return launch_qr_smem<true>(s, A, nullptr, Q, R, nullptr, batch, rows, cols, 0);
  const int L = rows < cols ? rows : cols;
  const size_t need = sizeof(double) * (size_t)batch * ((size_t)rows * cols + L);
  if (work == nullptr || work_bytes < need || batch > 0x7fffffffLL) return cudaErrorInvalidValue;
  qr_generic_kernel<<<(unsigned)batch, kQrGenThreads, 0, s>>>(A, Q, R, batch, rows, cols, work);
  return cudaGetLastError();
}

}  // namespace nd4b
