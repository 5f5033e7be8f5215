// qr_blocked.cu — batched 64x32 Householder QR with block reflectors on the FP64 tensor pipe (DMMA.8x8x4).
// Replaces the Givens QR of nd4js src/la/qr.js:80-145 for the C4 shape; same results contract as qr.cu
// (Q[64,32], R[32,32] exactly upper triangular, diag(R) >= 0).
//
// One warp per matrix, the matrix in registers in the accumulator layout of the TRANSPOSE:
//   lane (g,t) = (lane>>2, lane&3) owns columns 8i+g (i<4) and rows 8j+2t+e (j<8, e<2):  a[i][j][e].
// With that layout every product of the blocked algorithm is a chain of DMMA.8x8x4 whose fragments are the registers
// themselves (an accumulator tile X[g][2t+e] is the A fragment pair of a following product, k <-> 2t+e):
//   panel p (columns 8p..8p+7, one column per quad):  unblocked Householder, 8 steps; a step broadcasts the pivot
//       column from its quad (shuffles), forms all dot products with it (the pivot quad's is the squared norm) and
//       applies the rank-1 update — reflector scalars need one rsqrt and one reciprocal.
//   T of the compact WY form (H_1..H_8 = I - V T V^T):  G = V^T V by DMMA (both fragments are the same register),
//       T^-1 = striu(G) + diag(1/tau)  =>  T = (I+N)^-1 diag(tau), N = diag(tau) striu(G) nilpotent, and
//       (I+N)^-1 = (I-N)(I+N^2)(I+N^4) exactly: six 8x8 products, in both orientations where needed, so that no transpose is needed.
//   trailing columns:  W^T = C^T V,  Z^T = W^T T,  C^T -= Z^T V^T  (the last product needs V with rows on the g index:
//       the clean panel is written once to shared memory, row-major, and read back with 16-byte loads).
//   Q = H_1 .. H_32 [I;0] by blocks from the last panel: trailing columns as above with T^T, the panel's own columns
//       as E_p - V_p (T_p L_p^T) (L_p = unit lower 8x8 head of V_p) — no serial reflector loop in the Q phase.
// Work per matrix: 40 960 B of HBM traffic, 218 453 flop by the Householder convention.
#include "common.cuh"
#include "kernels.h"

namespace nd4b {

namespace {

constexpr int kQbWarps = 4;
constexpr int kVS = 34;                                   // row stride (doubles) of the per-warp V store
constexpr int kQbWarpDoubles = 64 * kVS + 4 * 64;         // V store + the four T's (accumulator layout, 2 per lane)
constexpr size_t kQbSmem = sizeof(double) * kQbWarpDoubles * kQbWarps;

struct Acc { double x, y; };  // 8x8 tile in accumulator layout: x = M[g][2t], y = M[g][2t+1]

// X * Y given X and Y^T in accumulator layout (A fragments = X, B fragments B[k<->2t+e][n=g] = Y^T[g][2t+e]).
__device__ __forceinline__ Acc mm8(const Acc& X, const Acc& YT) {
  Acc o{0.0, 0.0};
  dmma884(o.x, o.y, X.x, YT.x);
  dmma884(o.x, o.y, X.y, YT.y);
  return o;
}

// ---- panel P: unblocked Householder on columns 8P..8P+7 (my column: 8P+g), rows >= 8P ----
template <int P>
__device__ __forceinline__ void qb_panel(double (&a)[4][8][2], int lane, int g, int t, double& tau_q, double& sgn_q) {
  double inv_q = 0.0;
  tau_q = 0.0;
  sgn_q = 1.0;
#pragma unroll 1
  for (int kk = 0; kk < 8; kk++) {
    const int src = 4 * kk + t;
    double x[8][2];
#pragma unroll
    for (int j = P; j < 8; j++) {
      x[j][0] = shfl(a[P][j][0], src);
      x[j][1] = shfl(a[P][j][1], src);
    }
    // rows above the pivot row k = 8P+kk hold finished R entries
    x[P][0] = (2 * t >= kk) ? x[P][0] : 0.0;
    x[P][1] = (2 * t + 1 >= kk) ? x[P][1] : 0.0;
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int j = P; j < 8; j++) {
      d0 = fma(x[j][0], a[P][j][0], d0);
      d1 = fma(x[j][1], a[P][j][1], d1);
    }
    double d = d0 + d1;
    d += shfl_xor(d, 1);
    d += shfl_xor(d, 2);
    const double s = shfl(d, 4 * kk);  // |x|^2 (rows >= k) from the pivot quad
    const int e0 = kk & 1, t0 = kk >> 1;
    const int qsrc = (lane & ~3) | t0;
    const double x0 = shfl(e0 ? x[P][1] : x[P][0], qsrc);         // pivot element
    const double akj = shfl(e0 ? a[P][P][1] : a[P][P][0], qsrc);  // my column's element in the pivot row
    double beta = 0.0, tau = 0.0, v0 = 0.0, inv_v0 = 0.0;
    if (s != 0.0) {  // warp-uniform; s == 0: H = I
      const double rn = rsqrt(s);
      const double nrm = s * rn;
      beta = -copysign(nrm, x0);
      tau = fma(fabs(x0), rn, 1.0);
      v0 = x0 - beta;
      inv_v0 = 1.0 / v0;
    }
    // v = x~/v0 with x~ = (v0; x below), w = v^T a = (d - beta*akj)/v0, a -= tau*w*v
    const double w = (d - beta * akj) * inv_v0;
    const double f = (g > kk) ? -(tau * w) * inv_v0 : 0.0;
    if (t == t0) {
      if (e0) x[P][1] = v0; else x[P][0] = v0;
    }
#pragma unroll
    for (int j = P; j < 8; j++) {
      a[P][j][0] = fma(f, x[j][0], a[P][j][0]);
      a[P][j][1] = fma(f, x[j][1], a[P][j][1]);
    }
    if (g == kk) {
      tau_q = tau;
      inv_q = inv_v0;
      sgn_q = (beta < 0.0) ? -1.0 : 1.0;
      if (t == t0) {
        if (e0) a[P][P][1] = beta; else a[P][P][0] = beta;
      }
    }
  }
  // v = x / v0 below the diagonal (deferred: later steps never read a finished column)
  a[P][P][0] = (2 * t > g) ? a[P][P][0] * inv_q : a[P][P][0];
  a[P][P][1] = (2 * t + 1 > g) ? a[P][P][1] * inv_q : a[P][P][1];
#pragma unroll
  for (int j = P + 1; j < 8; j++) {
    a[P][j][0] *= inv_q;
    a[P][j][1] *= inv_q;
  }
}

// T of panel P in both orientations (accumulator layout) from the clean V held in a[P][j>=P].
template <int P>
__device__ __forceinline__ void qb_make_t(const double (&a)[4][8][2], int g, int t, double tau_q, Acc& T, Acc& TT) {
  Acc G{0.0, 0.0};
#pragma unroll
  for (int j = P; j < 8; j++) {
    dmma884(G.x, G.y, a[P][j][0], a[P][j][0]);
    dmma884(G.x, G.y, a[P][j][1], a[P][j][1]);
  }
  const double tq0 = shfl(tau_q, 4 * (2 * t)), tq1 = shfl(tau_q, 4 * (2 * t + 1));  // tau of columns 2t, 2t+1
  const int c0 = 2 * t, c1 = 2 * t + 1;
  Acc N, NT, I;
  N.x = (g < c0) ? tau_q * G.x : 0.0;
  N.y = (g < c1) ? tau_q * G.y : 0.0;
  NT.x = (c0 < g) ? tq0 * G.x : 0.0;  // N^T[g][c] = N[c][g] = tau_c G[c][g], G symmetric
  NT.y = (c1 < g) ? tq1 * G.y : 0.0;
  I.x = (g == c0) ? 1.0 : 0.0;
  I.y = (g == c1) ? 1.0 : 0.0;
  const Acc N2 = mm8(N, NT), N2T = mm8(NT, N);
  const Acc N4T = mm8(N2T, N2);
  const Acc ImN{I.x - N.x, I.y - N.y};
  const Acc IpN2T{I.x + N2T.x, I.y + N2T.y}, IpN4T{I.x + N4T.x, I.y + N4T.y};
  const Acc A1 = mm8(ImN, IpN2T);                             // (I-N)(I+N^2)
  const Acc Inv = mm8(A1, IpN4T), InvT = mm8(IpN4T, A1);      // (I+N)^-1 and its transpose
  T.x = Inv.x * tq0;     // T[a][b] = Inv[a][b] * tau_b
  T.y = Inv.y * tq1;
  TT.x = InvT.x * tau_q;  // T^T[g][c] = Inv[c][g] * tau_g
  TT.y = InvT.y * tau_q;
}

// C_i^T -= ((C_i^T V_P) * TB) V_P^T for the column blocks i > P; rows j >= J0 of C_i enter the dot products,
// rows j >= P are updated.  TB = T^T (accumulator layout) in the R phase, T in the Q phase.
template <int P, int J0>
__device__ __forceinline__ void qb_trailing(double (&a)[4][8][2], const double* vs, int g, int t, const Acc& TB) {
#pragma unroll
  for (int i = P + 1; i < 4; i++) {
    Acc W{0.0, 0.0}, W2{0.0, 0.0};
#pragma unroll
    for (int j = J0; j < 8; j++) {  // two independent accumulation chains
      dmma884(W.x, W.y, a[i][j][0], a[P][j][0]);
      dmma884(W2.x, W2.y, a[i][j][1], a[P][j][1]);
    }
    W.x += W2.x;
    W.y += W2.y;
    Acc Z = mm8(W, TB);
    Z.x = -Z.x;
    Z.y = -Z.y;
#pragma unroll
    for (int j = P; j < 8; j++) {
      const double2 v = *reinterpret_cast<const double2*>(vs + (8 * j + g) * kVS + 8 * P + 2 * t);
      dmma884(a[i][j][0], a[i][j][1], Z.x, v.x);
      dmma884(a[i][j][0], a[i][j][1], Z.y, v.y);
    }
  }
}

// R rows 8P..8P+7 (rows with beta < 0 negated, times post) for column block I.
template <int P, int I>
__device__ __forceinline__ void qb_store_r(const double (&a)[4][8][2], double* __restrict__ r_out, int g, int t,
                                           double s0, double s1) {
  const int row0 = 8 * P + 2 * t, col = 8 * I + g;
  double v0, v1;
  if (I < P) { v0 = 0.0; v1 = 0.0; }
  else if (I == P) { v0 = (2 * t <= g) ? s0 * a[I][P][0] : 0.0; v1 = (2 * t + 1 <= g) ? s1 * a[I][P][1] : 0.0; }
  else { v0 = s0 * a[I][P][0]; v1 = s1 * a[I][P][1]; }
  r_out[row0 * 32 + col] = v0;
  r_out[(row0 + 1) * 32 + col] = v1;
}

template <int P>
__device__ __forceinline__ void qb_r_phase(double (&a)[4][8][2], double* vs, double* ts,
                                           double* __restrict__ r_out, int lane, int g, int t, double post, double& sgn_p) {
  double tau_q;
  qb_panel<P>(a, lane, g, t, tau_q, sgn_p);
  const double s0 = shfl(sgn_p, 4 * (2 * t)) * post, s1 = shfl(sgn_p, 4 * (2 * t + 1)) * post;
  // R: diagonal block and the zero blocks left of it; then the head of the panel becomes the clean unit-lower V
  if (P > 0) qb_store_r<P, 0>(a, r_out, g, t, s0, s1);
  if (P > 1) qb_store_r<P, 1>(a, r_out, g, t, s0, s1);
  if (P > 2) qb_store_r<P, 2>(a, r_out, g, t, s0, s1);
  qb_store_r<P, P>(a, r_out, g, t, s0, s1);
  a[P][P][0] = (2 * t > g) ? a[P][P][0] : ((2 * t == g) ? 1.0 : 0.0);
  a[P][P][1] = (2 * t + 1 > g) ? a[P][P][1] : ((2 * t + 1 == g) ? 1.0 : 0.0);
#pragma unroll
  for (int j = P; j < 8; j++) {
    vs[(8 * j + 2 * t) * kVS + 8 * P + g] = a[P][j][0];
    vs[(8 * j + 2 * t + 1) * kVS + 8 * P + g] = a[P][j][1];
  }
  Acc T, TT;
  qb_make_t<P>(a, g, t, tau_q, T, TT);
  *reinterpret_cast<double2*>(ts + 64 * P + 2 * lane) = make_double2(T.x, T.y);
  __syncwarp();
  if (P < 3) {
    qb_trailing<P, P>(a, vs, g, t, TT);
    if (P < 1) qb_store_r<P, 1>(a, r_out, g, t, s0, s1);
    if (P < 2) qb_store_r<P, 2>(a, r_out, g, t, s0, s1);
    qb_store_r<P, 3>(a, r_out, g, t, s0, s1);
  }
}

template <int P>
__device__ __forceinline__ void qb_q_phase(double (&a)[4][8][2], const double* vs, const double* ts,
                                           int lane, int g, int t) {
  const double2 tt = *reinterpret_cast<const double2*>(ts + 64 * P + 2 * lane);
  const Acc T{tt.x, tt.y};
  // columns right of the panel: nonzero from row 8(P+1) on, filled from row 8P on by this update
  if (P < 3) qb_trailing<P, P + 1>(a, vs, g, t, T);
  // the panel's own columns: E_P - V_P (T_P L_P^T)
  const double2 l = *reinterpret_cast<const double2*>(vs + (8 * P + g) * kVS + 8 * P + 2 * t);
  Acc M = mm8(Acc{l.x, l.y}, T);  // (L_P T_P^T)[b][c] = M^T
  M.x = -M.x;
  M.y = -M.y;
#pragma unroll
  for (int j = 0; j < P; j++) { a[P][j][0] = 0.0; a[P][j][1] = 0.0; }
#pragma unroll
  for (int j = P; j < 8; j++) {
    const double2 v = *reinterpret_cast<const double2*>(vs + (8 * j + g) * kVS + 8 * P + 2 * t);
    double q0 = (j == P && g == 2 * t) ? 1.0 : 0.0, q1 = (j == P && g == 2 * t + 1) ? 1.0 : 0.0;
    dmma884(q0, q1, M.x, v.x);
    dmma884(q0, q1, M.y, v.y);
    a[P][j][0] = q0;
    a[P][j][1] = q1;
  }
}

__global__ void __launch_bounds__(kQbWarps * 32, 2)
qr64x32_blocked_kernel(const double* __restrict__ A, double* __restrict__ Q, double* __restrict__ R, int64_t batch) {
  extern __shared__ __align__(16) double qb_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t m = (int64_t)blockIdx.x * kQbWarps + warp;
  if (m >= batch) return;  // warp-uniform; no block-level barriers below
  double* vs = qb_smem + warp * kQbWarpDoubles;
  double* ts = vs + 64 * kVS;
  const double* a_in = A + m * 2048;

  double a[4][8][2];
#pragma unroll
  for (int j = 0; j < 8; j++)
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int i = 0; i < 4; i++) a[i][j][e] = ldg1_stream(a_in + (8 * j + 2 * t + e) * 32 + 8 * i + g);

  // scale guard (see pow2_prescale)
  double amax = 0.0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) amax = fmax(amax, fmax(fabs(a[i][j][0]), fabs(a[i][j][1])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, shfl_xor(amax, o));
  const double pre = pow2_prescale(amax);
  if (pre != 1.0) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) { a[i][j][0] *= pre; a[i][j][1] *= pre; }
  }
  const double post = 1.0 / pre;

  double* r_out = R + m * 1024;
  double sgn[4];
  qb_r_phase<0>(a, vs, ts, r_out, lane, g, t, post, sgn[0]);
  qb_r_phase<1>(a, vs, ts, r_out, lane, g, t, post, sgn[1]);
  qb_r_phase<2>(a, vs, ts, r_out, lane, g, t, post, sgn[2]);
  qb_r_phase<3>(a, vs, ts, r_out, lane, g, t, post, sgn[3]);

  qb_q_phase<3>(a, vs, ts, lane, g, t);
  qb_q_phase<2>(a, vs, ts, lane, g, t);
  qb_q_phase<1>(a, vs, ts, lane, g, t);
  qb_q_phase<0>(a, vs, ts, lane, g, t);

  double* q_out = Q + m * 2048;
#pragma unroll
  for (int j = 0; j < 8; j++)
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int i = 0; i < 4; i++) q_out[(8 * j + 2 * t + e) * 32 + 8 * i + g] = sgn[i] * a[i][j][e];
}

}  // namespace

cudaError_t launch_qr64x32_blocked(cudaStream_t s, const double* A, double* Q, double* R, int64_t batch) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(qr64x32_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQbSmem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  qr64x32_blocked_kernel<<<(unsigned)((batch + kQbWarps - 1) / kQbWarps), kQbWarps * 32, kQbSmem, s>>>(A, Q, R, batch);
  return cudaGetLastError();
}

}  // namespace nd4b
