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
#include <cstdlib>

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

// Shared tail of the 64x64 kernels.  On entry G (converged, columns mutually orthogonal) and the accumulated V are
// column-major in shared memory (column stride LD); computes sigma, the stable descending order, U = G diag(1/sigma)
// (zero columns completed to an orthonormal basis) and writes U, sv, V.  Called by all 256 threads after a barrier.
__device__ __forceinline__ void svd64_epilogue(double* Gs, double* Vs, double* sq, int* perm, int* zero_flag, int* done_flag,
                                               double* __restrict__ U, double* __restrict__ SV, double* __restrict__ V, int64_t m) {
  constexpr int N = 64, LD = kSvd64LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7, P = warp * 4 + (lane >> 3);
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

__global__ void __launch_bounds__(kSvd64Threads, 3)
svd64_smem_kernel(const double* __restrict__ A, double* __restrict__ U, double* __restrict__ SV, double* __restrict__ V,
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

  svd64_epilogue(Gs, Vs, sq, perm, zero_flag, done_flag, U, SV, V, m);
}

// ------------------------------------------------------------------------------------------------
// 64x64, register-resident: the production kernel.
//
// The shared-memory kernel above moves 4 KiB through shared memory per rotated pair and is bound by that
// bandwidth (ncu: l1tex 88 %, fp64 pipe 44 %).  Here every rotation is local to a thread:
//   * thread (row, q) of 256 holds a quarter (16 column slots) of row `row` of G and of row `row` of V in
//     registers; a column-pair rotation touches two registers per row, no data moves;
//   * pairs are visited in the odd-even (Brent-Luk) order with exchange: step A rotates slots (2i,2i+1), step B
//     slots (2i+1,2i+2), and the two rotated columns are written back swapped.  64 steps are one sweep
//     (every pair of columns meets exactly once, 2016 pairs), and the slot pattern has period 2, so all
//     register indices are static without unrolling the sweep;
//   * the only cross-thread quantities are the 32 (31) dot products g_p.g_q of a step: 8 products per thread,
//     recursive halving over the 8 row-lanes of a warp (7 shuffles), 8 warp partials combined through smem;
//   * column norms are cached per slot (a' = a - t d, b' = b + t d, Rutishauser) and recomputed exactly at the
//     start of every sweep; one lane per pair computes (c, s), so rotation set-up is not replicated;
//   * (c, s) of a step are broadcast through shared memory (uniform 16-byte loads);
//   * lanes are quarter-major (q = lane>>3) so that a quarter warp reads one (c, s) word per load.
// ------------------------------------------------------------------------------------------------
constexpr int kSvdR_DPART = 8 * 64;  // warp partials: 8 warps x (32 pair products | 64 slot norms)
constexpr size_t kSvd64RegSmem = sizeof(double) * (2 * 64 * kSvd64LD + 64 /*sq*/ + kSvdR_DPART + 2 * 64 /*cs*/ + 64 /*nslot*/) +
                                 sizeof(int) * (64 * 3 + 4);

// v[0..N/2) <- v[keep half] + partner's contribution; lanes with `bit` set keep the upper half
template <int N>
__device__ __forceinline__ void halve(double (&v)[N], bool bit, int mask) {
#pragma unroll
  for (int k = 0; k < N / 2; k++) {
    const double send = bit ? v[k] : v[k + N / 2];
    const double keep = bit ? v[k + N / 2] : v[k];
    v[k] = keep + shfl_xor(send, mask);
  }
}

__device__ __forceinline__ void rot_swap(double& xa, double& xb, double c, double s) {
  // (p, q) -> (c p - s q, s p + c q), stored exchanged: first slot <- new q, second slot <- new p
  const double np = fma(c, xa, -(s * xb));
  const double nq = fma(s, xa, c * xb);
  xa = nq;
  xb = np;
}

// One odd-even step for all 256 threads.  STEP_B selects the slot pattern; pw is the parameter warp.
template <bool STEP_B>
__device__ __forceinline__ void svd64_step(double (&g)[16], double (&v)[16], double* dpart, double2* cs_now, double* nslot,
                                           int* flags, int warp, int lane, int rl, int q, int pw, double tol2) {
  constexpr int N = 64;
  const bool b4 = rl & 4, b3 = rl & 2, b2 = rl & 1;
  // ---- P1: the dot products of this step's pairs ----
  double gR0 = 0.0, gL15 = 0.0, vR0 = 0.0, vL15 = 0.0;
  double pd[8];
  if (!STEP_B) {
#pragma unroll
    for (int j = 0; j < 8; j++) pd[j] = g[2 * j] * g[2 * j + 1];
  } else {
    gR0 = __shfl_down_sync(kFull, g[0], 8);   // lane + 8 = (rl, q + 1)
    gL15 = __shfl_up_sync(kFull, g[15], 8);   // lane - 8 = (rl, q - 1)
    vR0 = __shfl_down_sync(kFull, v[0], 8);
    vL15 = __shfl_up_sync(kFull, v[15], 8);
#pragma unroll
    for (int j = 0; j < 7; j++) pd[j] = g[2 * j + 1] * g[2 * j + 2];
    pd[7] = (q < 3) ? g[15] * gR0 : 0.0;
  }
  halve<8>(pd, b4, 4);
  {
    double h4[4];
#pragma unroll
    for (int k = 0; k < 4; k++) h4[k] = pd[k];
    halve<4>(h4, b3, 2);
    double h2[2] = {h4[0], h4[1]};
    halve<2>(h2, b2, 1);
    dpart[warp * 64 + 8 * q + rl] = h2[0];  // pair 8q + rl, rows of this warp
  }
  __syncthreads();  // B1

  // ---- P2: one lane per pair: threshold test, (c, s), cached-norm update ----
  if (warp == pw) {
    const int i = lane;
    double d = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) d += dpart[w * 64 + i];
    const int sp = STEP_B ? 2 * i + 1 : 2 * i, sqq = sp + 1;
    const bool have = sqq < N;
    const double na = nslot[sp], nb = have ? nslot[sqq] : 0.0;
    double c = 1.0, s = 0.0;
    if (have && d * d > tol2 * na * nb) {
      const double num = nb - na, den = 2.0 * d;
      const double h = sqrt(fma(num, num, den * den));
      double t = fabs(den) / (fabs(num) + h);
      if ((num < 0.0) != (den < 0.0)) t = -t;
      c = rsqrt(fma(t, t, 1.0));
      s = c * t;
      nslot[sp] = fmax(fma(t, d, nb), 0.0);  // exchanged: first slot now holds the rotated q column
      nslot[sqq] = fmax(fma(-t, d, na), 0.0);
      flags[0] = 1;
    } else if (have) {
      nslot[sp] = nb;
      nslot[sqq] = na;
    }
    cs_now[4 * (i & 7) + (i >> 3)] = make_double2(c, s);
  }
  __syncthreads();  // B2

  // ---- rotate G and V with the same coefficients ----
  if (!STEP_B) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const double2 r = cs_now[4 * j + q];
      rot_swap(g[2 * j], g[2 * j + 1], r.x, r.y);
      rot_swap(v[2 * j], v[2 * j + 1], r.x, r.y);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 7; j++) {
      const double2 r = cs_now[4 * j + q];
      rot_swap(g[2 * j + 1], g[2 * j + 2], r.x, r.y);
      rot_swap(v[2 * j + 1], v[2 * j + 2], r.x, r.y);
    }
    if (q < 3) {  // pair 8q+7 = (my slot 15, right neighbour's slot 0): first slot <- s p + c q
      const double2 r = cs_now[28 + q];
      g[15] = fma(r.y, g[15], r.x * gR0);
      v[15] = fma(r.y, v[15], r.x * vR0);
    }
    if (q > 0) {  // pair 8q-1 = (left neighbour's slot 15, my slot 0): second slot <- c p - s q
      const double2 r = cs_now[27 + q];
      g[0] = fma(r.x, gL15, -(r.y * g[0]));
      v[0] = fma(r.x, vL15, -(r.y * v[0]));
    }
  }
}

__global__ void __launch_bounds__(kSvd64Threads, 2)
svd64_kernel(const double* __restrict__ A, double* __restrict__ U, double* __restrict__ SV, double* __restrict__ V,
             int64_t batch, int* sweeps_out, int* fail_out) {
  constexpr int N = 64, LD = kSvd64LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Gs = reinterpret_cast<double*>(smem_raw);
  double* Vs = Gs + N * LD;
  double* sq = Vs + N * LD;
  double* dpart = sq + N;                                          // [8][64]
  double2* cs = reinterpret_cast<double2*>(dpart + kSvdR_DPART);   // [2][32], index 4*(i&7) + (i>>3) for pair i
  double* nslot = reinterpret_cast<double*>(cs + 64);              // [64] cached |g_slot|^2
  int* perm = reinterpret_cast<int*>(nslot + N);
  int* zero_flag = perm + N;
  int* done_flag = zero_flag + N;
  int* flags = done_flag + N;                                      // [0] = "some pair was rotated in this sweep"

  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane >> 3, rl = lane & 7, row = 8 * warp + rl;    // quarter-major lanes: a quarter warp shares q
  const double* a_in = A + m * (N * N) + row * N + 16 * q;

  double g[16], v[16];
#pragma unroll
  for (int s = 0; s < 16; s += 2) {
    const double2 t = ldg2_stream(a_in + s);
    g[s] = t.x;
    g[s + 1] = t.y;
  }
#pragma unroll
  for (int s = 0; s < 16; s++) v[s] = (16 * q + s == row) ? 1.0 : 0.0;
  if (tid == 0) flags[0] = 0;

  const double tol2 = (N * kEps) * (N * kEps);
  int sweeps = 0;
  bool converged = false;
  while (sweeps < kMaxSweeps && !converged) {
    sweeps++;
    // ---- exact slot norms at the start of the sweep ----
    {
      double n2[16];
#pragma unroll
      for (int s = 0; s < 16; s++) n2[s] = g[s] * g[s];
      halve<16>(n2, rl & 4, 4);
      double h8[8];
#pragma unroll
      for (int k = 0; k < 8; k++) h8[k] = n2[k];
      halve<8>(h8, rl & 2, 2);
      double h4[4];
#pragma unroll
      for (int k = 0; k < 4; k++) h4[k] = h8[k];
      halve<4>(h4, rl & 1, 1);
      dpart[warp * 64 + 16 * q + 2 * rl] = h4[0];  // slots 16q + 2rl + {0,1}
      dpart[warp * 64 + 16 * q + 2 * rl + 1] = h4[1];
      __syncthreads();
      if (tid < N) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += dpart[w * 64 + tid];
        nslot[tid] = t;
      }
      __syncthreads();
    }
#pragma unroll 1
    for (int sp2 = 0; sp2 < N / 2; sp2++) {
      svd64_step<false>(g, v, dpart, cs, nslot, flags, warp, lane, rl, q, (2 * sp2) & 7, tol2);
      svd64_step<true>(g, v, dpart, cs + 32, nslot, flags, warp, lane, rl, q, (2 * sp2 + 1) & 7, tol2);
    }
    __syncthreads();
    converged = (flags[0] == 0);
    __syncthreads();
    if (tid == 0) flags[0] = 0;
  }
  if (tid == 0) {
    if (sweeps_out) atomicMax(sweeps_out, sweeps);
    if (!converged && fail_out) atomicExch(fail_out, 1);
  }

  // ---- registers -> column-major shared memory, then the common epilogue ----
#pragma unroll
  for (int s = 0; s < 16; s++) {
    Gs[(16 * q + s) * LD + row] = g[s];
    Vs[(16 * q + s) * LD + row] = v[s];
  }
  __syncthreads();
  svd64_epilogue(Gs, Vs, sq, perm, zero_flag, done_flag, U, SV, V, m);
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
      cudaError_t e = cudaFuncSetAttribute(svd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSvd64RegSmem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(svd64_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSvd64Smem);
      if (e != cudaSuccess) return e;
      attr_set[dev] = true;
    }
    static int variant = -1;
    if (variant < 0) { const char* ev = getenv("ND4B_SVD_VARIANT"); variant = ev ? atoi(ev) : 0; }
    if (variant == 1) svd64_smem_kernel<<<(unsigned)batch, kSvd64Threads, kSvd64Smem, s>>>(A, U, sv, V, batch, sweeps, fail);
    else svd64_kernel<<<(unsigned)batch, kSvd64Threads, kSvd64RegSmem, s>>>(A, U, sv, V, batch, sweeps, fail);
    return cudaGetLastError();
  }
  const size_t need = svd_workspace_bytes(batch, rows, cols);
  if (work == nullptr || work_bytes < need) return cudaErrorInvalidValue;
  svd_generic_kernel<<<(unsigned)batch, kSvdGenThreads, 0, s>>>(A, U, sv, V, batch, rows, cols, sweeps, fail, work);
  return cudaGetLastError();
}

}  // namespace nd4b
