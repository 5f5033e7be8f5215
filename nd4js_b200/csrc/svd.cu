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
//  * svd64cb_kernel    : 64x64, production: one CTA (4 warps) per matrix, G and V in registers (a warp owns 16 whole
//                        columns), odd-even ordering with exchange, square-root-free scaled rotations (2 FMAs per
//                        element pair), dot products reduced inside the warp; see the section header below.
//  * svd64_smem_kernel : 64x64, first version kept as the A/B baseline (ND4B_SVD_VARIANT=1): G and V column-major in
//                        shared memory, 8 lanes per column pair, round-robin ordering.
//  * svd_generic_kernel: any shape; one CTA per matrix, G^T and V^T in global (L2-resident) scratch.
#include "common.cuh"
#include "kernels.h"
#include <float.h>
#include <cstdlib>

namespace nd4b {

constexpr int kMaxSweeps = 60;
constexpr double kEps = 2.220446049250313e-16;

// Numerical-zero rule of the epilogues: a column whose norm is below 2^-200 of the largest one is a zero singular value
// (its U vector comes from the orthonormal completion).  Together with the unit prescale below (largest entry in [1, 2)) and
// the guard in the rotation test this keeps every product of the threshold test `d^2 > tol^2 |p|^2 |q|^2` in the normal
// range: without it, columns that cancellation left at 1e-150 of the others (sparse rank-deficient input,
// _generic_test_svd_decomp.js:257-274) were never orthogonalised against each other — d^2 underflows — and U = G / sigma
// lost orthogonality.
constexpr double kZeroRel = 0x1p-200;
constexpr double kNormMin = 0x1p-400; // = kZeroRel^2: a column with |g|^2 below this (the largest entry is in [1, 2)) is numerically
                                      // zero and is left alone — two parallel columns of a sparse matrix otherwise shed a rounding residue
                                      // sixteen orders of magnitude smaller sweep after sweep without ever becoming orthogonal

// Exact power of two that brings the largest entry into [1, 2) (QR / SVD are scale-equivariant and every operation of the
// kernels commutes with a power-of-two scaling, so results do not change; zero, Inf and NaN give 1).
__device__ __forceinline__ double pow2_prescale_unit(double amax) {
  if (!(amax > 0.0) || !(amax < CUDART_INF)) return 1.0;
  const int e = -ilogb(amax);
  return scalbn(1.0, e > 1000 ? 1000 : e);
}

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

// Reciprocal square root for normal-range arguments without the library's special-case branches: hardware seed
// (MUFU.RSQ64H, ~2^-22) and two third-order Newton steps (full double precision up to ~1 ulp).
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double e = fma(-x * y, y, 1.0);   // 1 - x y^2
    y = fma(y * e, fma(0.375, e, 0.5), y);  // y (1 + e/2 + 3 e^2 / 8)
  }
  return y;
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
    if (!zero_flag[z] || done_flag[z]) continue;   // done on entry: a zero column the caller does not need
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
// PADDED: the matrix is rows x cols (both <= 64) inside the zero-padded 64 x 64 tile; only the L = min(rows, cols) leading
// singular triplets are written (U rows x L, sv L, V L x cols), and only zero columns among those are completed.
// Does column k (norm sk) come before column j (norm sj) in the output order?  Descending, stable, and TOTAL: a NaN norm
// ranks before every number (ties among NaNs by index), so that the ranks are a permutation whatever the input — with the
// plain comparisons a NaN column shares rank 0 with the largest one and a slot of perm[] is never written.
__device__ __forceinline__ bool sv_before(double sk, int k, double sj, int j) {
  const bool nk = sk != sk, nj = sj != sj;
  if (nk || nj) return nk && (!nj || k < j);
  return (sk > sj) || (sk == sj && k < j);
}

template <int T, bool PADDED = false>
__device__ __forceinline__ void svd64_epilogue(double* Gs, double* Vs, double* sq, int* perm, int* zero_flag, int* done_flag,
                                               double* __restrict__ U, double* __restrict__ SV, double* __restrict__ V, int64_t m,
                                               double sigma_scale, int rows = 64, int cols = 64) {
  constexpr int N = 64, LD = kSvd64LD;
  const int L = PADDED ? (rows < cols ? rows : cols) : N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NG = T / 8;  // 8-lane groups in the CTA; each handles columns P, P+NG, ...
  const int sub = lane & 7, P = warp * 4 + (lane >> 3);
  // singular values: column norms
#pragma unroll
  for (int h = 0; h < N / NG; h++) {
    const int col = P + NG * h;
    const double2* g = reinterpret_cast<const double2*>(Gs + col * LD + 2 * sub);
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const double2 x = g[8 * k]; a = fma(x.x, x.x, a); a = fma(x.y, x.y, a); }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) a += shfl_xor(a, o);
    if (sub == 0) sq[col] = sqrt(a);
  }
  __syncthreads();
  // PADDED: the exchanges of the odd-even ordering move every column through the slots, so the L real columns and the
  // padding columns are interleaved by now.  A padding column never rotated: its V vector is still a unit vector e_k
  // with k >= L, while the V vectors of the real columns are exactly zero there.  Padding columns rank behind all real ones.
  bool pad = false;
  if (PADDED) {
    if (tid < N)
      for (int k = L; k < N; k++) pad = pad || (Vs[tid * LD + k] != 0.0);
    if (tid < N) zero_flag[tid] = pad ? 1 : 0;   // scratch until the real flags are written below
    __syncthreads();
  }
  // stable descending rank + zero detection
  int rank = 0, z = 0;
  if (tid < N) {
    const double sj = sq[tid];
    double smax = 0.0;
    int pads_before = 0;
    for (int k = 0; k < N; k++) {
      const double sk = sq[k];
      const bool pk = PADDED && zero_flag[k];
      if (!pk) rank += sv_before(sk, k, sj, tid);
      else pads_before += (k < tid);
      smax = fmax(smax, sk);
    }
    if (pad) rank = L + pads_before;
    perm[rank] = tid;
    z = (sj < smax * kZeroRel) || (sj < DBL_MIN)   /* a NaN norm is not a zero: NaN in, NaN out */;
  }
  if (PADDED) __syncthreads();
  if (tid < N) {
    zero_flag[tid] = z;
    done_flag[tid] = pad ? 1 : 0;   // padding columns stay zero vectors: nothing to complete, never written
  }
  __syncthreads();
  // normalise the columns of G in place -> U columns
#pragma unroll
  for (int h = 0; h < N / NG; h++) {
    const int col = P + NG * h;
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
    if (warp == 0) complete_basis_warp(Gs, LD, PADDED ? (rows > cols ? rows : cols) : N, N, zero_flag, done_flag);
    __syncthreads();
  }

  if (!PADDED) {
    double* u_out = U + m * (N * N);
    double* v_out = V + m * (N * N);
    for (int e = tid; e < N * N; e += T) {
      const int i = e >> 6, l = e & 63;
      u_out[e] = Gs[perm[l] * LD + i];   // U[i][l]
      v_out[e] = Vs[perm[i] * LD + l];   // V[l'][j] with l' = i, j = l
    }
    if (tid < N) SV[m * N + tid] = zero_flag[perm[tid]] ? 0.0 : sq[perm[tid]] * sigma_scale;
  } else {
    double* u_out = U + m * (rows * L);
    double* v_out = V + m * (L * cols);
    // tall / square: A = (G normalised) S V_acc^T.  Wide: A^T was factorised, A = V_acc S (G normalised)^T.
    const double* usrc = rows >= cols ? Gs : Vs;
    const double* vsrc = rows >= cols ? Vs : Gs;
    for (int e = tid; e < rows * L; e += T) {
      const int i = e / L, l = e - i * L;
      u_out[e] = usrc[perm[l] * LD + i];
    }
    for (int e = tid; e < L * cols; e += T) {
      const int l = e / cols, j = e - l * cols;
      v_out[e] = vsrc[perm[l] * LD + j];
    }
    if (tid < L) SV[m * L + tid] = zero_flag[perm[tid]] ? 0.0 : sq[perm[tid]] * sigma_scale;
  }
}

__global__ void __launch_bounds__(kSvd64Threads, 3)
svd64_smem_kernel(const double* __restrict__ A, double* __restrict__ U, double* __restrict__ SV, double* __restrict__ V,
             int64_t batch, int* sweeps_out, int* fail_out, unsigned long long* sweep_sum) {
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
      if (d * d > tol2 * a * b && a > kNormMin && b > kNormMin) {  // uniform inside the 8-lane group
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
    if (sweep_sum) atomicAdd(sweep_sum, (unsigned long long)sweeps);
    if (!converged && fail_out) atomicExch(fail_out, 1);
  }

  svd64_epilogue<kSvd64Threads>(Gs, Vs, sq, perm, zero_flag, done_flag, U, SV, V, m, 1.0);
}

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

// ------------------------------------------------------------------------------------------------
// 64x64, register-resident, column-block per warp ("CB").
//
// Design history (profiles/): with G and V in shared memory a rotated pair moves 4 KiB through smem and the
// kernel is smem-bandwidth bound (svd64_smem_kernel above, 40 ms on C5); with rows split over several warps every step
// needs a cross-warp reduction, two CTA barriers and one parameter warp everybody waits for (29 ms).  Here every
// rotation is thread-local and a step needs no shared memory or barrier: a warp owns 16 whole columns
// (slots): lane l holds rows l and l+32 of G and of V for those 16 slots.  The 8 dot products of a step are
// reduced inside the warp (recursive halving over lane bits 4,3,2, then a 2-stage butterfly), the 4-lane group g
// computes the rotation of pair g, and the 8 (c, s) are gathered with shuffles: no shared memory, no barrier and no
// central parameter computation inside a step.  The odd-even pairing shifts by one slot between A and B steps, so the
// ownership boundary migrates instead: before a B step each warp hands its first column to its left neighbour, after
// it the (rotated) column comes back — one 4-double-per-lane exchange through shared memory and one barrier per step.
// ------------------------------------------------------------------------------------------------
//
// Scaled ("fast", square-root-free) rotations: column p is stored as g_p = D_p * gh_p with one scalar D_p per
// column.  The plane rotation  g_p' = c g_p - s g_q,  g_q' = s g_p + c g_q  then becomes
//     gh_p' = gh_p - alpha gh_q,   gh_q' = gh_q + beta gh_p,   D_p' = c D_p,  D_q' = c D_q,
//     alpha = t D_q / D_p,  beta = t D_p / D_q,  t = s / c,
// i.e. TWO independent FMAs per element pair instead of two multiplications and two FMAs; V shares the D of G
// because it undergoes the same rotations from D = 1.  |t| <= 1, so c >= 1/sqrt 2: within one sweep (63 rotations per
// column) D stays above 3e-10 and |gh| grows by at most 2^31.5; D is folded back into the columns after every sweep.
// D and 1/D are both tracked (1/c is a by-product of the rotation set-up), so no division is needed.
struct CbState {
  double g0[16], g1[16], v0[16], v1[16];  // rows l and l+32 of G-hat and V-hat for the warp's 16 slots
  double xg0, xg1, xv0, xv1;              // the borrowed boundary column (right neighbour's first slot) in B steps
  double e, o;                            // cached TRUE |g|^2 of slots 2*grp and 2*grp+1 (replicated in the 4-lane group)
  double de, dei, dq, dqi;                // scale D and 1/D of the even slot (de, dei) and of the odd slot (dq, dqi)
  double xn, xd, xdi;                     // norm and scale of the borrowed column
};

struct CbRot { double alpha, beta; };

// One lane per pair: threshold test and rotation set-up from the true norms and the true dot product.
// Outputs the scaled-rotation multipliers, and the norms/scales of the two columns AFTER the rotation (before exchange).
__device__ __forceinline__ CbRot cb_params(double dhat, double na, double nb, double Dp, double Dpi, double Dq, double Dqi,
                                           bool have, double tol2, double& na2, double& nb2, double& Dp2, double& Dpi2,
                                           double& Dq2, double& Dqi2, int& rotated, int& unsafe) {
  CbRot r;
  r.alpha = 0.0; r.beta = 0.0;
  na2 = na; nb2 = nb; Dp2 = Dp; Dpi2 = Dpi; Dq2 = Dq; Dqi2 = Dqi;
  const double d = dhat * Dp * Dq;
  // The set-up runs warp-uniformly (if any pair of the warp rotates, every lane executes it and pairs below the threshold
  // discard the result with selects): a divergent branch here costs a reconvergence barrier in every step.
  // both squared norms above kNormMin = 2^-400, as one integer compare of the high words (they are never negative)
  const double dd = d * d, thr = tol2 * na * nb;
  const bool rot = have && dd > thr && min(__double2hiint(na), __double2hiint(nb)) > 0x26f00000;
  if (__any_sync(kFull, rot)) {
    const double num = nb - na, den = 2.0 * d;
#ifndef ND4B_SVD_T_FP64
    // Critical path of a step: t = tan(theta) = sign * |den| / (|num| + sqrt(num^2 + den^2)).  Any t gives an exactly
    // orthogonal rotation as long as c = 1/sqrt(1 + t^2) is formed from that same t, and the off-diagonal element left
    // behind is only (relative error of t) * d — so t is computed in single precision (2^-22: the convergence of the
    // sweeps is unchanged to within a quarter of a sweep on average), on the FP32 pipe and the special-function unit
    // with 4-cycle dependent latency instead of a chain of sixteen FP64 operations.  num and den are first scaled by a
    // power of two taken from the larger exponent (exact), so that the single-precision range is never left.
    const int hm = max(__double2hiint(num) & 0x7ff00000, __double2hiint(den) & 0x7ff00000);
    const double scale = __hiloint2double(max(0x7fe00000 - hm, 0x00100000), 0);
    const float nf = (float)(num * scale), df = (float)(den * scale);   // the larger magnitude lands in [1, 2)
    const float an = fabsf(nf), ad = fabsf(df);
    const float Sf = fmaf(an, an, ad * ad);                              // in [1, 8)
    float yf;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(Sf));
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(fmaf(Sf, yf, an)));   // 1 / (|num| + sqrt(S)), denominator >= 1
    // t = den / (|num| + sqrt(S)) with the quotient's single-precision reciprocal applied to the DOUBLE numerator: when
    // |den| is more than 2^126 times smaller than |num| (a rounding residue that is still parallel to a large column:
    // sparse rank-deficient input) the single-precision numerator would flush to zero, the rotation would be the
    // identity, and the pair would be "rotated" in every sweep for ever
    const float rfs = __int_as_float(__float_as_int(rf) | (__float_as_int(nf) & 0x80000000));
    const double t = (den * scale) * (double)rfs;
    // c = 1/sqrt(1 + t^2) and 1/c = sqrt(1 + t^2) in full precision from the t actually used (off the critical path)
    const double w = fma(t, t, 1.0);                                     // in [1, 2]
    double c;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(c) : "d"(w));
    {
      const double e1 = fma(-w * c, c, 1.0);
      c = fma(c * e1, fma(0.375, e1, 0.5), c);
    }
    const double rc = w * c;
#else
    double numx = num, denx = den;
    if (fabs(numx) + fabs(denx) < 1e-140) { numx *= 0x1p600; denx *= 0x1p600; }  // only the ratio matters; keep num^2+den^2 normal
    // FP64 variant (A/B): t = sin(2 theta) / (1 + cos(2 theta)).  1/sqrt(S) by one cubic step from the
    // hardware seed y0 (2^-22 -> ~2^-66); the reciprocal of w = 1 + cos(2 theta) in [1,2] by two Newton steps from a seed
    // taken at the approximate w, so that the two special-function latencies overlap.
    const double S = fma(numx, numx, denx * denx);
    const double an = fabs(numx), ad = fabs(denx);
    double y0, r0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(S));
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(fma(an, y0, 1.0)));
    const double e = fma(-S * y0, y0, 1.0);
    const double rh = fma(y0 * e, fma(0.375, e, 0.5), y0);
    const double c2 = an * rh, s2 = ad * rh;                 // cos 2theta, |sin 2theta|
    const double w = 1.0 + c2;                               // 2 c^2
    const double r1 = fma(fma(-w, r0, 1.0), r0, r0);
    const double rw = fma(fma(-w, r1, 1.0), r1, r1);         // 1 / w
    double t = s2 * rw;
    if ((numx < 0.0) != (denx < 0.0)) t = -t;
    const double hc = 0.5 * w;                               // c^2
    double rc;                                               // 1/c, one cubic step (hc in [0.5, 1])
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(hc));
    {
      const double e1 = fma(-hc * rc, rc, 1.0);
      rc = fma(rc * e1, fma(0.375, e1, 0.5), rc);
    }
    const double c = hc * rc;
#endif
    // norms after the annihilating rotation: |c p - s q|^2 = |p|^2 - t d,  |s p + c q|^2 = |q|^2 + t d
    // (from (1 - t^2) d + t (|p|^2 - |q|^2) = 0); the cached norms are refreshed exactly every sweep
    const double na_r = fmax(fma(-t, d, na), 0.0);
    const double nb_r = fmax(fma(t, d, nb), 0.0);
    r.alpha = rot ? t * Dq * Dpi : 0.0;
    r.beta = rot ? t * Dp * Dqi : 0.0;
    Dp2 = rot ? c * Dp : Dp; Dq2 = rot ? c * Dq : Dq; Dpi2 = rot ? rc * Dpi : Dpi; Dqi2 = rot ? rc * Dqi : Dqi;
    na2 = rot ? na_r : na;
    nb2 = rot ? nb_r : nb;
    rotated = 1;
    // Quadratic-convergence exit.  With p' = c p - s q, the cosine of p' with any third column r changes by
    // s cos(q, r) |q| / |p'| (and that of q' by s cos(p, r) |p| / |q'|): at most tau = |t| max(|q|/|p|, |p|/|q|) times the
    // largest cosine present, and a column takes part in 63 rotations per sweep.  So when 126 * max tau * max|cos| over
    // the pairs rotated in a sweep stays below 64 eps, no cosine can exceed 64 eps afterwards — exactly what the
    // confirming sweep would establish — and that sweep is not run.  (For columns of very different norms tau is about
    // the cosine itself, not t: a null column of a rank-deficient matrix keeps the confirming sweep.)  The two maxima
    // are tracked as integers, on the high words: tol2 = 2^-92, so hi(d^2) - hi(thr) is 2^20 (log2 cos^2 + 92), hi(|t|)
    // is 2^20 (log2 |t| + 1023) and |hi(|q|^2) - hi(|p|^2)| / 2 is 2^20 log2 of the norm ratio; `unsafe` packs the cos^2
    // term in the upper and the tau term in the lower 16 bits, in sixteenths of a binade (test: svd64cb_kernel).
    if (rot) {
      const int e1 = max(__double2hiint(dd) - __double2hiint(thr), 0) >> 16;                       // 16 (log2 cos^2 + 92), >= 0 when rotated
      const int e2 = max((__double2hiint(t) & 0x7fffffff) - 0x3c000000 + (abs(__double2hiint(nb) - __double2hiint(na)) >> 1), 0) >> 16;   // 16 (log2 tau + 63)
      unsafe = max(unsafe & 0xffff0000, e1 << 16) | max(unsafe & 0xffff, min(e2, 0xffff));
    }
  }
  return r;
}

// reduce 8 per-lane partials so that every lane of 4-lane group grp holds the warp total of element grp
__device__ __forceinline__ double cb_reduce8(double (&pd)[8], int lane) {
  halve<8>(pd, (lane & 16) != 0, 16);
  double h4[4] = {pd[0], pd[1], pd[2], pd[3]};
  halve<4>(h4, (lane & 8) != 0, 8);
  double h2[2] = {h4[0], h4[1]};
  halve<2>(h2, (lane & 4) != 0, 4);
  double d = h2[0];
  d += shfl_xor(d, 2);
  d += shfl_xor(d, 1);
  return d;
}

// (p, q) -> (p - alpha q, q + beta p), stored exchanged: first slot <- new q, second slot <- new p
__device__ __forceinline__ void fast_rot_swap(double& xa, double& xb, double alpha, double beta) {
  const double np = fma(-alpha, xb, xa);
  const double nq = fma(beta, xa, xb);
  xa = nq;
  xb = np;
}

template <bool STEP_B>
__device__ __forceinline__ void cb_step(CbState& st, double2* wcs, int warp, int lane, double tol2, int& rotated, int& unsafe) {
  const int grp = lane >> 2;
  // slot pairs of this step in register terms: A: (2j, 2j+1); B: (2j+1, 2j+2) with slot 16 = the borrowed column
  double pd[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (!STEP_B) pd[j] = fma(st.g1[2 * j], st.g1[2 * j + 1], st.g0[2 * j] * st.g0[2 * j + 1]);
    else if (j < 7) pd[j] = fma(st.g1[2 * j + 1], st.g1[2 * j + 2], st.g0[2 * j + 1] * st.g0[2 * j + 2]);
    else pd[j] = fma(st.g1[15], st.xg1, st.g0[15] * st.xg0);
  }
  const double dhat = cb_reduce8(pd, lane);
  double na, nb, Dp, Dpi, Dq, Dqi;
  bool have = true;
  if (!STEP_B) { na = st.e; nb = st.o; Dp = st.de; Dpi = st.dei; Dq = st.dq; Dqi = st.dqi; }
  else {
    const double e_next = __shfl_down_sync(kFull, st.e, 4);      // even slot of the next group
    const double de_next = __shfl_down_sync(kFull, st.de, 4);
    const double dei_next = __shfl_down_sync(kFull, st.dei, 4);
    na = st.o; Dp = st.dq; Dpi = st.dqi;
    nb = (grp < 7) ? e_next : st.xn;
    Dq = (grp < 7) ? de_next : st.xd;
    Dqi = (grp < 7) ? dei_next : st.xdi;
    have = (grp < 7) || (warp < 3);                                // slot 64 does not exist
  }
  double na2, nb2, Dp2, Dpi2, Dq2, Dqi2;
  const CbRot r = cb_params(dhat, na, nb, Dp, Dpi, Dq, Dqi, have, tol2, na2, nb2, Dp2, Dpi2, Dq2, Dqi2, rotated, unsafe);
  // exchanged storage: first slot of the pair <- rotated q, second slot <- rotated p
  if (!STEP_B) {
    st.e = nb2; st.de = Dq2; st.dei = Dqi2;
    st.o = na2; st.dq = Dp2; st.dqi = Dpi2;
  } else {
    const double n_prev = __shfl_up_sync(kFull, na2, 4);          // group grp-1's second slot is my even slot
    const double d_prev = __shfl_up_sync(kFull, Dp2, 4);
    const double di_prev = __shfl_up_sync(kFull, Dpi2, 4);
    if (have) { st.o = nb2; st.dq = Dq2; st.dqi = Dqi2; }
    if (grp > 0) { st.e = n_prev; st.de = d_prev; st.dei = di_prev; }
    if (grp == 7 && have) { st.xn = na2; st.xd = Dp2; st.xdi = Dpi2; }
  }
  // (alpha, beta) of the 8 pairs to every lane: group leaders publish to a warp-private smem line, uniform 16-byte reads
  __syncwarp();
  if ((lane & 2) == 0) reinterpret_cast<double*>(wcs)[2 * grp + (lane & 1)] = (lane & 1) ? r.beta : r.alpha;   // 2 wavefronts (a 16-byte store from 8 scattered lanes takes 4)
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const double2 ab = wcs[j];
    if (!STEP_B) {
      fast_rot_swap(st.g0[2 * j], st.g0[2 * j + 1], ab.x, ab.y); fast_rot_swap(st.g1[2 * j], st.g1[2 * j + 1], ab.x, ab.y);
      fast_rot_swap(st.v0[2 * j], st.v0[2 * j + 1], ab.x, ab.y); fast_rot_swap(st.v1[2 * j], st.v1[2 * j + 1], ab.x, ab.y);
    } else if (j < 7) {
      fast_rot_swap(st.g0[2 * j + 1], st.g0[2 * j + 2], ab.x, ab.y); fast_rot_swap(st.g1[2 * j + 1], st.g1[2 * j + 2], ab.x, ab.y);
      fast_rot_swap(st.v0[2 * j + 1], st.v0[2 * j + 2], ab.x, ab.y); fast_rot_swap(st.v1[2 * j + 1], st.v1[2 * j + 2], ab.x, ab.y);
    } else if (warp < 3) {  // the last warp has no right neighbour: slot 63 stays where it is
      fast_rot_swap(st.g0[15], st.xg0, ab.x, ab.y); fast_rot_swap(st.g1[15], st.xg1, ab.x, ab.y);
      fast_rot_swap(st.v0[15], st.xv0, ab.x, ab.y); fast_rot_swap(st.v1[15], st.xv1, ab.x, ab.y);
    }
  }
}

constexpr size_t kSvd64CbSmem = sizeof(double) * (2 * 64 * kSvd64LD + 64 + 2 * 4 * 136 + 4 * 16) + sizeof(int) * (64 * 3 + 4);

// PRE: the iteration starts from (G1, V1) = (A V1, V1) with an orthogonal V1 prepared by svd_pre.cu instead of (A, I); A is
// then the G1 array and V1 its accumulated rotation.
template <int MINB, bool PADDED = false, bool PRE = false>
__global__ void __launch_bounds__(128, MINB)
svd64cb_kernel(const double* __restrict__ A, double* __restrict__ U, double* __restrict__ SV, double* __restrict__ V,
               int64_t batch, int* sweeps_out, int* fail_out, unsigned long long* sweep_sum, int rows = 64, int cols = 64,
               const double* __restrict__ V1 = nullptr) {
  constexpr int N = 64, LD = kSvd64LD, XS = 136;       // XS: doubles per exchange record (4*32 column values + norm, D, 1/D)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Gs = reinterpret_cast<double*>(smem_raw);
  double* Vs = Gs + N * LD;
  double* sq = Vs + N * LD;
  double* xch = sq + N;                                 // [2 directions][4 warps][XS]
  double2* wcs_all = reinterpret_cast<double2*>(xch + 2 * 4 * XS);   // [4 warps][8] (alpha, beta) of the current step
  int* perm = reinterpret_cast<int*>(wcs_all + 4 * 8);
  int* zero_flag = perm + N;
  int* done_flag = zero_flag + N;

  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = lane >> 2;
  const double* a_in = A + m * (PADDED ? rows * cols : N * N);
  double2* wcs = wcs_all + 8 * warp;
  double* wsc = reinterpret_cast<double*>(wcs);         // the same 16 doubles, reused for the per-slot scales between sweeps

  CbState st;
  if (!PADDED) {
#pragma unroll
    for (int s = 0; s < 16; s += 2) {
      const double2 t0 = ldg2_stream(a_in + lane * N + 16 * warp + s);
      const double2 t1 = ldg2_stream(a_in + (lane + 32) * N + 16 * warp + s);
      st.g0[s] = t0.x; st.g0[s + 1] = t0.y;
      st.g1[s] = t1.x; st.g1[s + 1] = t1.y;
    }
  } else if (rows >= cols) {   // rows x cols zero-padded into the 64 x 64 tile: zero rows and zero columns never rotate
#pragma unroll
    for (int s = 0; s < 16; s++) {
      const int col = 16 * warp + s;
      st.g0[s] = (lane < rows && col < cols) ? ldg1_stream(a_in + lane * cols + col) : 0.0;
      st.g1[s] = (lane + 32 < rows && col < cols) ? ldg1_stream(a_in + (lane + 32) * cols + col) : 0.0;
    }
  } else {   // wide: the transpose is factorised (more columns than rows could never become mutually orthogonal)
#pragma unroll
    for (int s = 0; s < 16; s++) {
      const int col = 16 * warp + s;   // column of A^T = row of A
      st.g0[s] = (lane < cols && col < rows) ? ldg1_stream(a_in + col * cols + lane) : 0.0;
      st.g1[s] = (lane + 32 < cols && col < rows) ? ldg1_stream(a_in + col * cols + lane + 32) : 0.0;
    }
  }
  if (PRE) {
    const double* v_in = V1 + m * (N * N);
#pragma unroll
    for (int s = 0; s < 16; s += 2) {
      const double2 t0 = ldg2_stream(v_in + lane * N + 16 * warp + s);
      const double2 t1 = ldg2_stream(v_in + (lane + 32) * N + 16 * warp + s);
      st.v0[s] = t0.x; st.v0[s + 1] = t0.y;
      st.v1[s] = t1.x; st.v1[s + 1] = t1.y;
    }
  } else {
#pragma unroll
    for (int s = 0; s < 16; s++) {
      st.v0[s] = (16 * warp + s == lane) ? 1.0 : 0.0;
      st.v1[s] = (16 * warp + s == lane + 32) ? 1.0 : 0.0;
    }
  }
  st.xg0 = st.xg1 = st.xv0 = st.xv1 = st.xn = 0.0;
  st.xd = st.xdi = 1.0;
  // scale guard (see pow2_prescale): squared column norms must neither overflow nor flush to zero
  double amax = 0.0;
#pragma unroll
  for (int s = 0; s < 16; s++) amax = fmax(amax, fmax(fabs(st.g0[s]), fabs(st.g1[s])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, shfl_xor(amax, o));
  if (lane == 0) sq[warp] = amax;
  __syncthreads();
  amax = fmax(fmax(sq[0], sq[1]), fmax(sq[2], sq[3]));
  const double pre = pow2_prescale_unit(amax);
  if (pre != 1.0) {
#pragma unroll
    for (int s = 0; s < 16; s++) { st.g0[s] *= pre; st.g1[s] *= pre; }
  }

  const double tol2 = (N * kEps) * (N * kEps);
  int sweeps = 0;
  bool converged = false;
  while (sweeps < kMaxSweeps && !converged) {
    sweeps++;
    int rotated = 0, unsafe = 0;
    st.de = st.dei = st.dq = st.dqi = 1.0;
    {  // exact slot norms: 16 values -> 2 per 4-lane group
      double n2[16];
#pragma unroll
      for (int s = 0; s < 16; s++) n2[s] = fma(st.g1[s], st.g1[s], st.g0[s] * st.g0[s]);
      halve<16>(n2, (lane & 16) != 0, 16);
      double h8[8];
#pragma unroll
      for (int k = 0; k < 8; k++) h8[k] = n2[k];
      halve<8>(h8, (lane & 8) != 0, 8);
      double h4[4] = {h8[0], h8[1], h8[2], h8[3]};
      halve<4>(h4, (lane & 4) != 0, 4);
      double e = h4[0], o = h4[1];
      e += shfl_xor(e, 2); o += shfl_xor(o, 2);
      e += shfl_xor(e, 1); o += shfl_xor(o, 1);
      st.e = e; st.o = o;
    }
#pragma unroll 1
    for (int sp2 = 0; sp2 < N / 2; sp2++) {
      cb_step<false>(st, wcs, warp, lane, tol2, rotated, unsafe);
      // hand my first column (slot 16w) to the left neighbour for the B step
      double* out = xch + warp * XS;
      if (warp > 0) {
        out[lane] = st.g0[0]; out[32 + lane] = st.g1[0]; out[64 + lane] = st.v0[0]; out[96 + lane] = st.v1[0];
        if (lane == 0) { out[128] = st.e; out[129] = st.de; out[130] = st.dei; }   // group 0's even slot
      }
      __syncthreads();
      if (warp < 3) {
        const double* in = xch + (warp + 1) * XS;
        st.xg0 = in[lane]; st.xg1 = in[32 + lane]; st.xv0 = in[64 + lane]; st.xv1 = in[96 + lane];
        st.xn = in[128]; st.xd = in[129]; st.xdi = in[130];
      }
      cb_step<true>(st, wcs, warp, lane, tol2, rotated, unsafe);
      // give the borrowed (rotated) column back
      double* out2 = xch + 4 * XS + (warp + 1) * XS;
      if (warp < 3) {
        out2[lane] = st.xg0; out2[32 + lane] = st.xg1; out2[64 + lane] = st.xv0; out2[96 + lane] = st.xv1;
        if (lane == 28) { out2[128] = st.xn; out2[129] = st.xd; out2[130] = st.xdi; }   // a lane of group 7
      }
      __syncthreads();
      if (warp > 0) {
        const double* in2 = xch + 4 * XS + warp * XS;
        st.g0[0] = in2[lane]; st.g1[0] = in2[32 + lane]; st.v0[0] = in2[64 + lane]; st.v1[0] = in2[96 + lane];
        if (grp == 0) { st.e = in2[128]; st.de = in2[129]; st.dei = in2[130]; }
      }
    }
    // fold the scales back into the columns: every lane needs D of all 16 slots of its warp
    __syncwarp();
    if ((lane & 3) == 0) { wsc[2 * grp] = st.de; wsc[2 * grp + 1] = st.dq; }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 16; s++) {
      const double D = wsc[s];
      st.g0[s] *= D; st.g1[s] *= D; st.v0[s] *= D; st.v1[s] *= D;
    }
    __syncwarp();
    {
      // block maxima of the two terms (cb_params) — the maxima of different warps must be combined: one warp's largest |t| can
      // meet another warp's largest cosine.  Converged when nothing was rotated, or when
      // log2 max|cos| + log2 max|t| < log2(64 eps / 126) - margin = -55:  (e1/16 - 92)/2 + (e2/16 - 63) < -55, i.e. e1/2 + e2 < 864
      const int m1 = __reduce_max_sync(kFull, (unsigned)unsafe >> 16), m2 = __reduce_max_sync(kFull, unsafe & 0xffff);
      int* red = reinterpret_cast<int*>(sq);   // sq is idle between the prologue and the epilogue
      if (lane == 0) { red[2 * warp] = m1; red[2 * warp + 1] = m2; }
      const bool any_rot = __syncthreads_or(rotated) != 0;   // also orders the writes above before the reads below
      const int b1 = max(max(red[0], red[2]), max(red[4], red[6])), b2 = max(max(red[1], red[3]), max(red[5], red[7]));
      converged = !any_rot || (b1 >> 1) + b2 < 864;
      __syncthreads();   // red is rewritten at the end of the next sweep only, but keep the two uses apart
    }
  }
  if (tid == 0) {
    if (sweeps_out) atomicMax(sweeps_out, sweeps);
    if (sweep_sum) atomicAdd(sweep_sum, (unsigned long long)sweeps);
    if (!converged && fail_out) atomicExch(fail_out, 1);
  }
#pragma unroll
  for (int s = 0; s < 16; s++) {
    Gs[(16 * warp + s) * LD + lane] = st.g0[s];
    Gs[(16 * warp + s) * LD + lane + 32] = st.g1[s];
    Vs[(16 * warp + s) * LD + lane] = st.v0[s];
    Vs[(16 * warp + s) * LD + lane + 32] = st.v1[s];
  }
  __syncthreads();
  svd64_epilogue<128, PADDED>(Gs, Vs, sq, perm, zero_flag, done_flag, U, SV, V, m, 1.0 / pre, rows, cols);
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

// T threads per matrix; a pair of columns is rotated by a group of LPP lanes (32, or 16 / 8 for short vectors, so that
// small matrices neither idle most lanes of a warp nor occupy eight warps each).
template <int T, int LPP>
__global__ void __launch_bounds__(T)
svd_generic_kernel(const double* __restrict__ A, double* __restrict__ U, double* __restrict__ SV, double* __restrict__ V,
                   int64_t batch, int rows, int cols, int* sweeps_out, int* fail_out, double* __restrict__ work,
                   unsigned long long* sweep_sum) {
  extern __shared__ __align__(16) double svd_gen_smem[];
  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const bool wide = rows < cols;
  const int mm = wide ? cols : rows, n = wide ? rows : cols;
  // work == nullptr: the per-matrix scratch (G^T, V^T, sigma, permutation, flags) lives in shared memory
  double* Gt = work ? work + m * svd_gen_scratch_doubles(rows, cols) : svd_gen_smem;
  double* Vt = Gt + (size_t)n * mm;
  double* sig = Vt + (size_t)n * n;
  int* perm = reinterpret_cast<int*>(sig + n);
  int* zero_flag = perm + n;
  int* done_flag = zero_flag + n;
  const double* a_in = A + m * (int64_t)rows * cols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = T / 32;

  // G = A (tall: Gt[j][i] = A[i][j]) or A^T (wide: Gt[j][i] = A[j][i])
  __shared__ double red_max[NW];
  double amax = 0.0;
  for (int64_t e = tid; e < (int64_t)rows * cols; e += T) amax = fmax(amax, fabs(a_in[e]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(kFull, amax, o));
  if (lane == 0) red_max[warp] = amax;
  __syncthreads();
  amax = 0.0;
  for (int w = 0; w < NW; w++) amax = fmax(amax, red_max[w]);
  const double pre = pow2_prescale_unit(amax), post = 1.0 / pre;   // scale guard, exact power of two
  for (int64_t e = tid; e < (int64_t)rows * cols; e += T) {
    const double x = a_in[e] * pre;
    if (wide) Gt[e] = x;
    else { const int i = (int)(e / cols), j = (int)(e % cols); Gt[(size_t)j * mm + i] = x; }
  }
  for (int64_t e = tid; e < (int64_t)n * n; e += T) Vt[e] = (e / n == e % n) ? 1.0 : 0.0;
  __syncthreads();

  const double tol2 = ((double)mm * kEps) * ((double)mm * kEps);
  const int n2 = n + (n & 1);
  int sweeps = 0;
  bool converged = (n < 2);
  for (; sweeps < kMaxSweeps && !converged;) {
    int rotated = 0;
    sweeps++;
    for (int s = 0; s < n2 - 1; s++) {
      for (int P = tid / LPP; P < n2 / 2; P += T / LPP) {
        // the lanes of a group stay together; groups of one warp may diverge (different trip counts, bye pairs, rotate or not)
        const unsigned gmask = LPP == 32 ? kFull : (((1u << LPP) - 1u) << ((lane / LPP) * LPP));
        const int sub = tid % LPP;
        int p, q;
        rr_pair(n2, s, P, p, q);
        if (q >= n) continue;  // bye (odd n)
        double* gp = Gt + (size_t)p * mm;
        double* gq = Gt + (size_t)q * mm;
        double a = 0.0, b = 0.0, d = 0.0;
        for (int i = sub; i < mm; i += LPP) {
          const double x = gp[i], y = gq[i];
          a = fma(x, x, a); b = fma(y, y, b); d = fma(x, y, d);
        }
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) {
          a += __shfl_xor_sync(gmask, a, o); b += __shfl_xor_sync(gmask, b, o); d += __shfl_xor_sync(gmask, d, o);
        }
        if (d * d > tol2 * a * b && a > kNormMin && b > kNormMin) {
          rotated = 1;
          const Rot r = make_rotation(a, b, d);
          for (int i = sub; i < mm; i += LPP) {
            const double x = gp[i], y = gq[i];
            gp[i] = fma(r.c, x, -r.s * y);
            gq[i] = fma(r.s, x, r.c * y);
          }
          double* vp = Vt + (size_t)p * n;
          double* vq = Vt + (size_t)q * n;
          for (int i = sub; i < n; i += LPP) {
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
    if (sweep_sum) atomicAdd(sweep_sum, (unsigned long long)sweeps);
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
  for (int j = tid; j < n; j += T) {
    const double sj = sig[j];
    int rank = 0;
    double smax = 0.0;
    for (int k = 0; k < n; k++) {
      const double sk = sig[k];
      rank += sv_before(sk, k, sj, j);
      smax = fmax(smax, sk);
    }
    perm[rank] = j;
    zero_flag[j] = (sj < smax * kZeroRel) || (sj < DBL_MIN)   /* a NaN norm is not a zero: NaN in, NaN out */;
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
  for (int64_t e = tid; e < (int64_t)rows * n; e += T) {
    const int i = (int)(e / n), l = (int)(e % n);
    u_out[e] = usrc[(size_t)perm[l] * uld + i];
  }
  for (int64_t e = tid; e < (int64_t)n * cols; e += T) {
    const int l = (int)(e / cols), j = (int)(e % cols);
    v_out[e] = vsrc[(size_t)perm[l] * vld + j];
  }
  for (int l = tid; l < n; l += T) SV[m * n + l] = zero_flag[perm[l]] ? 0.0 : sig[perm[l]] * post;
}

// ------------------------------------------------------------------------------------------------
// Tiny matrices (rows, cols <= 8): one lane per matrix, 32 matrices per warp, staged with 8-byte cp.async into odd-stride
// shared-memory slots.  Each lane runs a cyclic one-sided Jacobi on its own slot (the transpose when the matrix is wide),
// with the threshold, the ordering and sign conventions, the zero-column completion and the exactness on diagonal input
// of the kernels above; the results leave as coalesced stores that apply the descending order on the fly.
// ------------------------------------------------------------------------------------------------
constexpr int kSvdTinyWarps = 2;

__global__ void __launch_bounds__(kSvdTinyWarps * 32)
svd_tiny_kernel(const double* __restrict__ A, double* __restrict__ U, double* __restrict__ SV, double* __restrict__ V,
                int64_t batch, int rows, int cols, int* sweeps_out, int* fail_out, unsigned long long* sweep_sum) {
  extern __shared__ __align__(16) double svt_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool wide = rows < cols;
  const int mm = wide ? cols : rows, n = wide ? rows : cols;   // G is mm x n (n columns of length mm), V is n x n
  const int na = rows * cols, sg = na | 1, sv_ = (n * n) | 1, sl = sg + sv_ + 16;
  double* slots = svt_smem + (size_t)warp * 32 * sl;
  const int64_t m0 = ((int64_t)blockIdx.x * kSvdTinyWarps + warp) * 32;
  if (m0 >= batch) return;  // warp-uniform
  const int nmat = (int)min((int64_t)32, batch - m0);
  {
    const double* src = A + m0 * na;
    const uint32_t base_s = (uint32_t)__cvta_generic_to_shared(slots);
    const float inv = 1.0f / (float)na;
    for (int g = lane; g < nmat * na; g += 32) {
      const int q = (int)(((float)g + 0.5f) * inv), e = g - q * na;   // exact: g < 2048
      // tall / square: G = A (row-major mm x n); wide: G = A^T, i.e. G[i][j] = A[j][i]
      const int pos = wide ? (e % cols) * n + e / cols : e;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(base_s + (uint32_t)(q * sl + pos) * 8u), "l"(src + g) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
  }
  int my_sweeps = 0, my_fail = 0;
  if (lane < nmat) {
    double* G = slots + lane * sl;
    double* Vv = G + sg;
    double* sig = Vv + sv_;
    double* permd = sig + 8;
    double amax = 0.0;
    for (int e = 0; e < na; e++) amax = fmax(amax, fabs(G[e]));
    const double pre = pow2_prescale_unit(amax);
    if (pre != 1.0)
      for (int e = 0; e < na; e++) G[e] *= pre;
    for (int e = 0; e < n * n; e++) Vv[e] = 0.0;
    for (int j = 0; j < n; j++) Vv[j * n + j] = 1.0;
    const double tol2 = ((double)mm * kEps) * ((double)mm * kEps);
    bool converged = false;
    while (my_sweeps < kMaxSweeps && !converged) {
      my_sweeps++;
      converged = true;
      for (int p = 0; p < n - 1; p++)
        for (int q = p + 1; q < n; q++) {
          double a = 0.0, b = 0.0, d = 0.0;
          for (int i = 0; i < mm; i++) {
            const double x = G[i * n + p], y = G[i * n + q];
            a = fma(x, x, a); b = fma(y, y, b); d = fma(x, y, d);
          }
          if (!(d * d > tol2 * a * b && a > kNormMin && b > kNormMin)) continue;
          converged = false;
          const double zeta = (b - a) / (2.0 * d);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
          const double c = 1.0 / sqrt(fma(t, t, 1.0)), sn = c * t;
          for (int i = 0; i < mm; i++) {
            const double x = G[i * n + p], y = G[i * n + q];
            G[i * n + p] = c * x - sn * y;
            G[i * n + q] = sn * x + c * y;
          }
          for (int i = 0; i < n; i++) {
            const double x = Vv[i * n + p], y = Vv[i * n + q];
            Vv[i * n + p] = c * x - sn * y;
            Vv[i * n + q] = sn * x + c * y;
          }
        }
    }
    if (!converged) my_fail = 1;
    // singular values, stable descending order, zero detection
    double smax = 0.0;
    for (int j = 0; j < n; j++) {
      double a = 0.0;
      for (int i = 0; i < mm; i++) a = fma(G[i * n + j], G[i * n + j], a);
      sig[j] = sqrt(a);
      smax = fmax(smax, sig[j]);
    }
    unsigned zero_mask = 0, done_mask = 0;
    for (int j = 0; j < n; j++) {
      int rank = 0;
      for (int k = 0; k < n; k++) rank += sv_before(sig[k], k, sig[j], j);
      permd[rank] = (double)j;
      if (sig[j] < smax * kZeroRel || sig[j] < DBL_MIN) zero_mask |= 1u << j;
    }
    for (int j = 0; j < n; j++)
      if (!((zero_mask >> j) & 1)) {
        const double sj = sig[j];
        for (int i = 0; i < mm; i++) G[i * n + j] /= sj;
      }
    // zero columns -> completed to an orthonormal set (as complete_basis_warp does)
    for (int z = 0; z < n; z++) {
      if (!((zero_mask >> z) & 1)) continue;
      double best = DBL_MAX;
      int best_i = 0;
      for (int i = 0; i < mm; i++) {
        double rn = 0.0;
        for (int k = 0; k < n; k++)
          if (!((zero_mask >> k) & 1) || ((done_mask >> k) & 1)) rn = fma(G[i * n + k], G[i * n + k], rn);
        if (rn < best) { best = rn; best_i = i; }
      }
      for (int i = 0; i < mm; i++) G[i * n + z] = (i == best_i) ? 1.0 : 0.0;
      for (int pass = 0; pass < 2; pass++)
        for (int k = 0; k < n; k++) {
          if (k == z || (((zero_mask >> k) & 1) && !((done_mask >> k) & 1))) continue;
          double dot = 0.0;
          for (int i = 0; i < mm; i++) dot = fma(G[i * n + k], G[i * n + z], dot);
          if (dot != 0.0)
            for (int i = 0; i < mm; i++) G[i * n + z] = fma(-dot, G[i * n + k], G[i * n + z]);
        }
      double nn = 0.0;
      for (int i = 0; i < mm; i++) nn = fma(G[i * n + z], G[i * n + z], nn);
      if (nn != 1.0) {
        const double inv = 1.0 / sqrt(nn);
        for (int i = 0; i < mm; i++) G[i * n + z] *= inv;
      }
      done_mask |= 1u << z;
    }
    const double post = 1.0 / pre;
    for (int j = 0; j < n; j++) sig[j] = ((zero_mask >> j) & 1) ? 0.0 : sig[j] * post;
  }
  __syncwarp();
  {  // coalesced stores; n = L.  tall / square: U = G[:, perm], V_out[l][j] = V[j][perm[l]];  wide: U = V[:, perm], V_out[l][j] = G[j][perm[l]]
    const int nu = rows * n, nv = n * cols;
    double* ud = U + m0 * nu;
    const float invu = 1.0f / (float)nu;
    for (int g = lane; g < nmat * nu; g += 32) {
      const int q = (int)(((float)g + 0.5f) * invu), e = g - q * nu, i = e / n, l = e - i * n;
      const double* S0 = slots + q * sl;
      const int pl = (int)S0[sg + sv_ + 8 + l];
      ud[g] = wide ? S0[sg + i * n + pl] : S0[i * n + pl];
    }
    double* vd = V + m0 * nv;
    const float invv = 1.0f / (float)nv;
    for (int g = lane; g < nmat * nv; g += 32) {
      const int q = (int)(((float)g + 0.5f) * invv), e = g - q * nv, l = e / cols, j = e - l * cols;
      const double* S0 = slots + q * sl;
      const int pl = (int)S0[sg + sv_ + 8 + l];
      vd[g] = wide ? S0[j * n + pl] : S0[sg + j * n + pl];
    }
    double* sd = SV + m0 * n;
    for (int g = lane; g < nmat * n; g += 32) {
      const int q = g / n, l = g - q * n;
      const double* S0 = slots + q * sl;
      sd[g] = S0[sg + sv_ + (int)S0[sg + sv_ + 8 + l]];
    }
  }
  const int wmax = __reduce_max_sync(kFull, my_sweeps);
  const int wsum = __reduce_add_sync(kFull, my_sweeps);
  const int wfail = __reduce_max_sync(kFull, my_fail);
  if (lane == 0) {
    if (sweeps_out) atomicMax(sweeps_out, wmax);
    if (sweep_sum) atomicAdd(sweep_sum, (unsigned long long)wsum);
    if (wfail && fail_out) atomicExch(fail_out, 1);
  }
}

// Diagnostic: device counter that every following SVD launch on `device` adds its per-matrix sweep counts to (nullptr: off).
static unsigned long long* g_sweep_sum[64] = {nullptr};
static unsigned long long* g_pre_sweep_sum[64] = {nullptr};
void set_svd_sweep_counter(int device, unsigned long long* counter) {
  if (device >= 0 && device < 64) g_sweep_sum[device] = counter;
}
void set_svd_pre_sweep_counter(int device, unsigned long long* counter) {
  if (device >= 0 && device < 64) g_pre_sweep_sum[device] = counter;
}

constexpr size_t kSvdGenSmemLimit = 200 * 1024;

static bool svd_pre_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* ev = getenv("ND4B_SVD_PRE");
    on = ev ? atoi(ev) : 1;   // 0: the plain FP64 iteration from (A, I), for A/B timing
  }
  return on != 0;
}

size_t svd_workspace_bytes(int64_t batch, int rows, int cols) {
  if (rows == 64 && cols == 64 && svd_pre_enabled()) return svd64_pre_workspace_bytes(batch);   // V0, G1, V1 of the preconditioner
  if (rows <= 64 && cols <= 64) return 0;   // register kernel (padded) or scratch in shared memory
  if (sizeof(double) * svd_gen_scratch_doubles(rows, cols) <= kSvdGenSmemLimit) return 0;  // scratch in shared memory
  return sizeof(double) * (size_t)batch * svd_gen_scratch_doubles(rows, cols);
}

cudaError_t launch_svd_jac1(cudaStream_t s, const double* A, double* U, double* sv, double* V,
                            int64_t batch, int rows, int cols, int* sweeps, int* fail,
                            double* work, size_t work_bytes) {
  if (batch <= 0) return cudaSuccess;
  if (batch > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  int dev = 0;
  cudaGetDevice(&dev);
  unsigned long long* ssum = (dev >= 0 && dev < 64) ? g_sweep_sum[dev] : nullptr;
  // rows, cols <= 64 with many columns: zero-padded through the tuned 64 x 64 register kernel — 1.1-1.3 us per matrix whatever
  // the shape (ND4B_SVD_PADDED=0 switches the route off)
  static int padded = -1;
  if (padded < 0) {
    const char* ev = getenv("ND4B_SVD_PADDED");
    padded = ev ? atoi(ev) : 1;
  }
  if (padded && rows <= 8 && cols <= 8 && batch >= 64) {
    const int nn = rows < cols ? rows : cols;
    const size_t smem = sizeof(double) * kSvdTinyWarps * 32 * (size_t)((rows * cols | 1) + (nn * nn | 1) + 16);
    static bool tattr_set[64] = {false};
    if (dev >= 0 && dev < 64 && !tattr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(svd_tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (e != cudaSuccess) return e;
      tattr_set[dev] = true;
    }
    const int64_t grid = (batch + kSvdTinyWarps * 32 - 1) / (kSvdTinyWarps * 32);
    svd_tiny_kernel<<<(unsigned)grid, kSvdTinyWarps * 32, smem, s>>>(A, U, sv, V, batch, rows, cols, sweeps, fail, ssum);
    return cudaGetLastError();
  }
  // measured crossover against the generic kernel (8192 matrices): 32x32 generic 0.84 us vs padded 1.18 us per matrix,
  // 36x36 2.35 vs 1.22, 40x30 1.43 vs 1.10, 48x24 0.88 vs 1.01, 64x16 0.33 vs 0.88
  const int nmin_ = rows < cols ? rows : cols, mmax_ = rows > cols ? rows : cols;
  if (padded && rows <= 64 && cols <= 64 && !(rows == 64 && cols == 64) && (nmin_ > 32 || (mmax_ > 32 && nmin_ >= 28))) {
    static bool pattr_set[64] = {false};
    if (dev >= 0 && dev < 64 && !pattr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(svd64cb_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSvd64CbSmem);
      if (e != cudaSuccess) return e;
      pattr_set[dev] = true;
    }
    svd64cb_kernel<2, true><<<(unsigned)batch, 128, kSvd64CbSmem, s>>>(A, U, sv, V, batch, sweeps, fail, ssum, rows, cols);
    return cudaGetLastError();
  }
  if (rows == 64 && cols == 64) {
    static bool attr_set[64] = {false};
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(svd64cb_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSvd64CbSmem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(svd64_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSvd64Smem);
      if (e != cudaSuccess) return e;
      attr_set[dev] = true;
    }
    static int variant = -1;
    if (variant < 0) {
      const char* ev = getenv("ND4B_SVD_VARIANT");
      variant = ev ? atoi(ev) : 0;  // 1 = the shared-memory baseline kernel (kept for A/B profiling)
    }
    if (variant == 1) svd64_smem_kernel<<<(unsigned)batch, kSvd64Threads, kSvd64Smem, s>>>(A, U, sv, V, batch, sweeps, fail, ssum);
    else if (svd_pre_enabled() && work != nullptr && work_bytes >= svd64_pre_workspace_bytes(batch) &&
             ((reinterpret_cast<uintptr_t>(work) | reinterpret_cast<uintptr_t>(A)) & 15) == 0) {   // 16-byte vector accesses on V0, G1, V1
      // FP32 Jacobi -> orthogonalised V1, G1 = A V1 -> FP64 Jacobi from (G1, V1): svd_pre.cu
      static bool pre_attr_set[64] = {false};
      if (dev >= 0 && dev < 64 && !pre_attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(svd64cb_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSvd64CbSmem);
        if (e != cudaSuccess) return e;
        pre_attr_set[dev] = true;
      }
      const double *g1 = nullptr, *v1 = nullptr;
      cudaError_t e = launch_svd64_pre(s, A, batch, work, (dev >= 0 && dev < 64) ? g_pre_sweep_sum[dev] : nullptr, &g1, &v1);
      if (e != cudaSuccess) return e;
      svd64cb_kernel<2, false, true><<<(unsigned)batch, 128, kSvd64CbSmem, s>>>(g1, U, sv, V, batch, sweeps, fail, ssum, 64, 64, v1);
    } else svd64cb_kernel<2><<<(unsigned)batch, 128, kSvd64CbSmem, s>>>(A, U, sv, V, batch, sweeps, fail, ssum);
    return cudaGetLastError();
  }
  const size_t per_matrix = sizeof(double) * svd_gen_scratch_doubles(rows, cols);
  if (per_matrix <= kSvdGenSmemLimit) {  // same kernel, scratch in shared memory (L1 instead of L2 latency on every access)
    const int mmax = rows > cols ? rows : cols, nmin = rows < cols ? rows : cols;
    const int pairs = (nmin + 1) / 2;
    // short vectors: 8 or 16 lanes per column pair and only as many warps as the pairs of a step need
    auto go = [&](auto kernel, int threads) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSvdGenSmemLimit);
      if (e != cudaSuccess) return e;
      kernel<<<(unsigned)batch, threads, per_matrix, s>>>(A, U, sv, V, batch, rows, cols, sweeps, fail, nullptr, ssum);
      return cudaGetLastError();
    };
    if (mmax <= 16 && pairs <= 8) return go(svd_generic_kernel<64, 8>, 64);
    if (mmax <= 16) return go(svd_generic_kernel<128, 8>, 128);
    if (mmax <= 32 && pairs <= 8) return go(svd_generic_kernel<128, 16>, 128);
    if (mmax <= 32) return go(svd_generic_kernel<256, 16>, 256);
    return go(svd_generic_kernel<256, 32>, 256);
  }
  const size_t need = per_matrix * (size_t)batch;
  if (work == nullptr || work_bytes < need) return cudaErrorInvalidValue;
  svd_generic_kernel<256, 32><<<(unsigned)batch, 256, 0, s>>>(A, U, sv, V, batch, rows, cols, sweeps, fail, work, ssum);
  return cudaGetLastError();
}

}  // namespace nd4b
