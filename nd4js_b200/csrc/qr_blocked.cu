// qr_blocked.cu — batched 64x32 Householder QR with block reflectors on the FP64 tensor pipe (DMMA.8x8x4).
// Replaces the Givens QR of nd4js src/la/qr.js:80-145 for the C4 shape; same results contract as qr.cu
// (Q[64,32], R[32,32] exactly upper triangular, diag(R) >= 0).
//
// One warp per matrix, the matrix in registers in the accumulator layout of the TRANSPOSE:
//   lane (g,t) = (lane>>2, lane&3) owns columns 8i+g (i<4) and rows 8j+2t+e (j<8, e<2):  a[i][j][e].
// With that layout every product of the blocked algorithm is a chain of DMMA.8x8x4 whose fragments are the registers
// themselves (an accumulator tile X[g][2t+e] is the A fragment pair of a following product, k <-> 2t+e):
//   panel p (columns 8p..8p+7, one column per quad):  unblocked Householder, 8 steps; a step broadcasts the pivot
//       column through shared memory, forms all dot products with it (the pivot quad's is the squared norm) and
//       applies the rank-1 update — reflector scalars need one rsqrt and one reciprocal.
//   T of the compact WY form (H_1..H_8 = I - V T V^T):  G = V^T V by DMMA (both fragments are the same register),
//       T^-1 = striu(G) + diag(1/tau)  =>  T = (I+N)^-1 diag(tau), N = diag(tau) striu(G) nilpotent, and
//       (I+N)^-1 = (I-N)(I+N^2)(I+N^4) exactly: six 8x8 products, in both orientations where needed, so that no transpose is needed.
//   trailing columns:  W^T = C^T V,  Z^T = W^T T,  C^T -= Z^T V^T  (the last product needs V with rows on the g index:
//       the clean panel is written once to shared memory, row-major, and read back with 16-byte loads).
//   Q = H_1 .. H_32 [I;0] by blocks from the last panel: trailing columns as above with T^T, the panel's own columns
//       as E_p - V_p (T_p L_p^T) (L_p = unit lower 8x8 head of V_p) — no serial reflector loop in the Q phase.
// Work per matrix: 40 960 B of HBM traffic, 218 453 flop by the Householder convention.
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

// Phase profiling hook for tools/micro/qr_phase_prof.cu (which includes this file with QB_PROFILE defined).
#ifdef QB_PROFILE
__device__ unsigned long long qb_prof[16];
#define QB_MARK(i)                                                                              \
  do {                                                                                          \
    const long long n_ = clock64();                                                             \
    if ((threadIdx.x & 31) == 0) atomicAdd(&qb_prof[i], (unsigned long long)(n_ - qb_tm));      \
    qb_tm = n_;                                                                                 \
  } while (0)
#else
#define QB_MARK(i) do { } while (0)
#endif

namespace nd4b {

namespace {

// Per-warp shared memory: the clean V panels (panel p: rows 8p..63, 8 columns, row-major, no padding), the four T's
// (accumulator layout, 2 doubles per lane) and the pivot-column buffers: 16.6 KiB, so that 12 warps fit one SM.
__host__ __device__ constexpr int v_off(int p) { return 8 * (64 * p - 4 * p * (p - 1)); }  // 0, 512, 960, 1344
constexpr int kVDoubles = 1664;
constexpr int kXsOff = kVDoubles + 4 * 64;       // two pivot-column buffers (+ 64 spare)
constexpr int kMbarOff = kXsOff + 3 * 64;        // mbarrier of the bulk load
constexpr int kQbWarpDoubles = kMbarOff + 16;    // 2128 doubles = 17 024 B (a multiple of 128 B)
// The first 2048 doubles double as the landing zone of the bulk load of A (TMA, 16 KiB) and as the staging tile of the
// bulk store of Q: 8-byte per-lane global accesses in the fragment layout cost 4 L1 tag look-ups per instruction.

// Tensor-map variant (IO == 2): A arrives as two TMA tile loads (UTMALDG, box 16 columns x 64 rows, SWIZZLE_128B) and Q leaves
// as two tile stores (UTMASTG).  A row of a box is one 128-byte line whose 16-byte chunks are xor-ed with (row & 7), which is
// exactly what makes the transposing fragment accesses conflict-free: the sixteen lanes (g < 4, t) of a half-warp touch rows
// 2t+e — chunk (g >> 1) ^ (2t + e) takes eight different values — so every 8-byte access is 2 wavefronts instead of 8
// (profiles/r02_ncu_c4_blocked.txt: 1 440 of the 3 150 shared-memory wavefronts per matrix were these loads and stores).
// The swizzle needs 1 KiB-aligned boxes: the per-warp region is rounded up to 17 KiB.
constexpr int kQbWarpDoublesT = 2176;            // 17 408 B
__device__ __forceinline__ int tm_idx(int i, int j, int g, int t, int e) {   // element (row 8j+2t+e, column 8i+g) in the swizzled landing zone
  return (i >> 1) * 1024 + (8 * j + 2 * t + e) * 16 + ((((4 * (i & 1)) + (g >> 1)) ^ (2 * t + e)) << 1) + (g & 1);
}
// V panels in shared memory: row-major rows of 8, the column pairs of a row xor-ed with ((row >> 1) & 3): the 16-byte row reads
// stay conflict-free and the transposing 8-byte stores (lane (g,t) -> row 2t+e, column g) drop from 8 wavefronts to 4.
__device__ __forceinline__ int v_sw(int row) { return ((row >> 1) & 3) << 1; }

struct Acc { double x, y; };  // 8x8 tile in accumulator layout: x = M[g][2t], y = M[g][2t+1]

// X * Y given X and Y^T in accumulator layout (A fragments = X, B fragments B[k<->2t+e][n=g] = Y^T[g][2t+e]).
__device__ __forceinline__ Acc mm8(const Acc& X, const Acc& YT) {
  Acc o{0.0, 0.0};
  dmma884(o.x, o.y, X.x, YT.x);
  dmma884(o.x, o.y, X.y, YT.y);
  return o;
}

// Approximate reciprocal square root / reciprocal (MUFU.RSQ64H / MUFU.RCP64H, about 20 good bits); refined below.
__device__ __forceinline__ double rsqrt_approx(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rcp_approx(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}

// 16-byte shared-memory load that the compiler may not merge with an earlier load of the same address (the pivot column
// is read twice per step instead of being held in 32 registers across the scalar chain).
__device__ __forceinline__ double2 lds2_again(const double* p) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// One lane: bulk copy (TMA, UBLKCP) of `bytes` from global to shared memory, completion on an mbarrier it initialises.
__device__ __forceinline__ void bulk_load_start(void* dst, const void* src, unsigned bytes, void* mbar) {
  const unsigned mb = smem_u32(mbar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(mb) : "memory");
}
__device__ __forceinline__ void bulk_load_wait(void* mbar) {
  const unsigned mb = smem_u32(mbar);
  asm volatile("{\n.reg .pred p;\nQB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra QB_WAIT;\n}" ::"r"(mb) : "memory");
}
// One lane: bulk copy shared -> global; returns when the shared source has been read.
__device__ __forceinline__ void bulk_store(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// One lane: TMA tile load of the box at (x = column, y = row) of a 2-D tensor map into shared memory (UTMALDG), completion
// on an mbarrier that already expects the bytes; tile store shared -> global (UTMASTG).
__device__ __forceinline__ void tmap_load_2d(void* dst, const CUtensorMap* tm, int x, int y, void* mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(x), "r"(y), "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void tmap_store_2d(const CUtensorMap* tm, int x, int y, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<unsigned long long>(tm)), "r"(x), "r"(y), "r"(smem_u32(src)) : "memory");
}

// Predicated 16-byte shared store: lanes with pred == false issue nothing (no divergence, no wavefronts).
__device__ __forceinline__ void sts2_if(bool pred, double* p, double x, double y) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q st.shared.v2.f64 [%1], {%2,%3};\n}"
               ::"r"((unsigned)pred), "r"(smem_u32(p)), "d"(x), "d"(y) : "memory");
}

// x with its sign bit xor-ed by mask (0 or 0x80000000): sign flips without the FP64 pipe.
__device__ __forceinline__ double flip(double x, unsigned mask) {
  return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

// ---- panel P: unblocked Householder on columns 8P..8P+7 (my column: 8P+g), rows >= 8P ----
// xs: two 64-double buffers; the pivot column (rows above the pivot row zeroed) is published there by its quad and
// read back by every lane with broadcast 16-byte loads (a shuffle broadcast of 16 doubles costs 32 SHFL per step).
// The step is one dependent chain (dots -> |x|^2 -> reflector scalars -> update), so the scalar part is written for
// depth: branch-free, 1/sqrt and 1/v0 refined from overlapping 20-bit estimates.
template <int P, bool REREAD>
__device__ __forceinline__ void qb_panel(double (&a)[4][8][2], double* xs, int lane, int g, int t, double& tau_q, unsigned& sgn_q, Acc& Gu) {
  double inv_q = 0.0;
  tau_q = 0.0;
  sgn_q = 0u;
  Gu.x = 0.0;   // strict upper triangle of G = V^T V in accumulator layout, collected from the steps' own dot products:
  Gu.y = 0.0;   // for a finished column g < kk the step's x~^T a_g is v0_g v0_kk (v_g^T v_kk)
  if (g == 0) {
#pragma unroll
    for (int j = P; j < 8; j++) *reinterpret_cast<double2*>(xs + 8 * j + 2 * t) = make_double2(a[P][j][0], a[P][j][1]);
  }
#ifndef QB_PANEL_UNROLL
#define QB_PANEL_UNROLL 2
#endif
  // two steps per trip: the parity of kk (pivot-column buffer, which of a lane's two rows is the pivot row) becomes a
  // compile-time constant.  C4: 1.269 ms with 1, 1.233 with 2, 1.250 with 4, 1.376 with 8 (-DQB_PANEL_UNROLL for A/B builds)
  constexpr int kPanelUnroll = QB_PANEL_UNROLL;
#pragma unroll kPanelUnroll
  for (int kk = 0; kk < 8; kk++) {
    const double* xb = xs + 64 * (kk & 1);
    __syncwarp();
    const double x0 = xb[8 * P + kk];  // pivot element
    double x[8][2];
#ifdef QB_DOT8
    // eight independent chains while the panel is tall (P < 2: 16 / 14 products): depth 2 + 3 adds instead of 4 + 2
    constexpr bool kDot8 = P < 2;
#else
    constexpr bool kDot8 = false;
#endif
    double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0;  // four independent chains
    double d02 = 0.0, d03 = 0.0, d12 = 0.0, d13 = 0.0;
#pragma unroll
    for (int j = P; j < 8; j++) {
      const double2 v = *reinterpret_cast<const double2*>(xb + 8 * j + 2 * t);
      x[j][0] = v.x;
      x[j][1] = v.y;
      const int c = (j - P) & (kDot8 ? 3 : 1);
      if (c == 0) { d00 = fma(v.x, a[P][j][0], d00); d10 = fma(v.y, a[P][j][1], d10); }
      else if (c == 1) { d01 = fma(v.x, a[P][j][0], d01); d11 = fma(v.y, a[P][j][1], d11); }
      else if (c == 2) { d02 = fma(v.x, a[P][j][0], d02); d12 = fma(v.y, a[P][j][1], d12); }
      else { d03 = fma(v.x, a[P][j][0], d03); d13 = fma(v.y, a[P][j][1], d13); }
    }
    double d = kDot8 ? ((d00 + d01) + (d02 + d03)) + ((d10 + d11) + (d12 + d13)) : (d00 + d01) + (d10 + d11);
    // quad totals, and |x|^2 (rows >= k) = the pivot quad's total.  The pivot quad's two half sums are fetched while the own
    // quad finishes its butterfly: one shuffle latency off the chain for one more SHFL (since the transposing accesses of the
    // load / store phases are conflict-free, the shared-memory pipe has the room).
#ifdef QB_S_LATE
    d += shfl_xor(d, 1);
    d += shfl_xor(d, 2);
    const double s = shfl(d, 4 * kk);
#else
    d += shfl_xor(d, 1);
    const double s = shfl(d, 4 * kk) + shfl(d, 4 * kk + 2);
    d += shfl_xor(d, 2);
#endif
    const int e0 = kk & 1, t0 = kk >> 1;
    const double akj = shfl(e0 ? a[P][P][1] : a[P][P][0], (lane & ~3) | t0);  // my column's element in the pivot row
    // 2^-900 < s < inf (else a numerically zero column, H = I; NaN: H = I as well) — on the integer pipe: every FP64 instruction
    // of this chain queues behind the DMMA streams of the neighbouring warps (math-pipe throttle is 45 % of the panel's stalls)
    const bool ok = (unsigned)(__double2hiint(s) - 0x07B00001) < (unsigned)(0x7ff00000 - 0x07B00001);
    const double ss = ok ? s : 1.0;
    // nrm = sqrt(ss), rn = 1/nrm by one cubic step from the 20-bit estimate y0.  The update needs tau / v0^2, and with
    // |v0| = |x0| + nrm, tau = |v0| / nrm that is 1 / q, q = nrm |v0| = ss + |x0| nrm: one reciprocal, estimated early from the
    // approximate norm (the two special-function latencies overlap) and corrected to ~2^-57 by a second-order step at the exact q.
    const double ax0 = fabs(x0);
    const double y0 = rsqrt_approx(ss);
    const double sy = ss * y0;
    const double rq0 = rcp_approx(fma(ax0, sy, ss));
    const double e = fma(-sy, y0, 1.0);
    const double pe = fma(0.375, e, 0.5);
    const double nrm = fma(sy * e, pe, sy);
    const double rn = fma(y0 * e, pe, y0);
    const double cn = copysign(nrm, x0);  // -beta
    const double v0 = x0 + cn;
    const double tau = fma(ax0, rn, 1.0);
    const double q = fma(ax0, nrm, ss);
    const double eq = fma(-q, rq0, 1.0);                 // ~2^-19
    const double rq = fma(rq0, fma(eq, eq, eq), rq0);    // 1 / q
    // v = x~/v0 with x~ = (v0; x below): a -= tau (v^T a) v = a + f x~,  f = -(x~^T a) / q,  x~^T a = d - beta*akj
    const double wall = fma(cn, akj, d);   // x~^T a_g for every column g
    const double wsel = (ok && g > kk) ? wall : 0.0;
    const double f = -(wsel * rq);
    const double inv_v0 = ok ? copysign(nrm * rq, x0) : 0.0;   // 1 / v0 = nrm / q, sign of x0
    {  // G[g][kk] for the finished columns g < kk (inv_q is still 0 for the others); lands in lane (g, kk >> 1)
      const double gv = (wall * inv_v0) * inv_q;
      Gu.x = (t == (kk >> 1) && !(kk & 1)) ? gv : Gu.x;
      Gu.y = (t == (kk >> 1) && (kk & 1)) ? gv : Gu.y;
    }
    const bool prow = (t == t0);
#pragma unroll
    for (int j = P; j < 8; j++) {
      double2 v = REREAD ? lds2_again(xb + 8 * j + 2 * t) : make_double2(x[j][0], x[j][1]);
      if (j == P) {  // x~: v0 in the pivot row (branch-free)
        v.x = (prow && !e0) ? v0 : v.x;
        v.y = (prow && e0) ? v0 : v.y;
      }
      a[P][j][0] = fma(f, v.x, a[P][j][0]);
      a[P][j][1] = fma(f, v.y, a[P][j][1]);
    }
    {  // the next pivot quad publishes its column, rows above its pivot row zeroed (predicated stores)
      const bool nxt = (g == kk + 1);
      double* xn = xs + 64 * ((kk + 1) & 1);
      sts2_if(nxt, xn + 8 * P + 2 * t, (2 * t > kk) ? a[P][P][0] : 0.0, (2 * t + 1 > kk) ? a[P][P][1] : 0.0);
#pragma unroll
      for (int j = P + 1; j < 8; j++) sts2_if(nxt, xn + 8 * j + 2 * t, a[P][j][0], a[P][j][1]);
    }
    {  // the pivot quad records its reflector (selects, no divergence)
      const bool piv = (g == kk);
      tau_q = piv ? (ok ? tau : 0.0) : tau_q;
      inv_q = piv ? inv_v0 : inv_q;
      sgn_q = piv ? ((ok ? ~(unsigned)__double2hiint(cn) : 0u) & 0x80000000u) : sgn_q;  // beta = -cn < 0: flip the row of R / column of Q
      const double beta = ok ? -cn : 0.0;
      a[P][P][0] = (piv && prow && !e0) ? beta : a[P][P][0];
      a[P][P][1] = (piv && prow && e0) ? beta : a[P][P][1];
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

// T of panel P in both orientations (accumulator layout) from the strict upper triangle of G = V^T V that the panel steps
// collected (no Gram product: 2(8-P) dependent DMMAs less per panel).
template <int P>
__device__ __forceinline__ void qb_make_t(const Acc& Gu, int g, int t, double tau_q, Acc& T, Acc& TT) {
  const double tq0 = shfl(tau_q, 4 * (2 * t)), tq1 = shfl(tau_q, 4 * (2 * t + 1));  // tau of columns 2t, 2t+1
  const int c0 = 2 * t, c1 = 2 * t + 1;
  Acc N, I;
  N.x = (g < c0) ? tau_q * Gu.x : 0.0;
  N.y = (g < c1) ? tau_q * Gu.y : 0.0;
  I.x = (g == c0) ? 1.0 : 0.0;
  I.y = (g == c1) ? 1.0 : 0.0;
  const Acc NT = mm8(I, N);   // the transpose, exactly: I * (N^T) given (N^T)^T = N
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
    for (int j = J0; j < 8; j++) {  // two independent accumulation chains (-DQB_SINGLE_W: one, for A/B timing of the pipe contention)
      dmma884(W.x, W.y, a[i][j][0], a[P][j][0]);
#ifdef QB_SINGLE_W
      dmma884(W.x, W.y, a[i][j][1], a[P][j][1]);
#else
      dmma884(W2.x, W2.y, a[i][j][1], a[P][j][1]);
#endif
    }
    W.x += W2.x;
    W.y += W2.y;
    Acc Z = mm8(W, TB);
    Z.x = -Z.x;
    Z.y = -Z.y;
#pragma unroll
    for (int j = P; j < 8; j++) {
      const double2 v = *reinterpret_cast<const double2*>(vs + v_off(P) + (8 * (j - P) + g) * 8 + ((2 * t) ^ v_sw(g)));
      dmma884(a[i][j][0], a[i][j][1], Z.x, v.x);
      dmma884(a[i][j][0], a[i][j][1], Z.y, v.y);
    }
  }
}

// R rows 8P..8P+7 (rows with beta < 0 negated, times post) for column block I.  PADDED: the matrix is rows x cols inside the
// zero-padded 64 x 32 register tile; R is L x cols (L = min(rows, cols)) with row stride cols.
template <int P, int I, bool PADDED>
__device__ __forceinline__ void qb_store_r(const double (&a)[4][8][2], double* __restrict__ r_out, int g, int t,
                                           unsigned s0, unsigned s1, double post, int L, int cols) {
  const int row0 = 8 * P + 2 * t, col = 8 * I + g;
  double v0, v1;
  if (I < P) { v0 = 0.0; v1 = 0.0; }
  else if (I == P) { v0 = (2 * t <= g) ? flip(a[I][P][0], s0) : 0.0; v1 = (2 * t + 1 <= g) ? flip(a[I][P][1], s1) : 0.0; }
  else { v0 = flip(a[I][P][0], s0); v1 = flip(a[I][P][1], s1); }
  if (post != 1.0) { v0 *= post; v1 *= post; }  // warp-uniform, extreme magnitudes only
  if (!PADDED) {
    r_out[row0 * 32 + col] = v0;
    r_out[(row0 + 1) * 32 + col] = v1;
  } else if (col < cols) {
    if (row0 < L) r_out[row0 * cols + col] = v0;
    if (row0 + 1 < L) r_out[(row0 + 1) * cols + col] = v1;
  }
}

template <int P, bool REREAD, bool PADDED, bool LOCK = false>
__device__ __forceinline__ void qb_r_phase(double (&a)[4][8][2], double* vs, double* ts,
                                           double* __restrict__ r_out, int lane, int g, int t, double post, unsigned& sgn_p,
                                           long long& qb_tm, int L, int cols) {
  double tau_q;
  Acc Gu;
  if (LOCK) __syncthreads();   // every warp of the SM enters the panel together (no DMMA stream beside a panel chain)
  qb_panel<P, REREAD>(a, vs + kXsOff, lane, g, t, tau_q, sgn_p, Gu);
  if (LOCK) __syncthreads();
  QB_MARK(2);
  const unsigned s0 = __shfl_sync(kFull, sgn_p, 4 * (2 * t)), s1 = __shfl_sync(kFull, sgn_p, 4 * (2 * t + 1));
  // R: diagonal block and the zero blocks left of it; then the head of the panel becomes the clean unit-lower V
  if (P > 0) qb_store_r<P, 0, PADDED>(a, r_out, g, t, s0, s1, post, L, cols);
  if (P > 1) qb_store_r<P, 1, PADDED>(a, r_out, g, t, s0, s1, post, L, cols);
  if (P > 2) qb_store_r<P, 2, PADDED>(a, r_out, g, t, s0, s1, post, L, cols);
  qb_store_r<P, P, PADDED>(a, r_out, g, t, s0, s1, post, L, cols);
  a[P][P][0] = (2 * t > g) ? a[P][P][0] : ((2 * t == g) ? 1.0 : 0.0);
  a[P][P][1] = (2 * t + 1 > g) ? a[P][P][1] : ((2 * t + 1 == g) ? 1.0 : 0.0);
#pragma unroll
  for (int j = P; j < 8; j++) {
    vs[v_off(P) + (8 * (j - P) + 2 * t) * 8 + (g ^ (2 * t))] = a[P][j][0];      // v_sw(row 2t+e) = 2t
    vs[v_off(P) + (8 * (j - P) + 2 * t + 1) * 8 + (g ^ (2 * t))] = a[P][j][1];
  }
  QB_MARK(3);
  Acc T, TT;
  qb_make_t<P>(Gu, g, t, tau_q, T, TT);
  *reinterpret_cast<double2*>(ts + 64 * P + 2 * lane) = make_double2(T.x, T.y);
  __syncwarp();
  QB_MARK(4);
  if (P < 3) {
    qb_trailing<P, P>(a, vs, g, t, TT);
    if (P < 1) qb_store_r<P, 1, PADDED>(a, r_out, g, t, s0, s1, post, L, cols);
    if (P < 2) qb_store_r<P, 2, PADDED>(a, r_out, g, t, s0, s1, post, L, cols);
    qb_store_r<P, 3, PADDED>(a, r_out, g, t, s0, s1, post, L, cols);
  }
  QB_MARK(5);
}

template <int P>
__device__ __forceinline__ void qb_q_phase(double (&a)[4][8][2], const double* vs, const double* ts,
                                           int lane, int g, int t) {
  const double2 tt = *reinterpret_cast<const double2*>(ts + 64 * P + 2 * lane);
  const Acc T{tt.x, tt.y};
  // columns right of the panel: nonzero from row 8(P+1) on, filled from row 8P on by this update
  if (P < 3) qb_trailing<P, P + 1>(a, vs, g, t, T);
  // the panel's own columns: E_P - V_P (T_P L_P^T)
  const double2 l = *reinterpret_cast<const double2*>(vs + v_off(P) + g * 8 + ((2 * t) ^ v_sw(g)));
  Acc M = mm8(Acc{l.x, l.y}, T);  // (L_P T_P^T)[b][c] = M^T
  M.x = -M.x;
  M.y = -M.y;
#pragma unroll
  for (int j = 0; j < P; j++) { a[P][j][0] = 0.0; a[P][j][1] = 0.0; }
#pragma unroll
  for (int j = P; j < 8; j++) {
    const double2 v = *reinterpret_cast<const double2*>(vs + v_off(P) + (8 * (j - P) + g) * 8 + ((2 * t) ^ v_sw(g)));
    double q0 = (j == P && g == 2 * t) ? 1.0 : 0.0, q1 = (j == P && g == 2 * t + 1) ? 1.0 : 0.0;
    dmma884(q0, q1, M.x, v.x);
    dmma884(q0, q1, M.y, v.y);
    a[P][j][0] = q0;
    a[P][j][1] = q1;
  }
}

// PADDED: any rows <= 64, cols <= 32: the matrix is zero-padded into the 64 x 32 register tile on load (zero rows leave the
// reflectors unchanged, zero columns give identity reflectors after the last real column), and only the rows x L block
// of Q and the L x cols block of R are stored.
template <int WARPS, int MINB, bool REREAD, bool PADDED, bool TMAP, bool LOCK = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
qr64x32_blocked_kernel(const double* __restrict__ A, double* __restrict__ Q, double* __restrict__ R, int64_t batch,
                       int rows, int cols, const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_q) {
  extern __shared__ __align__(16) double qb_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  int64_t m = (int64_t)blockIdx.x * WARPS + warp;
  if (LOCK) { if (m >= batch) m = batch - 1; }   // phase-locked variant: surplus warps redo the last matrix (same bits, benign) to keep the barriers whole
  else if (m >= batch) return;  // warp-uniform; no block-level barriers
  double* vs;
  if (TMAP) {  // 1 KiB-aligned warp regions (SWIZZLE_128B); the launch adds 1 KiB of slack
    const unsigned base = smem_u32(qb_smem), pad = (1024u - (base & 1023u)) & 1023u;
    vs = reinterpret_cast<double*>(reinterpret_cast<char*>(qb_smem) + pad) + warp * kQbWarpDoublesT;
  } else {
    vs = qb_smem + warp * kQbWarpDoubles;
  }
  double* ts = vs + kVDoubles;
  const int L = PADDED ? (rows < cols ? rows : cols) : 32;
  const double* a_in = A + m * (PADDED ? rows * cols : 2048);

  long long qb_tm = clock64();
  (void)qb_tm;
  double a[4][8][2];
  const bool bulk = !TMAP && !PADDED && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Q)) & 15) == 0;  // kernel-uniform
  if (TMAP) {
    if (lane == 0) {
      const unsigned mb = smem_u32(vs + kMbarOff);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(16384u) : "memory");
      tmap_load_2d(vs, &tm_a, 0, (int)(64 * m), vs + kMbarOff);
      tmap_load_2d(vs + 1024, &tm_a, 16, (int)(64 * m), vs + kMbarOff);
    }
    __syncwarp();
    bulk_load_wait(vs + kMbarOff);
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int i = 0; i < 4; i++) a[i][j][e] = vs[tm_idx(i, j, g, t, e)];
    __syncwarp();  // the landing zone becomes the V store
  } else if (bulk) {
    if (lane == 0) bulk_load_start(vs, a_in, 16384, vs + kMbarOff);
    __syncwarp();
    bulk_load_wait(vs + kMbarOff);
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int i = 0; i < 4; i++) a[i][j][e] = vs[(8 * j + 2 * t + e) * 32 + 8 * i + g];
    __syncwarp();  // the landing zone becomes the V store
  } else if (!PADDED) {
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int i = 0; i < 4; i++) a[i][j][e] = ldg1_stream(a_in + (8 * j + 2 * t + e) * 32 + 8 * i + g);
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int row = 8 * j + 2 * t + e, col = 8 * i + g;
          a[i][j][e] = (row < rows && col < cols) ? ldg1_stream(a_in + row * cols + col) : 0.0;
        }
  }

  // scale guard (see pow2_prescale): only the binary exponent of the largest entry matters, so the maximum is taken over
  // the high words as integers (no FP64 instructions) and reduced with one REDUX
  unsigned hi = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      hi = max(hi, (unsigned)__double2hiint(a[i][j][0]) & 0x7fffffffu);
      hi = max(hi, (unsigned)__double2hiint(a[i][j][1]) & 0x7fffffffu);
    }
  hi = __reduce_max_sync(kFull, hi);
  const double amax = __hiloint2double((int)hi, 0);
  const double pre = pow2_prescale(amax);
  if (pre != 1.0) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) { a[i][j][0] *= pre; a[i][j][1] *= pre; }
  }
  const double post = 1.0 / pre;

  double* r_out = R + m * (PADDED ? L * cols : 1024);
  unsigned sgn[4];
  QB_MARK(0);
  qb_r_phase<0, REREAD, PADDED, LOCK>(a, vs, ts, r_out, lane, g, t, post, sgn[0], qb_tm, L, cols);
  qb_r_phase<1, REREAD, PADDED, LOCK>(a, vs, ts, r_out, lane, g, t, post, sgn[1], qb_tm, L, cols);
  qb_r_phase<2, REREAD, PADDED, LOCK>(a, vs, ts, r_out, lane, g, t, post, sgn[2], qb_tm, L, cols);
  qb_r_phase<3, REREAD, PADDED, LOCK>(a, vs, ts, r_out, lane, g, t, post, sgn[3], qb_tm, L, cols);

  qb_q_phase<3>(a, vs, ts, lane, g, t);
  qb_q_phase<2>(a, vs, ts, lane, g, t);
  qb_q_phase<1>(a, vs, ts, lane, g, t);
  qb_q_phase<0>(a, vs, ts, lane, g, t);
  QB_MARK(6);

  double* q_out = Q + m * (PADDED ? rows * L : 2048);
  if (TMAP) {
    __syncwarp();  // every lane is done with the V store
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int i = 0; i < 4; i++) vs[tm_idx(i, j, g, t, e)] = flip(a[i][j][e], sgn[i]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      tmap_store_2d(&tm_q, 0, (int)(64 * m), vs);
      tmap_store_2d(&tm_q, 16, (int)(64 * m), vs + 1024);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else if (bulk) {
    __syncwarp();  // every lane is done with the V store
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int i = 0; i < 4; i++) vs[(8 * j + 2 * t + e) * 32 + 8 * i + g] = flip(a[i][j][e], sgn[i]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) bulk_store(q_out, vs, 16384);
  } else if (!PADDED) {
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int i = 0; i < 4; i++) q_out[(8 * j + 2 * t + e) * 32 + 8 * i + g] = flip(a[i][j][e], sgn[i]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int row = 8 * j + 2 * t + e, col = 8 * i + g;
          if (row < rows && col < L) q_out[row * L + col] = flip(a[i][j][e], sgn[i]);
        }
  }
  QB_MARK(7);
}

// Block reflector P applied to the right-hand-side block Y (8 columns, same register layout as a column block of A):
// Y -= V_P (T_P^T (V_P^T Y)), i.e. Y^T -= ((Y^T V_P) T_P) V_P^T — what qb_trailing does to the column blocks right of a
// panel, here after the factorisation, with V_P still in a[P][j >= P], its row-major copy in vs and T_P in ts.
template <int P>
__device__ __forceinline__ void qb_apply_y(double (&yb)[8][2], const double (&a)[4][8][2], const double* vs, const double* ts,
                                           int g, int t) {
  Acc W{0.0, 0.0}, W2{0.0, 0.0};
#pragma unroll
  for (int j = P; j < 8; j++) {
    dmma884(W.x, W.y, yb[j][0], a[P][j][0]);
    dmma884(W2.x, W2.y, yb[j][1], a[P][j][1]);
  }
  W.x += W2.x;
  W.y += W2.y;
  Acc TT;   // T_P^T in accumulator layout from the row-major T_P
  TT.x = ts[64 * P + 8 * (2 * t) + g];
  TT.y = ts[64 * P + 8 * (2 * t + 1) + g];
  Acc Z = mm8(W, TT);
  Z.x = -Z.x;
  Z.y = -Z.y;
#pragma unroll
  for (int j = P; j < 8; j++) {
    const double2 v = *reinterpret_cast<const double2*>(vs + v_off(P) + (8 * (j - P) + g) * 8 + ((2 * t) ^ v_sw(g)));
    dmma884(yb[j][0], yb[j][1], Z.x, v.x);
    dmma884(yb[j][0], yb[j][1], Z.y, v.y);
  }
}

// _qr_decomp_inplace (src/la/qr.js:147-183) for M <= 64, N <= 32 and at most 8 right-hand sides: the R phases of the
// kernel above on the zero-padded matrix, then Q^T Y = H_4^T .. H_1^T Y from the stored block reflectors — Q is never
// formed.  R is M x N (zero rows below row min(M,N)), Q^T Y is M x L with the rows of the flipped reflectors negated
// like the rows of R (diag(R) >= 0).
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
qr64x32_inplace_kernel(const double* __restrict__ A, const double* __restrict__ Y, double* __restrict__ R,
                       double* __restrict__ QtY, int64_t batch, int rows, int cols, int nrhs) {
  extern __shared__ __align__(16) double qb_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t m = (int64_t)blockIdx.x * WARPS + warp;
  if (m >= batch) return;  // warp-uniform; the kernel has no block-level barriers
  double* vs = qb_smem + warp * kQbWarpDoubles;
  double* ts = vs + kVDoubles;
  const int L = rows < cols ? rows : cols;
  const double* a_in = A + m * (rows * cols);
  long long qb_tm = 0;
  (void)qb_tm;
  double a[4][8][2];
#pragma unroll
  for (int j = 0; j < 8; j++)
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int row = 8 * j + 2 * t + e, col = 8 * i + g;
        a[i][j][e] = (row < rows && col < cols) ? ldg1_stream(a_in + row * cols + col) : 0.0;
      }
  unsigned hi = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      hi = max(hi, (unsigned)__double2hiint(a[i][j][0]) & 0x7fffffffu);
      hi = max(hi, (unsigned)__double2hiint(a[i][j][1]) & 0x7fffffffu);
    }
  hi = __reduce_max_sync(kFull, hi);
  const double pre = pow2_prescale(__hiloint2double((int)hi, 0));
  if (pre != 1.0) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) { a[i][j][0] *= pre; a[i][j][1] *= pre; }
  }
  const double post = 1.0 / pre;

  // R is rows x cols here: the rows below min(rows, cols) are zero
  double* r_out = R + m * (rows * cols);
  for (int e = L * cols + lane; e < rows * cols; e += 32) r_out[e] = 0.0;
  unsigned sgn[4];
  qb_r_phase<0, true, true>(a, vs, ts, r_out, lane, g, t, post, sgn[0], qb_tm, L, cols);
  qb_r_phase<1, true, true>(a, vs, ts, r_out, lane, g, t, post, sgn[1], qb_tm, L, cols);
  qb_r_phase<2, true, true>(a, vs, ts, r_out, lane, g, t, post, sgn[2], qb_tm, L, cols);
  qb_r_phase<3, true, true>(a, vs, ts, r_out, lane, g, t, post, sgn[3], qb_tm, L, cols);

  double yb[8][2];
  const double* y_in = Y + m * (rows * nrhs);
#pragma unroll
  for (int j = 0; j < 8; j++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int row = 8 * j + 2 * t + e;
      yb[j][e] = (row < rows && g < nrhs) ? ldg1_stream(y_in + row * nrhs + g) : 0.0;
    }
  qb_apply_y<0>(yb, a, vs, ts, g, t);
  qb_apply_y<1>(yb, a, vs, ts, g, t);
  qb_apply_y<2>(yb, a, vs, ts, g, t);
  qb_apply_y<3>(yb, a, vs, ts, g, t);
  double* y_out = QtY + m * (rows * nrhs);
#pragma unroll
  for (int j = 0; j < 8; j++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int row = 8 * j + 2 * t + e;
      double v = yb[j][e];
      if (j < 4) v = flip(v, __shfl_sync(kFull, sgn[j < 4 ? j : 0], 4 * (2 * t + e)));   // row 8j+c belongs to reflector c of panel j
      if (row < rows && g < nrhs) y_out[row * nrhs + g] = v;
    }
}

}  // namespace

// 2-D tensor map over a batch of 64 x 32 matrices seen as one [batch*64, 32] row-major array: box = 16 columns x 64 rows
// (one 128-byte line per row), SWIZZLE_128B.  The encoder comes from the driver through the runtime (no libcuda link).
static bool qb_encode_map(CUtensorMap* tm, const void* base, int64_t batch) {
  typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode enc = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<Encode>(fn);
    else
      (void)cudaGetLastError();
  }
  if (!enc) return false;
  const cuuint64_t dims[2] = {32, (cuuint64_t)batch * 64};
  const cuuint64_t strides[1] = {256};
  const cuuint32_t box[2] = {16, 64}, estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int WARPS, int MINB, bool REREAD, bool PADDED, bool TMAP = false, bool LOCK = false>
static cudaError_t qb_launch(cudaStream_t s, const double* A, double* Q, double* R, int64_t batch, int rows, int cols,
                             const CUtensorMap* tm_a = nullptr, const CUtensorMap* tm_q = nullptr) {
  static bool attr_set[64] = {false};
  constexpr size_t smem = TMAP ? sizeof(double) * kQbWarpDoublesT * WARPS + 1024 : sizeof(double) * kQbWarpDoubles * WARPS;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(qr64x32_blocked_kernel<WARPS, MINB, REREAD, PADDED, TMAP, LOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  static const CUtensorMap none{};
  qr64x32_blocked_kernel<WARPS, MINB, REREAD, PADDED, TMAP, LOCK><<<(unsigned)((batch + WARPS - 1) / WARPS), WARPS * 32, smem, s>>>(
      A, Q, R, batch, rows, cols, tm_a ? *tm_a : none, tm_q ? *tm_q : none);
  return cudaGetLastError();
}

// variant 2: two CTAs per SM (216 registers, no spills, pivot column kept in registers) instead of three (168 registers,
// idle column blocks parked in local memory around the panel loops, pivot column read twice from shared memory): 1.38 vs 1.27 ms on C4.  Phase-locking the warps of
// a sub-partition with named barriers (so that no DMMA stream runs beside a panel chain) was measured slower (1.55-1.70 ms):
// the panel chains of the locked warps then collide on the shared-memory pipe instead.
cudaError_t launch_qr64x32_blocked(cudaStream_t s, const double* A, double* Q, double* R, int64_t batch, int variant) {
  if (variant == 2) return qb_launch<4, 2, false, false>(s, A, Q, R, batch, 64, 32);
  // default: TMA tile loads / stores through swizzled tensor maps (ND4B_QR_TMAP=0: the linear bulk copies of round 1)
  static const bool want_tmap = [] { const char* e = getenv("ND4B_QR_TMAP"); return !(e && e[0] == '0'); }();
  const bool aligned = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Q)) & 15) == 0;
  if (want_tmap && aligned && batch * 64 < 0x7fffffffLL) {
    CUtensorMap tm_a, tm_q;
    if (qb_encode_map(&tm_a, A, batch) && qb_encode_map(&tm_q, Q, batch)) {
      // variant 3 (A/B): one CTA of 12 warps per SM, CTA-wide barriers around every panel — no DMMA stream beside a panel chain
      if (variant == 3) return qb_launch<12, 1, true, false, true, true>(s, A, Q, R, batch, 64, 32, &tm_a, &tm_q);
      // CTAs of 2 warps, 6 per SM: a finished matrix frees its slot sooner (measured: 1 warp per CTA 1.071 ms, 2: 1.034, 3: 1.090 —
      // three warps leave a sub-partition short —, 4: 1.046, 6: 1.063)
      return qb_launch<2, 6, true, false, true>(s, A, Q, R, batch, 64, 32, &tm_a, &tm_q);
    }
  }
  return qb_launch<4, 3, true, false>(s, A, Q, R, batch, 64, 32);
}

// _qr_decomp_inplace for M <= 64, N <= 32, L <= 8 through the register kernel (R phases + block reflectors applied to Y)
cudaError_t launch_qr_inplace_blocked(cudaStream_t s, const double* A, const double* Y, double* R, double* QtY,
                                      int64_t batch, int rows, int cols, int nrhs) {
  static bool attr_set[64] = {false};
  constexpr size_t smem = sizeof(double) * kQbWarpDoubles * 4;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(qr64x32_inplace_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  qr64x32_inplace_kernel<4, 3><<<(unsigned)((batch + 3) / 4), 128, smem, s>>>(A, Y, R, QtY, batch, rows, cols, nrhs);
  return cudaGetLastError();
}

// rows <= 64, cols <= 32 through the zero-padded tile: the cost of a 64 x 32 factorisation whatever the shape
cudaError_t launch_qr_padded_blocked(cudaStream_t s, const double* A, double* Q, double* R, int64_t batch, int rows, int cols) {
  return qb_launch<4, 3, true, true>(s, A, Q, R, batch, rows, cols);
}

}  // namespace nd4b
