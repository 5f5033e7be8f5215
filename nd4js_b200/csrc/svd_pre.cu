// svd_pre.cu — single-precision preconditioner of the 64x64 one-sided Jacobi SVD (BASELINE config C5).
//
// A Jacobi sweep costs the same whether it rotates by 40 degrees or by 1e-9, and of the ~10 sweeps the FP64 kernel
// (svd.cu, svd64cb_kernel) needs on a random 64x64 matrix all but the last two only bring the off-diagonal cosines from
// O(1) down to ~1e-6 — work that does not need 53 bits.  So the sweeps are split:
//
//   1. svd64_pre32_kernel : one-sided Jacobi (block round-robin ordering with fixed register slots, square-root-free scaled
//                           rotations, warp-owned columns in registers) on an FP32 copy of A, with packed FP32x2 arithmetic (FFMA2 /
//                           FMUL2: the two rows a lane owns of a column are one 64-bit register pair), until a sweep sees no
//                           cosine above 1e-2 (quadratic convergence then leaves ~1e-6): ~8 cheap sweeps.  Only the
//                           accumulated rotation V0 (FP32, orthogonal to ~1e-5) leaves the kernel.
//   2. svd64_ortho_kernel : FP64, DMMA from shared memory: V1 = V0 (I + E)^(-1/2) with E = V0^T V0 - I, as the polynomial
//                           I - E/2 + 3 E^2 / 8 (- 5 E^3 / 16 when |E|_F says it matters) — orthogonal to working precision —
//                           and G1 = A V1, whose columns are orthogonal to ~1e-6.
//   3. svd64cb_kernel<.., PRE> (svd.cu): the FP64 Jacobi started from (G1, V1) instead of (A, I): two rotating sweeps and
//                           the confirming one, the same convergence test and epilogue as before, so the results obey the
//                           same bounds.  Any orthogonal V1 is a valid start: if step 1 or 2 goes wrong (NaN input, |E| not
//                           small) step 2 hands over (A, I) and step 3 is the plain algorithm.
//
// Contract, ordering and sign rules: see svd.cu (nd4js src/la/svd_jac_2sided.js:30-144, _svd_jac_utils.js:123-188).
#include "common.cuh"
#include "kernels.h"
#include <float.h>

namespace nd4b {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float hsum2(u64 x) { float lo, hi; unpack2(x, lo, hi); return lo + hi; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int N>
__device__ __forceinline__ void halvef(float (&v)[N], bool bit, int mask) {
#pragma unroll
  for (int k = 0; k < N / 2; k++) {
    const float send = bit ? v[k] : v[k + N / 2];
    const float keep = bit ? v[k + N / 2] : v[k];
    v[k] = keep + __shfl_xor_sync(kFull, send, mask);
  }
}

// every lane of 4-lane group grp ends up with the warp total of element grp
__device__ __forceinline__ float reduce8f(float (&pd)[8], int lane) {
  halvef<8>(pd, (lane & 16) != 0, 16);
  float h4[4] = {pd[0], pd[1], pd[2], pd[3]};
  halvef<4>(h4, (lane & 8) != 0, 8);
  float h2[2] = {h4[0], h4[1]};
  halvef<2>(h2, (lane & 4) != 0, 4);
  float d = h2[0];
  d += __shfl_xor_sync(kFull, d, 2);
  d += __shfl_xor_sync(kFull, d, 1);
  return d;
}

// Ordering.  The 64 columns are eight blocks of eight; a warp holds two blocks, A and B, in FIXED registers (slot i of a
// block never moves inside a round), lane l owning rows l and l+32 (one packed register pair) of every slot of G-hat and
// V-hat.  A sweep is seven rounds of a round-robin tournament between the blocks (warp w plays positions w and 7-w;
// after each round the blocks move one position on through shared memory, two barriers per round), and a round is the
// eight "cross" steps  { (A_i, B_(i+s) mod 8) : i = 0..7 },  s = 0..7, which meet every pair of the two blocks once; the
// 2 x 28 pairs inside the blocks are met in the first round by seven steps  { (X_i, X_(i xor m)) },  m = 1..7.  63 steps
// of 32 disjoint pairs = all 2016 pairs, like the odd-even ordering it replaces, but no column ever changes its
// register inside a round: the step bodies are unrolled with compile-time slot indices (the exchange of the odd-even
// ordering cost a register move per element and step, two barriers and a boundary column through shared memory per step).
//
// Scaled ("fast") rotations as in svd.cu: column = D * g-hat, a rotation is  p -= alpha q;  q += beta p  (one packed FMA
// each), D folded back at the end of every round.  The 4-lane group g keeps the cached true |.|^2, D and 1/D of A-slot g
// and of the B-slot it currently plays against (ownership of the B side moves one group on per cross step: 3 shuffles).
struct P32Col { float n, d, di; };   // cached |g|^2, scale D, 1/D of one column

constexpr float kPreTol2 = 1.4551915e-11f;   // (64 * 2^-24)^2: rotate while cos^2 is above the FP32 noise
constexpr float kPreBig2 = 1e-4f;            // a sweep that saw no cos^2 above this is the last one (then ~1e-6 remains)
constexpr int kPreMaxSweeps = 12;

// Threshold test and rotation set-up of one pair (p, q) from the true norms and the scaled dot product; updates the two
// column records and returns (-alpha, -alpha, beta, beta), zeros when the pair is not rotated.
__device__ __forceinline__ float4 p32_params(float dhat, P32Col& p, P32Col& q, int& big) {
  const float d = dhat * p.d * q.d;
  const float ab = p.n * q.n;
  const bool rot = d * d > kPreTol2 * ab && ab > 1e-25f;
  if (d * d > kPreBig2 * ab) big = 1;
  float alpha = 0.f, beta = 0.f;
  if (__any_sync(kFull, rot)) {   // warp-uniform set-up; lanes below the threshold discard it
    const float num = q.n - p.n, den = 2.f * d;
    const float im = rcp_approx(fmaxf(fabsf(num), fabsf(den)));
    const float n1 = fabsf(num) * im, d1 = fabsf(den) * im;      // the larger one is 1
    const float S = fmaf(n1, n1, d1 * d1);                       // in [1, 2]
    const float y = rsqrt_approx(S);
    float t = d1 * rcp_approx(fmaf(S, y, n1));                   // |den| / (|num| + sqrt(num^2 + den^2))
    t = __int_as_float(__float_as_int(t) | ((__float_as_int(num) ^ __float_as_int(den)) & 0x80000000));
    const float w = fmaf(t, t, 1.f);
    float c = rsqrt_approx(w);
    c = c * fmaf(-0.5f * w * c, c, 1.5f);                        // one Newton step: V0 stays orthogonal to ~1e-5 over a run
    const float rc = w * c;
    if (rot) {
      alpha = t * q.d * p.di;
      beta = t * p.d * q.di;
      p.n = fmaxf(fmaf(-t, d, p.n), 0.f);                        // |c p - s q|^2 = |p|^2 - t d,  |s p + c q|^2 = |q|^2 + t d
      q.n = fmaxf(fmaf(t, d, q.n), 0.f);
      p.d *= c; q.d *= c; p.di *= rc; q.di *= rc;
    }
  }
  return make_float4(-alpha, -alpha, beta, beta);
}

// (p, q) -> (p - alpha q, q + beta p) on both rows at once
__device__ __forceinline__ void rot2(u64& p, u64& q, u64 nalpha2, u64 beta2) {
  const u64 np = fma2(nalpha2, q, p);
  q = fma2(beta2, p, q);
  p = np;
}

constexpr size_t kPreSmem = 4 * 2 * 16 * 32 * sizeof(u64) + 4 * 16 * sizeof(float) + 4 * 8 * sizeof(float4) + 4 * sizeof(double);

// One CTA (4 warps) per matrix.  V0 (FP32, row-major [component][column slot]) is the only result.
__global__ void __launch_bounds__(128, 4)
svd64_pre32_kernel(const double* __restrict__ A, float* __restrict__ V0, int64_t batch, unsigned long long* sweep_sum) {
  extern __shared__ __align__(16) unsigned char pre_smem[];
  u64* xbuf = reinterpret_cast<u64*>(pre_smem);                       // [warp][block A/B][g 8 slots, v 8 slots][lane]
  float* nbuf = reinterpret_cast<float*>(xbuf + 4 * 2 * 16 * 32);      // [warp][16] norms travelling with the blocks
  float4* wcs_all = reinterpret_cast<float4*>(nbuf + 4 * 16);          // [warp][8] multipliers of the current step
  double* red = reinterpret_cast<double*>(wcs_all + 4 * 8);
  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = lane >> 2;
  const double* a_in = A + m * 4096;
  float4* wcs = wcs_all + 8 * warp;
  float* wsc = reinterpret_cast<float*>(wcs);   // the same 128 bytes hold the 16 slot scales when a round is folded

  u64 ga[8], gb[8], va[8], vb[8];
  {
    double t0[16], t1[16];
    double amax = 0.0;
#pragma unroll
    for (int s = 0; s < 16; s += 2) {
      const double2 x0 = ldg2(a_in + lane * 64 + 16 * warp + s);          // read again by the FP64 kernels: keep it in L2
      const double2 x1 = ldg2(a_in + (lane + 32) * 64 + 16 * warp + s);
      t0[s] = x0.x; t0[s + 1] = x0.y; t1[s] = x1.x; t1[s + 1] = x1.y;
      amax = fmax(amax, fmax(fmax(fabs(x0.x), fabs(x0.y)), fmax(fabs(x1.x), fabs(x1.y))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, shfl_xor(amax, o));
    if (lane == 0) red[warp] = amax;
    __syncthreads();
    amax = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
    // exact power of two that brings the largest entry into [1, 2): the FP32 copy then neither overflows nor loses the
    // matrix to underflow (entries below 2^-126 of the largest become zero, which only costs FP64 sweeps later)
    double pre = 1.0;
    if (amax > 0.0 && amax < CUDART_INF) { const int e = -ilogb(amax); pre = scalbn(1.0, e > 1000 ? 1000 : e); }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      ga[i] = pack2((float)(t0[i] * pre), (float)(t1[i] * pre));
      gb[i] = pack2((float)(t0[8 + i] * pre), (float)(t1[8 + i] * pre));
      va[i] = pack2((16 * warp + i == lane) ? 1.f : 0.f, (16 * warp + i == lane + 32) ? 1.f : 0.f);
      vb[i] = pack2((16 * warp + 8 + i == lane) ? 1.f : 0.f, (16 * warp + 8 + i == lane + 32) ? 1.f : 0.f);
    }
  }
  const int next_grp_lane = (((grp + 1) & 7) << 2) | (lane & 3);

  int sweeps = 0;
  bool more = true;
  while (sweeps < kPreMaxSweeps && more) {
    sweeps++;
    int big = 0;
    P32Col ca, cb;   // A-slot grp, and the B-slot this group currently plays against (B-slot grp between rounds)
    {  // exact norms: 16 values -> (A-slot grp, B-slot grp) in every lane of group grp
      float n2[16];
#pragma unroll
      for (int i = 0; i < 8; i++) { n2[2 * i] = hsum2(mul2(ga[i], ga[i])); n2[2 * i + 1] = hsum2(mul2(gb[i], gb[i])); }
      halvef<16>(n2, (lane & 16) != 0, 16);
      float h8[8];
#pragma unroll
      for (int k = 0; k < 8; k++) h8[k] = n2[k];
      halvef<8>(h8, (lane & 8) != 0, 8);
      float h4[4] = {h8[0], h8[1], h8[2], h8[3]};
      halvef<4>(h4, (lane & 4) != 0, 4);
      float e = h4[0], o = h4[1];
      e += __shfl_xor_sync(kFull, e, 2); o += __shfl_xor_sync(kFull, o, 2);
      e += __shfl_xor_sync(kFull, e, 1); o += __shfl_xor_sync(kFull, o, 1);
      ca.n = e; cb.n = o;
    }
    ca.d = ca.di = cb.d = cb.di = 1.f;

#pragma unroll 1
    for (int round = 0; round < 7; round++) {
      if (round == 0) {
        // pairs inside the blocks: step m pairs slot i with slot i xor m.  Groups i < i^m take the pair of block A, groups
        // i > i^m that of block B; the partner group's record of the other column comes and goes by shuffles.
#pragma unroll
        for (int msk = 1; msk < 8; msk++) {
          float pd[8];
#pragma unroll
          for (int j = 0; j < 8; j++)
            pd[j] = (j < (j ^ msk)) ? hsum2(mul2(ga[j], ga[j ^ msk])) : hsum2(mul2(gb[j ^ msk], gb[j]));
          const float dhat = reduce8f(pd, lane);
          const bool lower = grp < (grp ^ msk);
          P32Col rem;   // lower: the partner's A record (= q);  upper: the partner's B record (= p)
          rem.n = __shfl_xor_sync(kFull, lower ? cb.n : ca.n, 4 * msk);
          rem.d = __shfl_xor_sync(kFull, lower ? cb.d : ca.d, 4 * msk);
          rem.di = __shfl_xor_sync(kFull, lower ? cb.di : ca.di, 4 * msk);
          P32Col p = lower ? ca : rem, q = lower ? rem : cb;
          const float4 mult = p32_params(dhat, p, q, big);
          if (lower) ca = p; else cb = q;
          const P32Col back = lower ? q : p;
          rem.n = __shfl_xor_sync(kFull, back.n, 4 * msk);
          rem.d = __shfl_xor_sync(kFull, back.d, 4 * msk);
          rem.di = __shfl_xor_sync(kFull, back.di, 4 * msk);
          if (lower) cb = rem; else ca = rem;
          __syncwarp();
          #ifndef PRE32_WCS4
          if ((lane & 2) == 0) reinterpret_cast<float*>(wcs)[2 * grp + (lane & 1)] = (lane & 1) ? mult.z : mult.x;   // (-alpha, beta): 8 bytes per group, read back with uniform 8-byte loads and duplicated in registers — half the shared-memory wavefronts of the (-alpha, -alpha, beta, beta) form, which was the largest share of this kernel's LSU traffic (C5 15.82 -> 15.23 ms; -DPRE32_WCS4 for A/B)
#else
          reinterpret_cast<float*>(wcs)[lane] = (lane & 2) ? mult.z : mult.x;   // (-alpha, -alpha, beta, beta) of group grp: one 128-byte wavefront
#endif
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; j++) {
#ifndef PRE32_WCS4
            const float2 ab2 = reinterpret_cast<const float2*>(wcs)[j];
            const u64 na2 = pack2(ab2.x, ab2.x), b2 = pack2(ab2.y, ab2.y);
#else
            const float4 ab = wcs[j];
            const u64 na2 = pack2(ab.x, ab.y), b2 = pack2(ab.z, ab.w);
#endif
            if (j < (j ^ msk)) { rot2(ga[j], ga[j ^ msk], na2, b2); rot2(va[j], va[j ^ msk], na2, b2); }
            else { rot2(gb[j ^ msk], gb[j], na2, b2); rot2(vb[j ^ msk], vb[j], na2, b2); }
          }
        }
      }
      // the 64 pairs between the two blocks
#pragma unroll
      for (int s = 0; s < 8; s++) {
        float pd[8];
#pragma unroll
        for (int i = 0; i < 8; i++) pd[i] = hsum2(mul2(ga[i], gb[(i + s) & 7]));
        const float dhat = reduce8f(pd, lane);
        const float4 mult = p32_params(dhat, ca, cb, big);
        // group g played B-slot (g + s) mod 8; in the next step that slot belongs to group g - 1
        cb.n = __shfl_sync(kFull, cb.n, next_grp_lane);
        cb.d = __shfl_sync(kFull, cb.d, next_grp_lane);
        cb.di = __shfl_sync(kFull, cb.di, next_grp_lane);
        __syncwarp();
        #ifndef PRE32_WCS4
          if ((lane & 2) == 0) reinterpret_cast<float*>(wcs)[2 * grp + (lane & 1)] = (lane & 1) ? mult.z : mult.x;   // (-alpha, beta): 8 bytes per group, read back with uniform 8-byte loads and duplicated in registers — half the shared-memory wavefronts of the (-alpha, -alpha, beta, beta) form, which was the largest share of this kernel's LSU traffic (C5 15.82 -> 15.23 ms; -DPRE32_WCS4 for A/B)
#else
          reinterpret_cast<float*>(wcs)[lane] = (lane & 2) ? mult.z : mult.x;   // (-alpha, -alpha, beta, beta) of group grp: one 128-byte wavefront
#endif
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) {
#ifndef PRE32_WCS4
          const float2 ab2 = reinterpret_cast<const float2*>(wcs)[i];
          const u64 na2 = pack2(ab2.x, ab2.x), b2 = pack2(ab2.y, ab2.y);
#else
          const float4 ab = wcs[i];
          const u64 na2 = pack2(ab.x, ab.y), b2 = pack2(ab.z, ab.w);
#endif
          rot2(ga[i], gb[(i + s) & 7], na2, b2);
          rot2(va[i], vb[(i + s) & 7], na2, b2);
        }
      }
      // fold the scales back into the columns (after eight ownership moves group g holds B-slot g again)
      __syncwarp();
      if ((lane & 3) == 0) { wsc[grp] = ca.d; wsc[8 + grp] = cb.d; }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float da = wsc[i], db = wsc[8 + i];
        const u64 da2 = pack2(da, da), db2 = pack2(db, db);
        ga[i] = mul2(ga[i], da2); va[i] = mul2(va[i], da2);
        gb[i] = mul2(gb[i], db2); vb[i] = mul2(vb[i], db2);
      }
      ca.d = ca.di = cb.d = cb.di = 1.f;
      // Next round's partners.  The eight blocks play a round-robin tournament (circle method), but WHICH WARP hosts a pair is
      // free, and for any two consecutive rounds the hosts can be chosen so that every warp keeps one of its two blocks in its
      // slot and only swaps the other: after even rounds warps 0, 2 send their B block and warps 1, 3 their A block around the
      // cycle w1.A -> w0.B -> w2.B -> w3.A -> w1.A, after odd rounds warp 2 sends A and the others B around
      // w2.A -> w0.B -> w1.B -> w3.B -> w2.A; seven rounds of this (and the same again for the next sweep, from wherever the
      // blocks then are) meet every pair of blocks exactly once.  Half the shared-memory traffic of moving every block one
      // position on (which sent 2 and fetched 2 blocks per warp), and with the buffer alternating by round parity one barrier.
      {
        const bool odd = (round & 1) != 0;
        const bool send_a = odd ? (warp == 2) : ((warp & 1) != 0);
        const int src = odd ? ((0x1302 >> (4 * warp)) & 3) : ((0x2031 >> (4 * warp)) & 3);   // E: 1,3,0,2   O: 2,0,3,1
        u64* mine = xbuf + (round & 1) * (4 * 16 * 32) + warp * (16 * 32);
        float* nb = nbuf + (round & 1) * 32;
        if (send_a) {
#pragma unroll
          for (int i = 0; i < 8; i++) { mine[i * 32 + lane] = ga[i]; mine[(8 + i) * 32 + lane] = va[i]; }
          if ((lane & 3) == 0) nb[warp * 8 + grp] = ca.n;
        } else {
#pragma unroll
          for (int i = 0; i < 8; i++) { mine[i * 32 + lane] = gb[i]; mine[(8 + i) * 32 + lane] = vb[i]; }
          if ((lane & 3) == 0) nb[warp * 8 + grp] = cb.n;
        }
        __syncthreads();
        const u64* in = xbuf + (round & 1) * (4 * 16 * 32) + src * (16 * 32);
        if (send_a) {
#pragma unroll
          for (int i = 0; i < 8; i++) { ga[i] = in[i * 32 + lane]; va[i] = in[(8 + i) * 32 + lane]; }
          ca.n = nb[src * 8 + grp];
        } else {
#pragma unroll
          for (int i = 0; i < 8; i++) { gb[i] = in[i * 32 + lane]; vb[i] = in[(8 + i) * 32 + lane]; }
          cb.n = nb[src * 8 + grp];
        }
      }
    }
    more = __syncthreads_or(big) != 0;
  }
  if (tid == 0 && sweep_sum) atomicAdd(sweep_sum, (unsigned long long)sweeps);

  // the column order of V0 is the final arrangement of the blocks; G1 = A V1 inherits it and the epilogue of the FP64
  // kernel sorts by singular value anyway
  float* v_out = V0 + m * 4096;
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int s = 0; s < 8; s += 4) {
      float lo[4], hi[4];
#pragma unroll
      for (int k = 0; k < 4; k++) unpack2(h ? vb[s + k] : va[s + k], lo[k], hi[k]);
      *reinterpret_cast<float4*>(v_out + lane * 64 + 16 * warp + 8 * h + s) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<float4*>(v_out + (lane + 32) * 64 + 16 * warp + 8 * h + s) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// V1 = V0 (I + E)^(-1/2),  G1 = A V1  — FP64 on the DMMA pipe, operands in shared memory.
// 8 warps per matrix; warp w owns the 8 output rows 8w..8w+7 (eight 8x8 tiles), lane = 4g + t holds C[g][2t], C[g][2t+1]
// of every tile.  Row stride 68 doubles (4 mod 16) for every operand: an 8-byte fragment load is served half a warp at a time
// (lanes g < 4, t), and both the B pattern [k0+t][n0+g] and the A pattern [i0+g][k0+t] then touch the sixteen double-wide banks
// 4t + g / 4g + t once each.  (Round 2 first used 72 for the B operands — 8 mod 16 puts t and t+2 on the same banks:
// 11 264 of the 22 528 fragment-load wavefronts per matrix were conflicts and the kernel ran at 89 % of the LSU pipe,
// profiles/r02_ncu_c5_pipeline.txt.)  E is symmetric, so wherever E is the A operand it is read through its transpose.
// ------------------------------------------------------------------------------------------------
constexpr int kOrthoLD = 68, kOrthoLDA = 68;
constexpr size_t kOrthoSmem = sizeof(double) * (3 * 64 * kOrthoLD + 16);

template <class FA, class FB>
__device__ __forceinline__ void tile_row_product(double (&acc)[8][2], FA fa, FB fb) {
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll 2
  for (int k0 = 0; k0 < 64; k0 += 4) {
    const double a = fa(k0);
#pragma unroll
    for (int j = 0; j < 8; j++) dmma884(acc[j][0], acc[j][1], a, fb(k0, j));
  }
}

__global__ void __launch_bounds__(256, 2)
svd64_ortho_kernel(const double* __restrict__ A, const float* __restrict__ V0, double* __restrict__ G1, double* __restrict__ V1,
                   int64_t batch) {
  extern __shared__ __align__(16) double osm[];
  double* Bv = osm;                         // V0 (stride 72), later A (stride 68)
  double* Be = Bv + 64 * kOrthoLD;          // E, later V1
  double* Bp = Be + 64 * kOrthoLD;          // E^2, then P
  double* red = Bp + 64 * kOrthoLD;         // [8] partial |E|_F^2
  const int64_t m = blockIdx.x;
  if (m >= batch) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int i0 = 8 * warp;
  const double* a_in = A + m * 4096;
  const float* v_in = V0 + m * 4096;
  double* g_out = G1 + m * 4096;
  double* v_out = V1 + m * 4096;

  // A in registers until V0's buffer is free (row r = idx / 32, 16-byte column pair c2 = idx % 32)
  double2 areg[8];
#pragma unroll
  for (int r = 0; r < 8; r++) areg[r] = ldg2_stream(a_in + 2 * (tid + 256 * r));
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int idx = tid + 256 * r, row = idx >> 4, c4 = (idx & 15) * 4;
    const float4 x = __ldcs(reinterpret_cast<const float4*>(v_in) + idx);
    double* d = Bv + row * kOrthoLD + c4;
    *reinterpret_cast<double2*>(d) = make_double2((double)x.x, (double)x.y);
    *reinterpret_cast<double2*>(d + 2) = make_double2((double)x.z, (double)x.w);
  }
  __syncthreads();

  double acc[8][2];
  // E = V0^T V0 - I
  tile_row_product(acc, [&](int k0) { return Bv[(k0 + t) * kOrthoLD + i0 + g]; },
                   [&](int k0, int j) { return Bv[(k0 + t) * kOrthoLD + 8 * j + g]; });
  double ne = 0.0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int col = 8 * j + 2 * t;
    const double e0 = acc[j][0] - ((i0 + g == col) ? 1.0 : 0.0), e1 = acc[j][1] - ((i0 + g == col + 1) ? 1.0 : 0.0);
    ne = fma(e0, e0, fma(e1, e1, ne));
    *reinterpret_cast<double2*>(Be + (i0 + g) * kOrthoLD + col) = make_double2(e0, e1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ne += shfl_xor(ne, o);
  if (lane == 0) red[warp] = ne;
  __syncthreads();
  ne = 0.0;
#pragma unroll
  for (int w = 0; w < 8; w++) ne += red[w];

  if (!(ne <= 1e-4)) {
    // V0 is not close to orthogonal (NaN / Inf input, or the FP32 run went wrong): hand over the plain start (A, I)
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int idx = tid + 256 * r, row = idx >> 5, c2 = (idx & 31) * 2;
      *reinterpret_cast<double2*>(g_out + 2 * idx) = areg[r];
      *reinterpret_cast<double2*>(v_out + 2 * idx) = make_double2(row == c2 ? 1.0 : 0.0, row == c2 + 1 ? 1.0 : 0.0);
    }
    return;
  }

  // F = E E (E read through its transpose as the A operand)
  tile_row_product(acc, [&](int k0) { return Be[(k0 + t) * kOrthoLD + i0 + g]; },
                   [&](int k0, int j) { return Be[(k0 + t) * kOrthoLD + 8 * j + g]; });
  // The neglected term of (I + E)^(-1/2) leaves 5/8 |E^3| in V1^T V1 - I: below ~1e-15 it is noise
  const bool cubic = ne * sqrt(ne) > 2e-15;   // CTA-uniform
  if (cubic) {
#pragma unroll
    for (int j = 0; j < 8; j++)
      *reinterpret_cast<double2*>(Bp + (i0 + g) * kOrthoLD + 8 * j + 2 * t) = make_double2(acc[j][0], acc[j][1]);
    __syncthreads();
    double f3[8][2];
    tile_row_product(f3, [&](int k0) { return Be[(k0 + t) * kOrthoLD + i0 + g]; },
                     [&](int k0, int j) { return Bp[(k0 + t) * kOrthoLD + 8 * j + g]; });
    __syncthreads();   // every warp has read F before P overwrites it
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[j][0] = fma(-5.0 / 6.0, f3[j][0], acc[j][0]); acc[j][1] = fma(-5.0 / 6.0, f3[j][1], acc[j][1]); }   // 3/8 (F - 5/6 E^3)
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int col = 8 * j + 2 * t;
    const double2 e = *reinterpret_cast<const double2*>(Be + (i0 + g) * kOrthoLD + col);
    const double p0 = fma(0.375, acc[j][0], fma(-0.5, e.x, (i0 + g == col) ? 1.0 : 0.0));
    const double p1 = fma(0.375, acc[j][1], fma(-0.5, e.y, (i0 + g == col + 1) ? 1.0 : 0.0));
    *reinterpret_cast<double2*>(Bp + (i0 + g) * kOrthoLD + col) = make_double2(p0, p1);
  }
  __syncthreads();

  // V1 = V0 P  -> global and Be
  tile_row_product(acc, [&](int k0) { return Bv[(i0 + g) * kOrthoLD + k0 + t]; },
                   [&](int k0, int j) { return Bp[(k0 + t) * kOrthoLD + 8 * j + g]; });
  __syncthreads();   // all reads of V0 (Bv) and of E (Be) are done
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int col = 8 * j + 2 * t;
    *reinterpret_cast<double2*>(Be + (i0 + g) * kOrthoLD + col) = make_double2(acc[j][0], acc[j][1]);
    *reinterpret_cast<double2*>(v_out + (i0 + g) * 64 + col) = make_double2(acc[j][0], acc[j][1]);
  }
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int idx = tid + 256 * r, row = idx >> 5, c2 = (idx & 31) * 2;
    *reinterpret_cast<double2*>(Bv + row * kOrthoLDA + c2) = areg[r];
  }
  __syncthreads();

  // G1 = A V1
  tile_row_product(acc, [&](int k0) { return Bv[(i0 + g) * kOrthoLDA + k0 + t]; },
                   [&](int k0, int j) { return Be[(k0 + t) * kOrthoLD + 8 * j + g]; });
#pragma unroll
  for (int j = 0; j < 8; j++)
    *reinterpret_cast<double2*>(g_out + (i0 + g) * 64 + 8 * j + 2 * t) = make_double2(acc[j][0], acc[j][1]);
}

// workspace per matrix: V0 (4096 floats), G1, V1 (4096 doubles each)
size_t svd64_pre_workspace_bytes(int64_t batch) { return (size_t)batch * (4096 * 4 + 2 * 4096 * 8); }

cudaError_t launch_svd64_pre(cudaStream_t s, const double* A, int64_t batch, double* work, unsigned long long* sweep_sum,
                             const double** G1, const double** V1) {
  float* v0 = reinterpret_cast<float*>(work);
  double* g1 = work + (size_t)batch * 2048;
  double* v1 = g1 + (size_t)batch * 4096;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(svd64_ortho_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOrthoSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(svd64_pre32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPreSmem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  svd64_pre32_kernel<<<(unsigned)batch, 128, kPreSmem, s>>>(A, v0, batch, sweep_sum);
  svd64_ortho_kernel<<<(unsigned)batch, 256, kOrthoSmem, s>>>(A, v0, g1, v1, batch);
  *G1 = g1;
  *V1 = v1;
  return cudaGetLastError();
}

}  // namespace nd4b
