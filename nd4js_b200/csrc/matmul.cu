// matmul.cu — broadcast-batched fp64 GEMM kernels (replaces matmul2_RR, nd4js src/la/matmul.js:31-74).
//
// Kernels:
//  * matmul32_kernel     : I=K=J=32, one warp per matrix triple, operands loaded straight from HBM
//                          into DMMA fragment registers with 16-byte loads.  HBM-bound
//                          (24 576 B and 65 536 flop per matrix, AI 2.67 flop/B).
//  * gemm_pipe_kernel    : large / mid-sized products; cp.async-pipelined CTA tiles, DMMA.8x8x4 from smem.
//  * gemm_tiled_kernel   : any I,K,J (odd sizes, unaligned operands); register-staged CTA tiles.
//  * matmul_small_kernel : tiny matrices (3x3, 4x4, ...): one lane per matrix through cp.async-staged smem.
//  * matmul_frag8_kernel : the rest up to 8x8: DMMA fragments straight from HBM.
// The reference accumulates each C_ij sequentially in k with separately rounded mul and add; the
// tensor-core path accumulates with fused multiply-adds in a different k order, so results agree
// componentwise to a few ulp of (|A||B|)_ij (tests/ state the bound), not bit for bit.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace nd4b {

__device__ __forceinline__ void decode_batch(const BatchMap& map, int64_t m_local, int64_t& ao, int64_t& bo) {
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

// ------------------------------------------------------------------------------------------------
// 32x32x32, warp per matrix.
//
// lane = 4g+t.  k is consumed in the order k(w,h,t) = 8w+2t+h so that thread t's two A values of
// k-steps (w,0),(w,1) are one 16-byte load; tile columns are n(x,e,g) = 16x+e+2g so that a thread's
// two B values of tiles (x,0),(x,1) are one 16-byte load and its four C values of (x,0),(x,1) are
// 32 contiguous bytes.  Permuting k and n inside an MMA is free (same permutation on both operands /
// on B and C).  Per warp instruction: A load 8 rows x 64 B, B load 4 rows x 128 B, C store 8 rows x 64 B.
// ------------------------------------------------------------------------------------------------
constexpr int kMM32Warps = 8;

__global__ void __launch_bounds__(kMM32Warps * 32)
matmul32_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                int64_t batch, BatchMap map) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t m = (int64_t)blockIdx.x * kMM32Warps + warp;
  if (m >= batch) return;
  int64_t ao, bo;
  decode_batch(map, m, ao, bo);
  const double* a = A + ao + g * 32 + 2 * t;
  const double* b = B + bo + (2 * t) * 32 + 2 * g;
  double* c = C + m * 1024 + g * 32 + 4 * t;

  double2 bf[4][2][2];
#pragma unroll
  for (int w = 0; w < 4; w++)
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int x = 0; x < 2; x++) bf[w][h][x] = ldg2(b + (8 * w + h) * 32 + 16 * x);

  double2 af[4][4];
#pragma unroll
  for (int rb = 0; rb < 4; rb++)
#pragma unroll
    for (int w = 0; w < 4; w++) af[rb][w] = ldg2(a + rb * 256 + 8 * w);

#pragma unroll
  for (int rb = 0; rb < 4; rb++) {
    double acc[2][2][2];
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
      for (int e = 0; e < 2; e++) acc[x][e][0] = acc[x][e][1] = 0.0;
#pragma unroll
    for (int w = 0; w < 4; w++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const double av = h ? af[rb][w].y : af[rb][w].x;
#pragma unroll
        for (int x = 0; x < 2; x++) {
          dmma884(acc[x][0][0], acc[x][0][1], av, bf[w][h][x].x);
          dmma884(acc[x][1][0], acc[x][1][1], av, bf[w][h][x].y);
        }
      }
#pragma unroll
    for (int x = 0; x < 2; x++) {
      double* p = c + rb * 256 + 16 * x;
      stg2_stream(p, acc[x][0][0], acc[x][1][0]);
      stg2_stream(p + 2, acc[x][0][1], acc[x][1][1]);
    }
  }
}

// The broadcast form of the same product ([batch,32,32] x [1,32,32]): every unit multiplies by the SAME B.  The kernel
// above would fetch B's fragments again for every matrix (8 KiB per unit through L1 / L2 beside the 8 KiB of A from HBM); here a
// warp loads them once and keeps them in registers for MPW consecutive matrices, the A fragments of the next matrix in flight
// while the current one is multiplied.  Same fragment layout, same accumulation order: bit-identical results.
template <int MPW, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
matmul32_bcast_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                      int64_t batch, BatchMap map) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t m0 = ((int64_t)blockIdx.x * WARPS + warp) * MPW;
  if (m0 >= batch) return;
  int64_t ao, bo;
  decode_batch(map, m0, ao, bo);
  const double* b = B + bo + (2 * t) * 32 + 2 * g;
  double2 bf[4][2][2];
#pragma unroll
  for (int w = 0; w < 4; w++)
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int x = 0; x < 2; x++) bf[w][h][x] = ldg2(b + (8 * w + h) * 32 + 16 * x);

  // the row block just multiplied is refilled from the next matrix at once: sixteen 16-byte loads per lane stay in flight
  // without a second set of registers
  double2 af[4][4];
  const double* a = A + m0 * map.a_lin + g * 32 + 2 * t;
#pragma unroll
  for (int rb = 0; rb < 4; rb++)
#pragma unroll
    for (int w = 0; w < 4; w++) af[rb][w] = ldg2_stream(a + rb * 256 + 8 * w);
#pragma unroll
  for (int i = 0; i < MPW; i++) {
    const int64_t m = m0 + i;
    if (m >= batch) break;   // warp-uniform
    const bool more = (i + 1 < MPW) && (m + 1 < batch);
    double* c = C + m * 1024 + g * 32 + 4 * t;
#pragma unroll
    for (int rb = 0; rb < 4; rb++) {
      double acc[2][2][2];
#pragma unroll
      for (int x = 0; x < 2; x++)
#pragma unroll
        for (int e = 0; e < 2; e++) acc[x][e][0] = acc[x][e][1] = 0.0;
#pragma unroll
      for (int w = 0; w < 4; w++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const double av = h ? af[rb][w].y : af[rb][w].x;
#pragma unroll
          for (int x = 0; x < 2; x++) {
            dmma884(acc[x][0][0], acc[x][0][1], av, bf[w][h][x].x);
            dmma884(acc[x][1][0], acc[x][1][1], av, bf[w][h][x].y);
          }
        }
      if (more) {
#pragma unroll
        for (int w = 0; w < 4; w++) af[rb][w] = ldg2_stream(a + (i + 1) * map.a_lin + rb * 256 + 8 * w);
      }
#pragma unroll
      for (int x = 0; x < 2; x++) {
        double* p = c + rb * 256 + 16 * x;
        stg2_stream(p, acc[x][0][0], acc[x][1][0]);
        stg2_stream(p + 2, acc[x][0][1], acc[x][1][1]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// General tiled GEMM.  CTA = WR x WC warps, each warp owns a (8*TM) x (8*TN) block of C.
// smem tiles are padded so that DMMA fragment reads (8 rows x 4 k, resp. 4 k x 8 cols of doubles)
// hit 16 distinct 8-byte bank pairs per half warp:  row stride == 4 (mod 16) doubles.
// ------------------------------------------------------------------------------------------------
template <int WR, int WC, int TM, int TN>
struct TileCfg {
  static constexpr int BM = WR * TM * 8, BN = WC * TN * 8, BK = 16;
  static constexpr int LDA = BK + 4;  // 20 doubles
  static constexpr int LDB = BN + 4;  // == 4 mod 16 because BN is a multiple of 16
  static constexpr int THREADS = WR * WC * 32;
  static constexpr int A_ELEMS = BM * BK, B_ELEMS = BK * BN;
  static constexpr int A_PER_THR = A_ELEMS / THREADS, B_PER_THR = B_ELEMS / THREADS;
  static_assert(BN % 16 == 0, "BN must be a multiple of 16");
  static_assert(A_ELEMS % (2 * THREADS) == 0 && B_ELEMS % (2 * THREADS) == 0, "tile/threads mismatch");
};

template <int WR, int WC, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(WR * WC * 32)
gemm_tiled_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                  int64_t batch, int I, int K, int J, BatchMap map, int tiles_m, int tiles_n) {
  using Cfg = TileCfg<WR, WC, TM, TN>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, LDA = Cfg::LDA, LDB = Cfg::LDB;
  constexpr int T = Cfg::THREADS;
  __shared__ __align__(16) double As[2][BM * LDA];
  __shared__ __align__(16) double Bs[2][BK * LDB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp / WC, wc = warp % WC;

  int64_t tile = blockIdx.x;
  const int tn = (int)(tile % tiles_n); tile /= tiles_n;
  const int tm = (int)(tile % tiles_m); tile /= tiles_m;
  const int64_t m = tile;
  if (m >= batch) return;
  int64_t ao, bo;
  decode_batch(map, m, ao, bo);
  const double* a = A + ao;
  const double* b = B + bo;
  double* c = C + m * (int64_t)I * J;
  const int row0 = tm * BM, col0 = tn * BN;

  // global -> register staging.  VEC: 16-byte loads (requires K,J even and 16-byte aligned bases).
  constexpr int AV = VEC ? Cfg::A_PER_THR / 2 : Cfg::A_PER_THR;
  constexpr int BV = VEC ? Cfg::B_PER_THR / 2 : Cfg::B_PER_THR;
  double ra[Cfg::A_PER_THR], rb[Cfg::B_PER_THR];

  auto load_tiles = [&](int k0) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < AV; i++) {
        const int e = (tid + i * T) * 2, r = e / BK, kk = e % BK;
        const int gr = row0 + r, gk = k0 + kk;
        double2 v = make_double2(0.0, 0.0);
        if (gr < I && gk < K) v = ldg2(a + (int64_t)gr * K + gk);  // K even => gk+1 < K
        ra[2 * i] = v.x; ra[2 * i + 1] = v.y;
      }
#pragma unroll
      for (int i = 0; i < BV; i++) {
        const int e = (tid + i * T) * 2, kk = e / BN, cc = e % BN;
        const int gk = k0 + kk, gc = col0 + cc;
        double2 v = make_double2(0.0, 0.0);
        if (gk < K && gc < J) v = ldg2(b + (int64_t)gk * J + gc);
        rb[2 * i] = v.x; rb[2 * i + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < AV; i++) {
        const int e = tid + i * T, r = e / BK, kk = e % BK;
        const int gr = row0 + r, gk = k0 + kk;
        ra[i] = (gr < I && gk < K) ? __ldg(a + (int64_t)gr * K + gk) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < BV; i++) {
        const int e = tid + i * T, kk = e / BN, cc = e % BN;
        const int gk = k0 + kk, gc = col0 + cc;
        rb[i] = (gk < K && gc < J) ? __ldg(b + (int64_t)gk * J + gc) : 0.0;
      }
    }
  };
  auto store_tiles = [&](int buf) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < AV; i++) {
        const int e = (tid + i * T) * 2, r = e / BK, kk = e % BK;
        *reinterpret_cast<double2*>(&As[buf][r * LDA + kk]) = make_double2(ra[2 * i], ra[2 * i + 1]);
      }
#pragma unroll
      for (int i = 0; i < BV; i++) {
        const int e = (tid + i * T) * 2, kk = e / BN, cc = e % BN;
        *reinterpret_cast<double2*>(&Bs[buf][kk * LDB + cc]) = make_double2(rb[2 * i], rb[2 * i + 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < AV; i++) {
        const int e = tid + i * T, r = e / BK, kk = e % BK;
        As[buf][r * LDA + kk] = ra[i];
      }
#pragma unroll
      for (int i = 0; i < BV; i++) {
        const int e = tid + i * T, kk = e / BN, cc = e % BN;
        Bs[buf][kk * LDB + cc] = rb[i];
      }
    }
  };

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
    const double* as = &As[buf][(wr * TM * 8 + g) * LDA + t];
    const double* bs = &Bs[buf][t * LDB + wc * TN * 8 + g];
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      double af[TM], bfr[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) af[i] = as[i * 8 * LDA + ks * 4];
#pragma unroll
      for (int j = 0; j < TN; j++) bfr[j] = bs[ks * 4 * LDB + j * 8];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bfr[j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: thread holds C[row g of tile i][cols 2t,2t+1 of tile j]
#pragma unroll
  for (int i = 0; i < TM; i++) {
    const int gr = row0 + (wr * TM + i) * 8 + g;
    if (gr >= I) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      const int gc = col0 + (wc * TN + j) * 8 + 2 * t;
      double* p = c + (int64_t)gr * J + gc;
      if (VEC) {
        if (gc < J) *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else {
        if (gc < J) p[0] = acc[i][j][0];
        if (gc + 1 < J) p[1] = acc[i][j][1];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pipelined tiled GEMM for 16-byte-aligned operands with even K and J: the same CTA tiling and padded
// smem layout as gemm_tiled_kernel, but tiles are brought in by cp.async (LDGSTS, 16 bytes, zero-fill past
// the edges) through a STAGES-deep ring, so global latency is hidden behind the DMMA stream instead of being
// exposed at every k-tile.  This is the kernel behind the compute-bound figures (C1 and the large-N probes).
// ------------------------------------------------------------------------------------------------
template <int WR, int WC, int TM, int TN, int STAGES, int BK_ = 16, int KS = 1>
struct PipeCfg {
  static constexpr int BM = WR * TM * 8, BN = WC * TN * 8, BK = BK_;
  static constexpr int LDA = BK + 4, LDB = BN + 4;
  static constexpr int THREADS = WR * WC * KS * 32;   // KS warp groups share a tile: each takes 1/KS of every k-tile (split-K inside the CTA)
  static_assert((BK / 4) % KS == 0, "k-steps per tile must divide among the warp groups");
  static constexpr int A_CHUNKS = BM * BK / 2, B_CHUNKS = BK * BN / 2;  // 16-byte chunks per tile
  static constexpr int STAGE_DOUBLES = BM * LDA + BK * LDB;
  static constexpr size_t SMEM = sizeof(double) * STAGES * STAGE_DOUBLES;
  static_assert(BN % 16 == 0, "BN must be a multiple of 16");
  static_assert(A_CHUNKS % THREADS == 0 && B_CHUNKS % THREADS == 0, "tile/threads mismatch");
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

template <int WR, int WC, int TM, int TN, int STAGES, int BK_ = 16, int KS = 1>
__global__ void __launch_bounds__(WR * WC * KS * 32)
gemm_pipe_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                 int64_t batch, int I, int K, int J, BatchMap map, int tiles_m, int tiles_n) {
  using Cfg = PipeCfg<WR, WC, TM, TN, STAGES, BK_, KS>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, LDA = Cfg::LDA, LDB = Cfg::LDB, T = Cfg::THREADS;
  extern __shared__ __align__(16) double gemm_smem[];

  const int tid = threadIdx.x, lane = tid & 31, warp_all = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int kgrp = warp_all / (WR * WC), warp = warp_all % (WR * WC);   // warp group kgrp takes k-steps kgrp*(BK/4/KS) .. of every tile
  const int wr = warp / WC, wc = warp % WC;

  int64_t tile = blockIdx.x;
  const int tn = (int)(tile % tiles_n); tile /= tiles_n;
  const int tm = (int)(tile % tiles_m); tile /= tiles_m;
  const int64_t m = tile;
  if (m >= batch) return;
  int64_t ao, bo;
  decode_batch(map, m, ao, bo);
  const double* a = A + ao;
  const double* b = B + bo;
  double* c = C + m * (int64_t)I * J;
  const int row0 = tm * BM, col0 = tn * BN;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(gemm_smem);

  auto issue_tile = [&](int kt, int stage) {
    const int k0 = kt * BK;
    const uint32_t as = smem_base + (uint32_t)(stage * Cfg::STAGE_DOUBLES) * 8u;
    const uint32_t bs = as + (uint32_t)(BM * LDA) * 8u;
#pragma unroll
    for (int i = 0; i < Cfg::A_CHUNKS / T; i++) {
      const int ch = tid + i * T, r = ch / (BK / 2), kk = (ch % (BK / 2)) * 2;
      const int gr = row0 + r, gk = k0 + kk;
      const bool ok = gr < I && gk < K;
      cp_async16(as + (uint32_t)(r * LDA + kk) * 8u, ok ? a + (int64_t)gr * K + gk : a, ok);
    }
#pragma unroll
    for (int i = 0; i < Cfg::B_CHUNKS / T; i++) {
      const int ch = tid + i * T, kk = ch / (BN / 2), cc = (ch % (BN / 2)) * 2;
      const int gk = k0 + kk, gc = col0 + cc;
      const bool ok = gk < K && gc < J;
      cp_async16(bs + (uint32_t)(kk * LDB + cc) * 8u, ok ? b + (int64_t)gk * J + gc : b, ok);
    }
  };

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (K + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) issue_tile(s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < nk; kt++) {
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
    __syncthreads();  // tile kt has landed for everybody; everybody is done with the stage refilled below
    if (kt + STAGES - 1 < nk) issue_tile(kt + STAGES - 1, (kt + STAGES - 1) % STAGES);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double* as = gemm_smem + (kt % STAGES) * Cfg::STAGE_DOUBLES + (wr * TM * 8 + g) * LDA + t;
    const double* bs = gemm_smem + (kt % STAGES) * Cfg::STAGE_DOUBLES + BM * LDA + t * LDB + wc * TN * 8 + g;
#pragma unroll
    for (int kq = 0; kq < BK / 4 / KS; kq++) {
      const int ks = kgrp * (BK / 4 / KS) + kq;
      double af[TM], bfr[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) af[i] = as[i * 8 * LDA + ks * 4];
#pragma unroll
      for (int j = 0; j < TN; j++) bfr[j] = bs[ks * 4 * LDB + j * 8];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bfr[j]);
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (KS > 1) {   // the groups' partial tiles meet in shared memory (the ring is idle now); group 0 stores
    __syncthreads();
    double* red = gemm_smem + (size_t)(warp * 32 + lane) * (TM * TN * 2);
    for (int gq = KS - 1; gq > 0; gq--) {
      if (kgrp == gq) {
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
          for (int j = 0; j < TN; j++) *reinterpret_cast<double2*>(red + (i * TN + j) * 2) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
      __syncthreads();
      if (kgrp == 0) {
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
          for (int j = 0; j < TN; j++) {
            const double2 p = *reinterpret_cast<const double2*>(red + (i * TN + j) * 2);
            acc[i][j][0] += p.x;
            acc[i][j][1] += p.y;
          }
      }
      if (gq > 1) __syncthreads();
    }
    if (kgrp != 0) return;
  }

#pragma unroll
  for (int i = 0; i < TM; i++) {
    const int gr = row0 + (wr * TM + i) * 8 + g;
    if (gr >= I) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      const int gc = col0 + (wc * TN + j) * 8 + 2 * t;
      if (gc < J) *reinterpret_cast<double2*>(c + (int64_t)gr * J + gc) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

template <int WR, int WC, int TM, int TN, int STAGES, int BK_ = 16, int KS = 1>
static cudaError_t launch_pipe(cudaStream_t s, const double* A, const double* B, double* C,
                               int64_t batch, int I, int K, int J, const BatchMap& map) {
  using Cfg = PipeCfg<WR, WC, TM, TN, STAGES, BK_, KS>;
  const int tiles_m = (I + Cfg::BM - 1) / Cfg::BM, tiles_n = (J + Cfg::BN - 1) / Cfg::BN;
  const int64_t grid = batch * tiles_m * tiles_n;
  if (grid <= 0 || grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  auto kern = gemm_pipe_kernel<WR, WC, TM, TN, STAGES, BK_, KS>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  kern<<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, s>>>(A, B, C, batch, I, K, J, map, tiles_m, tiles_n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// The same pipelined GEMM fed by the TMA engine: a producer warp issues bulk asynchronous copies (cp.async.bulk, SASS
// UBLKCP: one per tile row, 16-byte aligned, landing in the same padded, bank-conflict-free rows the DMMA fragments are
// read from) and signals a per-stage "full" mbarrier through the copies' transaction bytes; the consumer warps wait on
// it, run the DMMA k-steps of the stage and arrive on the stage's "empty" mbarrier, which the producer waits on before it
// refills the stage.  No consumer thread issues a load instruction or a CTA-wide barrier inside the main loop.  Tensor-map
// copies (cp.async.bulk.tensor) would land the tile densely: with 8-byte fragment reads neither the dense layout nor the
// hardware swizzles are conflict-free for the B operand, which is why the rows are copied one by one into padded rows.
// For full tiles only (I % BM == 0, J % BN == 0, K % BK == 0: C1, the 4096^3 probe); other shapes take gemm_pipe_kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mb, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t mb, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mb) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nGB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra GB_WAIT;\n}" ::"r"(mb), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mb) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
}

template <int WR, int WC, int TM, int TN, int STAGES, int BK_ = 16>
__global__ void __launch_bounds__(WR * WC * 32 + 32)
gemm_bulk_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                 int64_t batch, int I, int K, int J, BatchMap map, int tiles_m, int tiles_n) {
  using Cfg = PipeCfg<WR, WC, TM, TN, STAGES, BK_>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, LDA = Cfg::LDA, LDB = Cfg::LDB, NCW = WR * WC;
  extern __shared__ __align__(16) double gemm_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  int64_t tile = blockIdx.x;
  const int tn = (int)(tile % tiles_n); tile /= tiles_n;
  const int tm = (int)(tile % tiles_m); tile /= tiles_m;
  const int64_t m = tile;
  if (m >= batch) return;
  int64_t ao, bo;
  decode_batch(map, m, ao, bo);
  const int row0 = tm * BM, col0 = tn * BN;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(gemm_smem);
  const uint32_t bar_base = smem_base + (uint32_t)(STAGES * Cfg::STAGE_DOUBLES) * 8u;   // full[STAGES], empty[STAGES]
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < STAGES; st++) { mbar_init(bar_base + 8u * st, 1); mbar_init(bar_base + 8u * (STAGES + st), NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nk = K / BK;

  if (warp == NCW) {
    // ---- producer warp: one bulk copy per tile row ----
    const double* a = A + ao + (int64_t)row0 * K;
    const double* b = B + bo + col0;
    for (int kt = 0; kt < nk; kt++) {
      const int st = kt % STAGES;
      if (kt >= STAGES) mbar_wait(bar_base + 8u * (STAGES + st), ((kt / STAGES) - 1) & 1);   // the consumers have drained the stage
      const uint32_t full = bar_base + 8u * st;
      if (lane == 0) mbar_expect_tx(full, (uint32_t)((BM * BK + BK * BN) * 8));
      __syncwarp();
      const uint32_t as = smem_base + (uint32_t)(st * Cfg::STAGE_DOUBLES) * 8u, bs = as + (uint32_t)(BM * LDA) * 8u;
      const int k0 = kt * BK;
      for (int r = lane; r < BM; r += 32) bulk_g2s(as + (uint32_t)(r * LDA) * 8u, a + (int64_t)r * K + k0, BK * 8, full);
      for (int r = lane; r < BK; r += 32) bulk_g2s(bs + (uint32_t)(r * LDB) * 8u, b + (int64_t)(k0 + r) * J, BN * 8, full);
    }
    return;
  }

  // ---- consumer warps ----
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp / WC, wc = warp % WC;
  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int kt = 0; kt < nk; kt++) {
    const int st = kt % STAGES;
    mbar_wait(bar_base + 8u * st, (kt / STAGES) & 1);
    const double* as = gemm_smem + st * Cfg::STAGE_DOUBLES + (wr * TM * 8 + g) * LDA + t;
    const double* bs = gemm_smem + st * Cfg::STAGE_DOUBLES + BM * LDA + t * LDB + wc * TN * 8 + g;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      double af[TM], bfr[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) af[i] = as[i * 8 * LDA + ks * 4];
#pragma unroll
      for (int j = 0; j < TN; j++) bfr[j] = bs[ks * 4 * LDB + j * 8];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bfr[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_base + 8u * (STAGES + st));   // this warp is done reading the stage
  }
  double* c = C + m * (int64_t)I * J;
#pragma unroll
  for (int i = 0; i < TM; i++) {
    const int gr = row0 + (wr * TM + i) * 8 + g;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      const int gc = col0 + (wc * TN + j) * 8 + 2 * t;
      *reinterpret_cast<double2*>(c + (int64_t)gr * J + gc) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

static int gemm_bulk_mode() {
  static int on = -1;
  if (on < 0) {
    const char* ev = getenv("ND4B_GEMM_TMA");
    on = ev ? atoi(ev) : 1;   // 0: cp.async (LDGSTS) tiles for every shape; 1: bulk copies for 128x128 tiles; 2: for every full-tile shape
  }
  return on;
}

template <int WR, int WC, int TM, int TN, int STAGES, int BK_ = 16>
static cudaError_t launch_bulk_or_pipe(cudaStream_t s, const double* A, const double* B, double* C,
                                       int64_t batch, int I, int K, int J, const BatchMap& map) {
  using Cfg = PipeCfg<WR, WC, TM, TN, STAGES, BK_>;
  // Measured on B200 (tools/gemm_sweep.py, ND4B_GEMM_TMA=0/1): with 128x128x32 tiles (A rows of 256 B, B rows of 1 KiB) the
  // bulk-copy pipeline equals the cp.async one (4096^3: 31.7 vs 32.2 TFLOP/s); with the small tiles of launch-sized
  // products the per-row copies are 128 B each and the copy engine's latency per request dominates (512^3: 47 vs 17 us),
  // so those keep cp.async unless ND4B_GEMM_TMA=2 forces the bulk path.
  const int mode = gemm_bulk_mode();
  if (mode == 0 || (mode == 1 && Cfg::BM < 128) || I % Cfg::BM || J % Cfg::BN || K % Cfg::BK)
    return launch_pipe<WR, WC, TM, TN, STAGES, BK_>(s, A, B, C, batch, I, K, J, map);
  const int tiles_m = I / Cfg::BM, tiles_n = J / Cfg::BN;
  const int64_t grid = batch * tiles_m * tiles_n;
  if (grid <= 0 || grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  auto kern = gemm_bulk_kernel<WR, WC, TM, TN, STAGES, BK_>;
  constexpr size_t smem = Cfg::SMEM + 2 * STAGES * 8;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  kern<<<(unsigned)grid, Cfg::THREADS + 32, smem, s>>>(A, B, C, batch, I, K, J, map, tiles_m, tiles_n);
  return cudaGetLastError();
}

template <int WR, int WC, int TM, int TN>
static cudaError_t launch_tiled(cudaStream_t s, const double* A, const double* B, double* C,
                                int64_t batch, int I, int K, int J, const BatchMap& map, bool vec) {
  using Cfg = TileCfg<WR, WC, TM, TN>;
  const int tiles_m = (I + Cfg::BM - 1) / Cfg::BM, tiles_n = (J + Cfg::BN - 1) / Cfg::BN;
  const int64_t grid = batch * tiles_m * tiles_n;
  if (grid <= 0 || grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  if (vec)
    gemm_tiled_kernel<WR, WC, TM, TN, true><<<(unsigned)grid, Cfg::THREADS, 0, s>>>(A, B, C, batch, I, K, J, map, tiles_m, tiles_n);
  else
    gemm_tiled_kernel<WR, WC, TM, TN, false><<<(unsigned)grid, Cfg::THREADS, 0, s>>>(A, B, C, batch, I, K, J, map, tiles_m, tiles_n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Tiny matrices (I, K, J <= 8 and I*K + K*J + I*J <= 100: 3x3, 4x4, 6x6 * 6x1, ...): HBM-bound by a wide margin (a 4x4
// product moves 384 B for 128 flop), and a DMMA tile would be mostly padding.  One lane per matrix, 32 matrices per warp:
// the warp stages its 32 A and B operands through shared memory with coalesced 8-byte loads (a contiguous batch is one
// contiguous block; broadcast operands are addressed through the odometer, lane q holding matrix q's offsets), every
// lane then forms its product from its own padded slots (odd stride: conflict-free), C_ij accumulated with FMAs over k
// ascending, and the results leave through shared memory as coalesced stores again.
// ------------------------------------------------------------------------------------------------
constexpr int kSmallWarps = 4;

__host__ __device__ inline int small_stride(int n) { return n | 1; }   // odd number of doubles per matrix slot

// Stages `cnt` doubles of each of the warp's nmat matrices (matrix q starts at base + off_q, off_q held by lane q) into
// padded slots with 8-byte asynchronous copies: nothing waits until the caller's cp.async.wait_all, so every load of the
// warp's 32 operands is in flight at once (a load -> store loop through registers waits for HBM once per iteration).
__device__ __forceinline__ void small_stage(const double* __restrict__ base, int64_t my_off, int cnt, int stride,
                                            double* slots, int nmat, int lane) {
  const float inv = 1.0f / (float)cnt;
  const int total = nmat * cnt;
  const uint32_t slots_s = (uint32_t)__cvta_generic_to_shared(slots);
  for (int g0 = 0; g0 < total; g0 += 32) {         // warp-uniform trip count: every lane takes part in the shuffle
    const int g = g0 + lane;
    const bool in = g < total;
    const int q = in ? (int)(((float)g + 0.5f) * inv) : 0, e = g - q * cnt;   // exact: g < 2048
    const int64_t off = __shfl_sync(kFull, my_off, q);
    if (in) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(slots_s + (uint32_t)(q * stride + e) * 8u), "l"(base + off + e) : "memory");
  }
}

// CI, CK, CJ > 0: compile-time dimensions (3x3, 4x4 products and their matrix-vector forms): the lane's product is fully
// unrolled from registers; 0: run-time dimensions.
template <int CI, int CK, int CJ>
__global__ void __launch_bounds__(kSmallWarps * 32)
matmul_small_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                    int64_t batch, int I_, int K_, int J_, BatchMap map) {
  const int I = CI ? CI : I_, K = CK ? CK : K_, J = CJ ? CJ : J_;
  extern __shared__ __align__(16) double small_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ik = I * K, kj = K * J, ij = I * J;
  const int sa = small_stride(ik), sb = small_stride(kj), sc = small_stride(ij);
  double* as = small_smem + (size_t)warp * 32 * (sa + sb + sc);
  double* bs = as + 32 * sa;
  double* cs = bs + 32 * sb;
  const int64_t m0 = ((int64_t)blockIdx.x * kSmallWarps + warp) * 32;
  if (m0 >= batch) return;  // warp-uniform
  const int nmat = (int)min((int64_t)32, batch - m0);
  int64_t ao = 0, bo = 0;
  if (lane < nmat) decode_batch(map, m0 + lane, ao, bo);
  small_stage(A, ao, ik, sa, as, nmat, lane);
  small_stage(B, bo, kj, sb, bs, nmat, lane);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
  if (lane < nmat) {
    const double* a = as + lane * sa;
    const double* b = bs + lane * sb;
    double* c = cs + lane * sc;
    if (CI) {
      double br[(CK ? CK : 1) * (CJ ? CJ : 1)];
#pragma unroll
      for (int e = 0; e < CK * CJ; e++) br[e] = b[e];
#pragma unroll
      for (int i = 0; i < CI; i++) {
        double ar[CK ? CK : 1];
#pragma unroll
        for (int k = 0; k < CK; k++) ar[k] = a[i * CK + k];
#pragma unroll
        for (int j = 0; j < CJ; j++) {
          double acc = ar[0] * br[j];
#pragma unroll
          for (int k = 1; k < CK; k++) acc = fma(ar[k], br[k * CJ + j], acc);
          c[i * CJ + j] = acc;
        }
      }
    } else {
      for (int i = 0; i < I; i++)
        for (int j = 0; j < J; j++) {
          double acc = a[i * K] * b[j];
          for (int k = 1; k < K; k++) acc = fma(a[i * K + k], b[k * J + j], acc);
          c[i * J + j] = acc;
        }
    }
  }
  __syncwarp();
  double* dst = C + m0 * ij;
  const float inv_ij = 1.0f / (float)ij;
  for (int g = lane; g < nmat * ij; g += 32) {
    const int q = (int)(((float)g + 0.5f) * inv_ij), e = g - q * ij;
    dst[g] = cs[q * sc + e];
  }
}

template <int CI, int CK, int CJ>
static cudaError_t launch_matmul_small_t(cudaStream_t s, const double* A, const double* B, double* C,
                                         int64_t batch, int I, int K, int J, const BatchMap& map) {
  const size_t smem = sizeof(double) * kSmallWarps * 32 * (size_t)(small_stride(I * K) + small_stride(K * J) + small_stride(I * J));
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(matmul_small_kernel<CI, CK, CJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int64_t grid = (batch + kSmallWarps * 32 - 1) / (kSmallWarps * 32);
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  matmul_small_kernel<CI, CK, CJ><<<(unsigned)grid, kSmallWarps * 32, smem, s>>>(A, B, C, batch, I, K, J, map);
  return cudaGetLastError();
}

static cudaError_t launch_matmul_small(cudaStream_t s, const double* A, const double* B, double* C,
                                       int64_t batch, int I, int K, int J, const BatchMap& map) {
  if (I == 4 && K == 4 && J == 4) return launch_matmul_small_t<4, 4, 4>(s, A, B, C, batch, I, K, J, map);
  if (I == 3 && K == 3 && J == 3) return launch_matmul_small_t<3, 3, 3>(s, A, B, C, batch, I, K, J, map);
  if (I == 2 && K == 2 && J == 2) return launch_matmul_small_t<2, 2, 2>(s, A, B, C, batch, I, K, J, map);
  if (I == 5 && K == 5 && J == 5) return launch_matmul_small_t<5, 5, 5>(s, A, B, C, batch, I, K, J, map);
  if (I == 4 && K == 4 && J == 1) return launch_matmul_small_t<4, 4, 1>(s, A, B, C, batch, I, K, J, map);
  if (I == 3 && K == 3 && J == 1) return launch_matmul_small_t<3, 3, 1>(s, A, B, C, batch, I, K, J, map);
  return launch_matmul_small_t<0, 0, 0>(s, A, B, C, batch, I, K, J, map);
}

// ------------------------------------------------------------------------------------------------
// I, K, J <= 8 beyond the one-lane-per-matrix kernel (e.g. 8x8 . 8x8): one DMMA.8x8x4 tile per matrix and k-half, the
// fragments loaded straight from HBM (lane 4g+t: A[g][t], A[g][t+4], B[t][g], B[t+4][g]; out-of-range entries are zeros)
// and C stored from the accumulator fragment (C[g][2t], C[g][2t+1]).  A warp walks over its matrices kFragUnroll at a
// time so that kFragUnroll x 4 loads per lane are in flight; for 8x8 operands every 32-byte sector fetched is used.
// ------------------------------------------------------------------------------------------------
constexpr int kFragWarps = 8, kFragUnroll = 8;

__global__ void __launch_bounds__(kFragWarps * 32)
matmul_frag8_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                    int64_t batch, int I, int K, int J, BatchMap map) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t m0 = ((int64_t)blockIdx.x * kFragWarps + warp) * kFragUnroll;
  if (m0 >= batch) return;  // warp-uniform
  const bool a0_in = g < I && t < K, a1_in = g < I && t + 4 < K;
  const bool b0_in = t < K && g < J, b1_in = t + 4 < K && g < J;
  double a0[kFragUnroll], a1[kFragUnroll], b0[kFragUnroll], b1[kFragUnroll];
#pragma unroll
  for (int u = 0; u < kFragUnroll; u++) {
    a0[u] = a1[u] = b0[u] = b1[u] = 0.0;
    const int64_t m = m0 + u;
    if (m < batch) {
      int64_t ao, bo;
      decode_batch(map, m, ao, bo);
      const double* a = A + ao;
      const double* b = B + bo;
      if (a0_in) a0[u] = ldg1_stream(a + g * K + t);
      if (a1_in) a1[u] = ldg1_stream(a + g * K + t + 4);
      if (b0_in) b0[u] = ldg1_stream(b + t * J + g);
      if (b1_in) b1[u] = ldg1_stream(b + (t + 4) * J + g);
    }
  }
#pragma unroll
  for (int u = 0; u < kFragUnroll; u++) {
    const int64_t m = m0 + u;
    if (m < batch) {   // warp-uniform
      double c0 = 0.0, c1 = 0.0;
      dmma884(c0, c1, a0[u], b0[u]);
      if (K > 4) dmma884(c0, c1, a1[u], b1[u]);
      double* c = C + m * (int64_t)(I * J) + g * J + 2 * t;
      if (g < I && 2 * t < J) c[0] = c0;
      if (g < I && 2 * t + 1 < J) c[1] = c1;
    }
  }
}

static cudaError_t launch_matmul_frag8(cudaStream_t s, const double* A, const double* B, double* C,
                                       int64_t batch, int I, int K, int J, const BatchMap& map) {
  const int per_cta = kFragWarps * kFragUnroll;
  const int64_t grid = (batch + per_cta - 1) / per_cta;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  matmul_frag8_kernel<<<(unsigned)grid, kFragWarps * 32, 0, s>>>(A, B, C, batch, I, K, J, map);
  return cudaGetLastError();
}

// The same idea for odd or unaligned operands up to 8*TI x 8*TK x 8*TJ (9x9, 17x17, ...), which the tiled kernels can
// only serve through their scalar path on tiles of 16 or 32: TI x TJ accumulator tiles per warp, every fragment an
// 8-byte load predicated on the matrix bounds (zeros outside), two matrices per warp in flight.
template <int TI, int TK, int TJ>
__global__ void __launch_bounds__(kFragWarps * 32)
matmul_frag_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                   int64_t batch, int I, int K, int J, BatchMap map) {
  constexpr int U = 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t m0 = ((int64_t)blockIdx.x * kFragWarps + warp) * U;
  if (m0 >= batch) return;  // warp-uniform
  double af[U][TI][2 * TK], bf[U][2 * TK][TJ];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int64_t m = m0 + u;
    int64_t ao = 0, bo = 0;
    const bool live = m < batch;
    if (live) decode_batch(map, m, ao, bo);
    const double* a = A + ao;
    const double* b = B + bo;
#pragma unroll
    for (int i = 0; i < TI; i++)
#pragma unroll
      for (int k = 0; k < 2 * TK; k++) {
        const int row = 8 * i + g, kk = 4 * k + t;
        af[u][i][k] = (live && row < I && kk < K) ? ldg1_stream(a + row * K + kk) : 0.0;
      }
#pragma unroll
    for (int k = 0; k < 2 * TK; k++)
#pragma unroll
      for (int j = 0; j < TJ; j++) {
        const int kk = 4 * k + t, col = 8 * j + g;
        bf[u][k][j] = (live && kk < K && col < J) ? ldg1_stream(b + kk * J + col) : 0.0;
      }
  }
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int64_t m = m0 + u;
    if (m >= batch) break;   // warp-uniform
    double* c = C + m * (int64_t)(I * J);
#pragma unroll
    for (int i = 0; i < TI; i++)
#pragma unroll
      for (int j = 0; j < TJ; j++) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int k = 0; k < 2 * TK; k++) dmma884(c0, c1, af[u][i][k], bf[u][k][j]);
        const int row = 8 * i + g, col = 8 * j + 2 * t;
        if (row < I && col < J) c[row * J + col] = c0;
        if (row < I && col + 1 < J) c[row * J + col + 1] = c1;
      }
  }
}

template <int TI, int TK, int TJ>
static cudaError_t launch_matmul_frag(cudaStream_t s, const double* A, const double* B, double* C,
                                      int64_t batch, int I, int K, int J, const BatchMap& map) {
  const int per_cta = kFragWarps * 2;
  const int64_t grid = (batch + per_cta - 1) / per_cta;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  matmul_frag_kernel<TI, TK, TJ><<<(unsigned)grid, kFragWarps * 32, 0, s>>>(A, B, C, batch, I, K, J, map);
  return cudaGetLastError();
}

cudaError_t launch_matmul(cudaStream_t s, const double* A, const double* B, double* C,
                          int64_t batch, int I, int K, int J, const BatchMap& map, int sm_count) {
  if (batch <= 0) return cudaSuccess;
  const bool aligned = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
                         reinterpret_cast<uintptr_t>(C)) & 15) == 0;
  bool str_even = (map.a_lin < 0 || map.a_lin % 2 == 0) && (map.b_lin < 0 || map.b_lin % 2 == 0);
  for (int d = 0; d < map.nd; d++) str_even = str_even && (map.a_str[d] % 2 == 0) && (map.b_str[d] % 2 == 0);
  if (I == 32 && K == 32 && J == 32 && aligned && str_even) {
    const int64_t grid = (batch + kMM32Warps - 1) / kMM32Warps;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    // one B for every unit (and A walked linearly): B's fragments stay in registers over four matrices per warp
    static const bool bcast_on = [] { const char* e = getenv("ND4B_MM32_BCAST"); return !(e && e[0] == '0'); }();
    if (bcast_on && map.b_lin == 0 && map.a_lin == 1024 && batch >= 4096) {
      constexpr int MPW = 4, W = 4;
      const int64_t g4 = (batch + (int64_t)W * MPW - 1) / ((int64_t)W * MPW);
      matmul32_bcast_kernel<MPW, W, 3><<<(unsigned)g4, W * 32, 0, s>>>(A, B, C, batch, map);
      return cudaGetLastError();
    }
    matmul32_kernel<<<(unsigned)grid, kMM32Warps * 32, 0, s>>>(A, B, C, batch, map);
    return cudaGetLastError();
  }
  if (I <= 8 && K <= 8 && J <= 8 && batch >= 256) {
    if (I * K + K * J + I * J <= 100) return launch_matmul_small(s, A, B, C, batch, I, K, J, map);
    return launch_matmul_frag8(s, A, B, C, batch, I, K, J, map);
  }
  const bool vec = aligned && str_even && (K % 2 == 0) && (J % 2 == 0) && (((int64_t)I * J) % 2 == 0);
  if (!vec && batch >= 256) {   // odd / unaligned small operands: fragments straight from HBM instead of the scalar tile path
    if (I <= 16 && K <= 16 && J <= 16) return launch_matmul_frag<2, 2, 2>(s, A, B, C, batch, I, K, J, map);
    if (I <= 24 && K <= 24 && J <= 24) return launch_matmul_frag<3, 3, 3>(s, A, B, C, batch, I, K, J, map);
  }
  // Tile choice: 64x64 tiles when they still give >= 2 CTAs per SM, else 64x32 (4 warps of 16x32) to
  // spread a single mid-sized product (e.g. 512^3 -> 128 CTAs) over the 148 SMs, else 16x16 per warp.
  if (vec && K >= 64) {
    // pipelined kernels: 128x128 tiles when they fill the machine, else 64x64, else 64x32 (e.g. one 512^3 -> 128 CTAs)
    const int64_t t128 = batch * ((I + 127) / 128) * ((J + 127) / 128);
    const int64_t t64 = batch * ((I + 63) / 64) * ((J + 63) / 64);
    // BK = 32 halves the barriers per flop of the 128x128 configuration (4096^3: 32.2 vs 31.7 TFLOP/s)
    if (I >= 96 && J >= 96 && t128 >= sm_count) return launch_bulk_or_pipe<2, 4, 8, 4, 3, 32>(s, A, B, C, batch, I, K, J, map);
    // 64x64 tiles from 1.5 CTAs per SM on (1024^3 = 256 tiles: 80 us vs 101 us with 64x32 tiles)
    if (I >= 48 && J >= 48 && 2 * t64 >= 3LL * sm_count) return launch_bulk_or_pipe<2, 2, 4, 4, 4>(s, A, B, C, batch, I, K, J, map);
    if (I >= 48 && J >= 24) {
      // few tiles (e.g. one 512^3): 32x32 tiles with 4 warps of 16x16 put >= 2 CTAs on most SMs
      // measured on one 512^3: 64x32 tiles/4 warps 18.7 us, 64x32/8 warps 17.6 us, 32x32/4 warps (256 CTAs) 16.4 us;
      // a second sweep (BK 16/32/64, 2-4 stages, 32x32 / 64x32 / 32x64 tiles, 4 or 8 warps) stayed within 15.4-18.6 us:
      // at this size the time is fill + wave quantisation (108 SMs hold two CTAs, 40 hold one), not the tile shape.
      // Split-K over a 2- or 4-CTA cluster with a DSMEM reduction (64x64 tiles, 128 / 256 CTAs) was built and measured:
      // 18.9 / 20.9 us — a 64x64 CTA of 4 warps alone on an SM runs its main loop at ~52 % of the DMMA peak.
      const int64_t t6432 = batch * ((I + 63) / 64) * ((J + 31) / 32);
      if (t6432 < 2LL * sm_count) {
        // split-K inside the CTA: two (or four) warp groups share the 32x32 tile and each takes half (a quarter) of every k-tile,
        // so that a sub-partition holds 2-4 warps per resident CTA instead of one (ND4B_GEMM_KS=0|1|2|3 for A/B timing)
        static const int ks_mode = [] { const char* e = getenv("ND4B_GEMM_KS"); return e ? atoi(e) : 0; }();
        if (ks_mode == 1) return launch_pipe<2, 2, 2, 2, 4, 16, 2>(s, A, B, C, batch, I, K, J, map);
        if (ks_mode == 2) return launch_pipe<2, 2, 2, 2, 3, 32, 2>(s, A, B, C, batch, I, K, J, map);
        if (ks_mode == 3) return launch_pipe<2, 2, 2, 2, 3, 32, 4>(s, A, B, C, batch, I, K, J, map);
        return launch_bulk_or_pipe<2, 2, 2, 2, 4>(s, A, B, C, batch, I, K, J, map);
      }
      return launch_bulk_or_pipe<4, 1, 2, 4, 4>(s, A, B, C, batch, I, K, J, map);
    }
  }
  const int64_t t6464 = batch * ((I + 63) / 64) * ((J + 63) / 64);
  const int64_t t6432 = batch * ((I + 63) / 64) * ((J + 31) / 32);
  if (I >= 48 && J >= 48 && t6464 >= 2LL * sm_count) return launch_tiled<2, 2, 4, 4>(s, A, B, C, batch, I, K, J, map, vec);
  if (I >= 48 && J >= 24 && t6432 >= sm_count / 2) return launch_tiled<4, 1, 2, 4>(s, A, B, C, batch, I, K, J, map, vec);
  if (I > 16 || J > 16) return launch_tiled<2, 1, 2, 4>(s, A, B, C, batch, I, K, J, map, vec);
  return launch_tiled<1, 1, 2, 2>(s, A, B, C, batch, I, K, J, map, vec);
}

}  // namespace nd4b
