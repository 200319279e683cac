// K4 — the whole DepthBottleneckUni (yolov6/layers/common.py:898-927, deploy form) in ONE kernel:
//
//     y = SiLU( W2 * SiLU( DW_k( SiLU( W1 * x + b1 ) ) + bd ) + b2 )
//         x : [n,h,w,c_] fp16 NHWC (a channel slice of the RepHDW concat buffer)      c_  <= 64
//         t1: 3c_ = `mid` channels, NEVER in HBM (round 1: written by one GEMM, re-read by dwpw_kernel)  mid <= 192
//         y : [n,h,w,c_] (the next channel slice of the same buffer)
//
// Persistent, warp-specialised CTA, one per SM (it owns the SM's 512 TMEM columns and ~226 KB of shared memory);
// a CTA loops over 10 x 20-pixel output tiles, and per tile over 64-channel blocks `cb` of the mid channels:
//
//   warp 13 lane 0  TMA: W1 / W2 panels once (resident), then per tile the HALO tile of x — (10+k-1) x (20+k-1) pixels
//                   x 64 channels as ONE 4-D box, SWIZZLE_128B: rows = halo pixels, i.e. directly the K-major A operand
//                   of the expand GEMM (out-of-image pixels and channels >= c_ are zero-filled by the TMA unit)
//   warp 12 lane 0  tcgen05.mma issuer:  MMA1(cb): acc1[cb&1][3 M tiles x 64 cols] = Xhalo[384 x 64] * W1[cb]^T
//                                        MMA2(cb): acc2[2 M tiles x c_ cols]     += A2[256 x 64]    * W2[cb]^T
//   warps 8-11      epilogue 1: acc1 -> +b1 -> SiLU -> 0 for pixels outside the image (the depth-wise conv pads t1, not
//                   x) -> fp16 -> T1[cb&1] in shared memory (pixel rows of 128 B, 16-byte chunk ^ (row & 7): conflict
//                   free for one-row-per-thread writes);  epilogue 2 (once per tile): acc2 -> +b2 -> SiLU -> global
//   warps 0-7       depth-wise taps: warp = 5 x 5 pixel unit, lane = channel pair; 25 x k*k packed FFMA2 from T1, +bd,
//                   SiLU, fp16 pairs into the hand-swizzled SW128 A2 tile -> fence.proxy.async -> MMA2
//
// so that the three pipes the block needs run CONCURRENTLY: FFMA (taps), MUFU (the three SiLUs, mostly in the epilogue
// warps) and the tensor core; T1 and acc1 are double buffered (epilogue 1 of block cb+1 overlaps the taps of block cb).
// All barriers are mbarriers with explicit phase arithmetic on the running step s = tile_iteration * nblk + cb.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"

namespace mafb200 {

constexpr int kBnTX = 20, kBnTY = 10;      // output tile
constexpr int kBnCB = 64;                  // mid channels per block = one 128-byte swizzle row
constexpr int kBnTapWarps = 8, kBnEpiWarps = 4;
constexpr int kBnThreads = 32 * (kBnTapWarps + kBnEpiWarps + 2);  // + MMA warp + TMA warp = 448
constexpr int kBnMaxMid = 192, kBnMaxCin = 64;
constexpr int kBnAcc1Cols = 3 * kBnCB;     // 192: three M tiles of the halo
constexpr int kBnAcc2Col0 = 2 * kBnAcc1Cols;  // 384
constexpr int kBnA2Bytes = kBnTX * kBnTY * 128;  // 25600: 200 pixel rows (the MMA reads 256; the rest is never used)

struct BneckParams {
  CUtensorMap tm_x;    // x as {c_, W, H, N}, box {64, TW, TH, 1}, SWIZZLE_128B
  CUtensorMap tm_w1;   // W1 packed fp16 [mid_pad][64], box {64, mid_pad}, SWIZZLE_128B
  CUtensorMap tm_w1b;  // the same W1, box {64, 64}: one 64-channel block (2-CTA shape)
  CUtensorMap tm_w2;   // W2 packed fp16 [tile_n][mid_pad], box {64, tile_n}, SWIZZLE_128B
  const float* b1;     // [mid_pad]
  const float* dw_w;   // [k*k][mid_pad] fp32
  const float* dw_b;   // [mid_pad]
  const float* b2;     // [tile_n]
  __half* out;
  int32_t out_ld;
  int32_t H, W, N, tile_n, mid_pad, nblk;
  int32_t tiles_x, tiles_y, tiles;  // tiles = n * tiles_y * tiles_x
  uint32_t idesc1, idesc2;
  long long* trace;  // debug (mafb200_bottleneck_trace): clock64 stamps of CTA 0, [role][step][4]; nullptr = off
};

constexpr int kBnTraceSteps = 64;
#define BN_TRACE(role, s, slot)                                                                              \
  do {                                                                                                       \
    if (p.trace != nullptr && blockIdx.x == 0 && (s) < kBnTraceSteps && lane == 0)                           \
      p.trace[((role) * kBnTraceSteps + (s)) * 4 + (slot)] = clock64();                                      \
  } while (0)

__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}

template <int K>
__global__ void __launch_bounds__(kBnThreads, 1) bneck_kernel(const __grid_constant__ BneckParams p) {
  constexpr int P = K / 2;
  constexpr int TW = kBnTX + K - 1, TH = kBnTY + K - 1;
  constexpr int kHaloRows = TH * TW;
  constexpr uint32_t kXBytes = kHaloRows * 128;  // a multiple of 1024 for K = 3 (33792) and K = 5 (43008)
  static_assert(kXBytes % 1024 == 0 && kHaloRows <= 384, "halo tile must be whole swizzle atoms and fit 3 M tiles");

  extern __shared__ uint8_t smem_bn_raw[];
  uint8_t* smem = smem_bn_raw + ((1024u - (smem_u32(smem_bn_raw) & 1023u)) & 1023u);
  const int nblk = p.nblk, mid_pad = p.mid_pad;
  const int w2_blk = p.tile_n * 128;
  uint8_t* s_w1 = smem;                                   // [mid_pad rows][128 B]
  uint8_t* s_w2 = s_w1 + mid_pad * 128;                   // nblk x [tile_n rows][128 B]
  uint8_t* s_x = s_w2 + nblk * w2_blk;                    // [kHaloRows][128 B]; MMA1 reads 384 rows (tail = garbage rows)
  uint8_t* s_a2 = s_x + kXBytes;                          // [200 rows][128 B]; MMA2 reads 256 rows
  uint8_t* s_t1 = s_a2 + kBnA2Bytes;                      // 2 x [kHaloRows][128 B], chunk-swizzled by hand
  float* s_dw = reinterpret_cast<float*>(s_t1 + 2 * kXBytes);  // [K*K][mid_pad]
  float* s_b1 = s_dw + K * K * mid_pad;                   // [mid_pad]
  float* s_bd = s_b1 + mid_pad;                           // [mid_pad]
  float* s_b2 = s_bd + mid_pad;                           // [tile_n]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b2 + p.tile_n + (p.tile_n & 1));
  uint64_t* bar_w = bars;                // W1 + W2 landed
  uint64_t* bar_x_full = bars + 1;       // halo tile landed (phase per tile)
  uint64_t* bar_x_empty = bars + 2;      // last MMA1 of the tile has read it
  uint64_t* bar_acc1_full = bars + 3;    // [2] MMA1(s) complete
  uint64_t* bar_acc1_empty = bars + 5;   // [2] epilogue 1 has drained acc1[s & 1]        (4 warps)
  uint64_t* bar_t1_full = bars + 7;      // [2] epilogue 1 has written T1[s & 1]          (4 warps)
  uint64_t* bar_t1_empty = bars + 9;     // [2] the tap warps have read T1[s & 1]         (8 warps)
  uint64_t* bar_a2_full = bars + 11;     // the tap warps have written A2                 (8 warps)
  uint64_t* bar_a2_empty = bars + 12;    // MMA2(s) complete
  uint64_t* bar_acc2_full = bars + 13;   // last MMA2 of the tile complete
  uint64_t* bar_acc2_empty = bars + 14;  // epilogue 2 has drained acc2                   (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = (p.tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int total_steps = my_tiles * nblk;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_x_full, 1);
    mbar_init(bar_x_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_acc1_full[i], 1);
      mbar_init(&bar_acc1_empty[i], kBnEpiWarps);
      mbar_init(&bar_t1_full[i], kBnEpiWarps);
      mbar_init(&bar_t1_empty[i], kBnTapWarps);
    }
    mbar_init(bar_a2_full, kBnTapWarps);
    mbar_init(bar_a2_empty, 1);
    mbar_init(bar_acc2_full, 1);
    mbar_init(bar_acc2_empty, kBnEpiWarps);
    fence_barrier_init();
  }
  if (warp == kBnTapWarps + kBnEpiWarps) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // constants of the whole kernel (weights of the model): read before griddepcontrol.wait
  for (int i = threadIdx.x; i < K * K * mid_pad; i += kBnThreads) s_dw[i] = __ldg(p.dw_w + i);
  for (int i = threadIdx.x; i < mid_pad; i += kBnThreads) {
    s_b1[i] = __ldg(p.b1 + i);
    s_bd[i] = __ldg(p.dw_b + i);
  }
  for (int i = threadIdx.x; i < p.tile_n; i += kBnThreads) s_b2[i] = __ldg(p.b2 + i);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp == kBnTapWarps + kBnEpiWarps + 1) {
    // ================================ TMA producer ===========================================================
    if (lane == 0 && my_tiles > 0) {
      tma_prefetch_desc(&p.tm_x);
      mbar_arrive_expect_tx(bar_w, mid_pad * 128 + nblk * w2_blk);
      tma_load_2d(s_w1, &p.tm_w1, bar_w, 0, 0);
      for (int cb = 0; cb < nblk; ++cb) tma_load_2d(s_w2 + cb * w2_blk, &p.tm_w2, bar_w, cb * kBnCB, 0);
      pdl_wait();  // x is produced by the previous kernel; every store of this kernel follows causally
      int it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const int img = t / tiles_per_img, r = t - img * tiles_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        mbar_wait(bar_x_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(bar_x_full, kXBytes);
        tma_load_tile_4d(s_x, &p.tm_x, bar_x_full, 0, tx * kBnTX - P, ty * kBnTY - P, img);
      }
    }
  } else if (warp == kBnTapWarps + kBnEpiWarps) {
    // ================================ MMA issuer =============================================================
    if (lane == 0 && my_tiles > 0) {
      mbar_wait(bar_w, 0);
      tc_fence_after_sync();
      auto mma1 = [&](int s) {
        const int it = s / nblk, cb = s - it * nblk, buf = s & 1;
        if (cb == 0) {
          mbar_wait(bar_x_full, it & 1);
          tc_fence_after_sync();
        }
        mbar_wait(&bar_acc1_empty[buf], ((s >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint64_t db = umma_smem_desc_sw128(smem_u32(s_w1 + cb * (kBnCB * 128)));
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          const uint64_t da = umma_smem_desc_sw128(smem_u32(s_x + mt * 16384));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem_base + buf * kBnAcc1Cols + mt * kBnCB, da + 2 * k, db + 2 * k, p.idesc1, k != 0 ? 1u : 0u);
        }
        tc_commit(&bar_acc1_full[buf]);
        if (cb == nblk - 1) tc_commit(bar_x_empty);  // the halo tile may be overwritten once these MMAs are done
        BN_TRACE(0, s, 0);
      };
      auto mma2 = [&](int s) {
        const int it = s / nblk, cb = s - it * nblk;
        if (cb == 0) {
          mbar_wait(bar_acc2_empty, (it & 1) ^ 1);
          tc_fence_after_sync();
        }
        mbar_wait(bar_a2_full, s & 1);
        tc_fence_after_sync();
        const uint64_t db = umma_smem_desc_sw128(smem_u32(s_w2 + cb * w2_blk));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint64_t da = umma_smem_desc_sw128(smem_u32(s_a2 + mt * 16384));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem_base + kBnAcc2Col0 + mt * p.tile_n, da + 2 * k, db + 2 * k, p.idesc2, (cb | k) != 0 ? 1u : 0u);
        }
        tc_commit(bar_a2_empty);
        if (cb == nblk - 1) tc_commit(bar_acc2_full);
        BN_TRACE(0, s, 1);
      };
      // MMA1 runs two steps ahead of MMA2: MMA1(s + 2) only needs epilogue 1 of step s (acc1[s & 1] drained), which
      // precedes the end of the taps of step s that MMA2(s) waits for — so it is issued first and epilogue 1 never
      // waits for the tensor core (measured with the clock64 trace: 3000 cycles of bubble per step the other way round)
      mma1(0);
      if (total_steps > 1) mma1(1);
      for (int s = 0; s < total_steps; ++s) {
        if (s + 2 < total_steps) mma1(s + 2);
        mma2(s);
      }
    }
  } else if (warp >= kBnTapWarps) {
    // ================================ epilogue warps =========================================================
    const int q = warp & 3;  // TMEM lane quarter of this warp (kBnTapWarps % 4 == 0)
    auto epilogue2 = [&](int it) {
      const int t = blockIdx.x + it * gridDim.x;
      const int img = t / tiles_per_img, r = t - img * tiles_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      mbar_wait(bar_acc2_full, it & 1);
      tc_fence_after_sync();
      // 2 M tiles x tile_n columns per thread in 32-column groups; the TMEM load of the next group is in flight while
      // this one is processed, bias in registers before the math (same reasons as epilogue 1)
      const int groups_per_mt = (p.tile_n + 31) >> 5;  // 1 or 2
      const int n_groups = 2 * groups_per_mt;
      const uint32_t taddr0 = tmem_base + kBnAcc2Col0 + (static_cast<uint32_t>(q * 32) << 16);
      uint32_t rr[2][32];
      __syncwarp();
      tmem_ld_32x32b_x32(taddr0, rr[0]);
#pragma unroll 1
      for (int gsel = 0; gsel < n_groups; gsel += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int gi = gsel + u;
          if (gi >= n_groups) break;
          const int mt = gi / groups_per_mt, c = (gi - mt * groups_per_mt) * 32;
          float bias[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(&bias[j]) = *reinterpret_cast<const float4*>(s_b2 + c + j);
          tmem_ld_wait();
          if (gi + 1 < n_groups) {
            const int mt2 = (gi + 1) / groups_per_mt, c2 = ((gi + 1) - mt2 * groups_per_mt) * 32;
            __syncwarp();
            tmem_ld_32x32b_x32(taddr0 + mt2 * p.tile_n + c2, rr[u ^ 1]);
          }
          const int prow = mt * 128 + q * 32 + lane;
          const int py = prow / kBnTX, px = prow - py * kBnTX;
          const int gy = ty * kBnTY + py, gx = tx * kBnTX + px;
          const bool ok = prow < kBnTX * kBnTY && gy < p.H && gx < p.W;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            pk[j] = pack_half2(silu_fast(__uint_as_float(rr[u][2 * j]) + bias[2 * j]),
                               silu_fast(__uint_as_float(rr[u][2 * j + 1]) + bias[2 * j + 1]));
          if (ok) {
            __half* orow = p.out + ((static_cast<size_t>(img) * p.H + gy) * p.W + gx) * p.out_ld + c;
#pragma unroll
            for (int h16 = 0; h16 < 2; ++h16) {
              const int cc = c + 16 * h16;
              if (cc + 16 <= p.N) {
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + 16 * h16),
                             "r"(pk[8 * h16]), "r"(pk[8 * h16 + 1]), "r"(pk[8 * h16 + 2]), "r"(pk[8 * h16 + 3]),
                             "r"(pk[8 * h16 + 4]), "r"(pk[8 * h16 + 5]), "r"(pk[8 * h16 + 6]), "r"(pk[8 * h16 + 7])
                             : "memory");
              } else if (cc + 8 <= p.N) {  // N % 16 == 8
                *reinterpret_cast<uint4*>(orow + 16 * h16) = make_uint4(pk[8 * h16], pk[8 * h16 + 1], pk[8 * h16 + 2], pk[8 * h16 + 3]);
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc2_empty);
    };

    for (int s = 0; s < total_steps; ++s) {
      const int it = s / nblk, cb = s - it * nblk, buf = s & 1;
      const int t = blockIdx.x + it * gridDim.x;
      const int img_r = t % tiles_per_img;
      const int ty = img_r / p.tiles_x, tx = img_r - ty * p.tiles_x;
      const int hy0 = ty * kBnTY - P, hx0 = tx * kBnTX - P;  // image coordinates of halo pixel (0, 0)
      if (q == 0) BN_TRACE(1, s, 0);
      mbar_wait(&bar_acc1_full[buf], (s >> 1) & 1);
      tc_fence_after_sync();
      if (q == 0) BN_TRACE(1, s, 1);
      mbar_wait(&bar_t1_empty[buf], ((s >> 1) & 1) ^ 1);
      if (q == 0) BN_TRACE(1, s, 2);
      uint8_t* t1 = s_t1 + buf * kXBytes;
      const float* b1 = s_b1 + cb * kBnCB;
      // 6 half rows (3 M tiles x 2 x 32 columns) per thread.  The TMEM load of half h + 1 is in flight while half h is
      // processed; the 32 bias values go to registers BEFORE the math so that the 32 SiLU chains are independent of the
      // shared-memory stores (the compiler must assume the T1 stores alias the bias array: with the loads in between,
      // the chains ran 8 at a time — 35 cycles per element in the clock64 trace).
      uint32_t rr[2][32];
      const uint32_t taddr0 = tmem_base + buf * kBnAcc1Cols + (static_cast<uint32_t>(q * 32) << 16);
      __syncwarp();
      tmem_ld_32x32b_x32(taddr0, rr[0]);
#pragma unroll
      for (int hh = 0; hh < 6; ++hh) {
        const int mt = hh >> 1, half = hh & 1;
        float bias[32];
#pragma unroll
        for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(&bias[c]) = *reinterpret_cast<const float4*>(b1 + 32 * half + c);
        tmem_ld_wait();
        if (hh + 1 < 6) {
          __syncwarp();
          tmem_ld_32x32b_x32(taddr0 + ((hh + 1) >> 1) * kBnCB + 32 * ((hh + 1) & 1), rr[(hh + 1) & 1]);
        }
        const int prow = mt * 128 + q * 32 + lane;  // halo pixel of this thread
        const int py = prow / TW, px = prow - py * TW;
        const int gy = hy0 + py, gx = hx0 + px;
        const bool row_ok = prow < kHaloRows;
        const bool inside = row_ok && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = silu_fast(__uint_as_float(rr[hh & 1][2 * j]) + bias[2 * j]);
          const float b = silu_fast(__uint_as_float(rr[hh & 1][2 * j + 1]) + bias[2 * j + 1]);
          pk[j] = inside ? pack_half2(a, b) : 0u;  // the depth-wise conv zero-pads t1 (common.py:915-923)
        }
        if (row_ok) {
          uint8_t* trow = t1 + prow * 128;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            *reinterpret_cast<uint4*>(trow + (((4 * half + ch) ^ (prow & 7)) << 4)) =
                make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bar_acc1_empty[buf]);
        mbar_arrive(&bar_t1_full[buf]);
      }
      if (q == 0) BN_TRACE(1, s, 3);
      // the output tile whose last block was step s - 1: its MMA2 completes while the taps of step s run
      if (s >= 1 && (s - 1) % nblk == nblk - 1) epilogue2((s - 1) / nblk);
    }
    if (total_steps > 0) epilogue2(my_tiles - 1);
  } else {
    // ================================ depth-wise tap warps ===================================================
    const int uy = warp >> 2, ux = warp & 3;  // 2 x 4 units of 5 x 5 pixels
    const int oy0 = uy * 5, ox0 = ux * 5;
    const int p0 = oy0 * TW + ox0;            // halo row of the unit's first window pixel
    for (int s = 0; s < total_steps; ++s) {
      const int cb = s % nblk, buf = s & 1;
      const int c0 = cb * kBnCB;
      float2 wreg[K * K];
#pragma unroll
      for (int t = 0; t < K * K; ++t) wreg[t] = *reinterpret_cast<const float2*>(s_dw + t * mid_pad + c0 + 2 * lane);
      const float2 bv = *reinterpret_cast<const float2*>(s_bd + c0 + 2 * lane);
      // T1 row r holds its 16-byte chunk j at position j ^ (r & 7): one base pointer per value of (row & 7), so that
      // every window load below is [base + immediate]
      const uint8_t* t1 = s_t1 + buf * kXBytes + p0 * 128 + ((lane & 3) << 2);
      const uint8_t* tb[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) tb[v] = t1 + ((((p0 + v) & 7) ^ (lane >> 2)) << 4);

      if (warp == 0) BN_TRACE(2, s, 0);
      mbar_wait(&bar_t1_full[buf], (s >> 1) & 1);
      if (warp == 0) BN_TRACE(2, s, 1);

      float2 acc[5][5];
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int r = 0; r < 5; ++r) acc[i][r] = bv;
#pragma unroll
      for (int d = 0; d < 5 + K - 1; ++d) {
        float2 win[5 + K - 1];
#pragma unroll
        for (int j = 0; j < 5 + K - 1; ++j)
          win[j] = __half22float2(*reinterpret_cast<const __half2*>(tb[(d * TW + j) & 7] + (d * TW + j) * 128));
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int ky = d - i;
          if (ky < 0 || ky >= K) continue;
#pragma unroll
          for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) acc[i][r] = ffma2(win[r + kx], wreg[ky * K + kx], acc[i][r]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_t1_empty[buf]);  // this warp's reads of T1[buf] are complete
      if (warp == 0) BN_TRACE(2, s, 2);

      mbar_wait(bar_a2_empty, (s & 1) ^ 1);  // MMA2 of the previous step has read the A2 tile
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const int prow = (oy0 + i) * kBnTX + ox0 + r;
          *reinterpret_cast<uint32_t*>(s_a2 + prow * 128 + ((((lane >> 2) ^ (prow & 7)) << 4) | ((lane & 3) << 2))) =
              pack_half2(silu_fast(acc[i][r].x), silu_fast(acc[i][r].y));
        }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a2_full);
      if (warp == 0) BN_TRACE(2, s, 3);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kBnTapWarps + kBnEpiWarps) tmem_dealloc(tmem_base, 512);
}

// ================================================================================================================
// K4, second shape — two CTAs per SM, every warp does every phase (the structure of dwpw_kernel, which measured
// faster than the warp-specialised persistent shape above because 16 tap warps per SM hide each other's latencies):
//   one CTA = one 8 x 16 output tile of one image; halo (8+k-1) x (16+k-1) <= 240 pixels = 2 M tiles of the expand GEMM;
//   per 64-channel block cb of the mid channels:
//     thread 0: MMA1(cb) = Xhalo[256 x 64] * W1[cb]^T -> acc1 (128 TMEM columns); W1 / W2 blocks stream through one
//               shared-memory slot each (TMA), so shared memory does not grow with `mid`
//     all 8 warps: epilogue 1 (warp = (M tile, lane quarter), one halo pixel per thread): +b1, SiLU, 0 outside the image,
//               fp16 -> T1 (chunk-swizzled);  __syncthreads;  thread 0 issues MMA1(cb+1) — it overlaps the taps
//     all 8 warps: taps, warp = 4 x 4 pixel unit, lane = channel pair (16 x k*k FFMA2), +bd, SiLU -> A2 (SW128 by hand)
//     thread 0: MMA2(cb): acc2[128 x c_] += A2 * W2[cb]^T
//   final epilogue: acc2 -> +b2 -> SiLU -> global (warps 0-3 / 4-7 take the two 32-column halves).
// ~105 KB of shared memory and 256 TMEM columns per CTA.
constexpr int kB2TX = 16, kB2TY = 8;
constexpr int kB2Threads = 256;

template <int K>
__global__ void __launch_bounds__(kB2Threads, 2) bneck2_kernel(const __grid_constant__ BneckParams p) {
  constexpr int P = K / 2;
  constexpr int TW = kB2TX + K - 1, TH = kB2TY + K - 1;
  constexpr int kHaloRows = TH * TW;               // 240 (k = 5) / 180 (k = 3)
  constexpr uint32_t kXBytes = kHaloRows * 128;
  constexpr int U = 4;                             // unit = 4 x 4 output pixels per warp
  static_assert(kHaloRows <= 256, "halo must fit two M tiles");

  extern __shared__ uint8_t smem_b2_raw[];
  uint8_t* smem = smem_b2_raw + ((1024u - (smem_u32(smem_b2_raw) & 1023u)) & 1023u);
  constexpr uint32_t kXAlloc = ((kXBytes + 1023) / 1024) * 1024;
  uint8_t* s_x = smem;                              // [kHaloRows][128 B]; MMA1 reads 256 rows (the tail = rows of A2: unused)
  uint8_t* s_a2 = s_x + kXAlloc;                    // [128 rows][128 B]
  uint8_t* s_w1 = s_a2 + 16384;                     // 2 slots x [64 rows][128 B]
  uint8_t* s_w2 = s_w1 + 2 * 8192;                  // 2 slots x [tile_n rows][128 B] (8 KB apart)
  uint8_t* s_t1 = s_w2 + 2 * 8192;                  // [kHaloRows][128 B], chunk-swizzled by hand
  float* s_b2 = reinterpret_cast<float*>(s_t1 + ((kXBytes + 127) / 128) * 128);  // [64]
  float* s_b1 = s_b2 + 64;                          // [mid_pad]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b1 + p.mid_pad);
  uint64_t* bar_x = bars;
  uint64_t* bar_w1 = bars + 1;   // [2] per slot
  uint64_t* bar_w2 = bars + 3;   // [2] per slot
  uint64_t* bar_mma1 = bars + 5;
  uint64_t* bar_mma2 = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_x = blockIdx.x % p.tiles_x, tile_y = blockIdx.x / p.tiles_x;
  const int img = blockIdx.y;
  const int x0 = tile_x * kB2TX, y0 = tile_y * kB2TY;
  const int nblk = p.nblk, mid_pad = p.mid_pad;
  const int w2_bytes = p.tile_n * 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tm_x);
    tma_prefetch_desc(&p.tm_w1b);
    tma_prefetch_desc(&p.tm_w2);
    mbar_init(bar_x, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_w1[i], 1);
      mbar_init(&bar_w2[i], 1);
    }
    mbar_init(bar_mma1, 1);
    mbar_init(bar_mma2, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.tile_n; i += kB2Threads) s_b2[i] = __ldg(p.b2 + i);
  for (int i = threadIdx.x; i < p.mid_pad; i += kB2Threads) s_b1[i] = __ldg(p.b1 + i);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc2_col = 128;

  auto issue_mma1 = [&](int cb) {  // thread 0 only
    if (cb == 0) mbar_wait(bar_x, 0);
    mbar_wait(&bar_w1[cb & 1], (cb >> 1) & 1);
    tc_fence_after_sync();
    const uint64_t db = umma_smem_desc_sw128(smem_u32(s_w1 + (cb & 1) * 8192));
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const uint64_t da = umma_smem_desc_sw128(smem_u32(s_x + mt * 16384));
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base + mt * kBnCB, da + 2 * k, db + 2 * k, p.idesc1, k != 0 ? 1u : 0u);
    }
    tc_commit(bar_mma1);
  };

  if (threadIdx.x == 0) {
    pdl_launch_dependents();
    for (int i = 0; i < 2 && i < nblk; ++i) {  // weights are constants: may start before the previous kernel ends
      mbar_arrive_expect_tx(&bar_w1[i], 8192);
      tma_load_2d(s_w1 + i * 8192, &p.tm_w1b, &bar_w1[i], 0, i * kBnCB);
      mbar_arrive_expect_tx(&bar_w2[i], w2_bytes);
      tma_load_2d(s_w2 + i * 8192, &p.tm_w2, &bar_w2[i], i * kBnCB, 0);
    }
    pdl_wait();  // x (and, causally, every output store) follows the previous kernels
    mbar_arrive_expect_tx(bar_x, kXBytes);
    tma_load_tile_4d(s_x, &p.tm_x, bar_x, 0, x0 - P, y0 - P, img);
    issue_mma1(0);
  }

  // epilogue-1 role of this thread: one halo pixel
  const int e_mt = warp >> 2, e_q = warp & 3;
  const int e_row = e_mt * 128 + e_q * 32 + lane;
  const int e_py = e_row / TW, e_px = e_row - e_py * TW;
  const bool e_row_ok = e_row < kHaloRows;
  const bool e_inside = e_row_ok && y0 - P + e_py >= 0 && y0 - P + e_py < p.H && x0 - P + e_px >= 0 && x0 - P + e_px < p.W;
  const uint32_t e_taddr = tmem_base + e_mt * kBnCB + (static_cast<uint32_t>(e_q * 32) << 16);
  // tap role: 4 x 4 unit
  const int uy = warp >> 2, ux = warp & 3;
  const int oy0 = uy * U, ox0 = ux * U;
  const int p0 = oy0 * TW + ox0;

  for (int cb = 0; cb < nblk; ++cb) {
    const int c0 = cb * kBnCB;
    // ---- epilogue 1: acc1 -> T1 ------------------------------------------------------------------------------------
    mbar_wait(bar_mma1, cb & 1);
    tc_fence_after_sync();
    if (threadIdx.x == 0 && cb + 2 < nblk) {  // MMA1(cb) is complete: its W1 slot takes the panel of block cb + 2
      mbar_arrive_expect_tx(&bar_w1[cb & 1], 8192);
      tma_load_2d(s_w1 + (cb & 1) * 8192, &p.tm_w1b, &bar_w1[cb & 1], 0, c0 + 2 * kBnCB);
    }
    {
      uint32_t rr[2][32];
      __syncwarp();
      tmem_ld_32x32b_x32(e_taddr, rr[0]);
      tmem_ld_32x32b_x32(e_taddr + 32, rr[1]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // bias to registers BEFORE the math: the T1 stores below may alias s_b1 as far as the compiler knows
        float bias[32];
#pragma unroll
        for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(&bias[c]) = *reinterpret_cast<const float4*>(s_b1 + c0 + 32 * h + c);
        if (h == 0) tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = silu_fast(__uint_as_float(rr[h][2 * j]) + bias[2 * j]);
          const float b = silu_fast(__uint_as_float(rr[h][2 * j + 1]) + bias[2 * j + 1]);
          pk[j] = e_inside ? pack_half2(a, b) : 0u;  // the depth-wise conv zero-pads t1 (common.py:915-923)
        }
        if (e_row_ok) {
          uint8_t* trow = s_t1 + e_row * 128;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            *reinterpret_cast<uint4*>(trow + (((4 * h + ch) ^ (e_row & 7)) << 4)) =
                make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        }
      }
    }
    // depth-wise weights of this thread's two channels (L2-resident, coalesced): in flight across the barrier
    float2 wreg[K * K];
#pragma unroll
    for (int t = 0; t < K * K; ++t) wreg[t] = __ldg(reinterpret_cast<const float2*>(p.dw_w + static_cast<size_t>(t) * mid_pad + c0 + 2 * lane));
    const float2 bv = __ldg(reinterpret_cast<const float2*>(p.dw_b + c0 + 2 * lane));
    tc_fence_before_sync();
    __syncthreads();  // T1 complete, acc1 drained
    tc_fence_after_sync();
    if (threadIdx.x == 0 && cb + 1 < nblk) issue_mma1(cb + 1);  // overlaps the taps below

    // ---- taps: T1 -> registers --------------------------------------------------------------------------------------
    const uint8_t* t1 = s_t1 + p0 * 128 + ((lane & 3) << 2);
    const uint8_t* tb[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) tb[v] = t1 + ((((p0 + v) & 7) ^ (lane >> 2)) << 4);
    float2 acc[U][U];
#pragma unroll
    for (int i = 0; i < U; ++i)
#pragma unroll
      for (int r = 0; r < U; ++r) acc[i][r] = bv;
#pragma unroll
    for (int d = 0; d < U + K - 1; ++d) {
      float2 win[U + K - 1];
#pragma unroll
      for (int j = 0; j < U + K - 1; ++j)
        win[j] = __half22float2(*reinterpret_cast<const __half2*>(tb[(d * TW + j) & 7] + (d * TW + j) * 128));
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int ky = d - i;
        if (ky < 0 || ky >= K) continue;
#pragma unroll
        for (int r = 0; r < U; ++r)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) acc[i][r] = ffma2(win[r + kx], wreg[ky * K + kx], acc[i][r]);
      }
    }
    // ---- A2 tile + MMA2 -----------------------------------------------------------------------------------------------
    if (cb > 0) mbar_wait(bar_mma2, (cb - 1) & 1);  // MMA2 of the previous block has read A2 (and its W2 slot)
    if (threadIdx.x == 0 && cb > 0 && cb + 1 < nblk) {  // the freed W2 slot takes the panel of block cb + 1
      mbar_arrive_expect_tx(&bar_w2[(cb + 1) & 1], w2_bytes);
      tma_load_2d(s_w2 + ((cb + 1) & 1) * 8192, &p.tm_w2, &bar_w2[(cb + 1) & 1], c0 + kBnCB, 0);
    }
#pragma unroll
    for (int i = 0; i < U; ++i)
#pragma unroll
      for (int r = 0; r < U; ++r) {
        const int prow = (oy0 + i) * kB2TX + ox0 + r;
        *reinterpret_cast<uint32_t*>(s_a2 + prow * 128 + ((((lane >> 2) ^ (prow & 7)) << 4) | ((lane & 3) << 2))) =
            pack_half2(silu_fast(acc[i][r].x), silu_fast(acc[i][r].y));
      }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();  // A2 complete; every warp is done reading T1
    tc_fence_after_sync();
    if (threadIdx.x == 0) {
      mbar_wait(&bar_w2[cb & 1], (cb >> 1) & 1);
      tc_fence_after_sync();
      const uint64_t da = umma_smem_desc_sw128(smem_u32(s_a2));
      const uint64_t db = umma_smem_desc_sw128(smem_u32(s_w2 + (cb & 1) * 8192));
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base + acc2_col, da + 2 * k, db + 2 * k, p.idesc2, (cb | k) != 0 ? 1u : 0u);
      tc_commit(bar_mma2);
    }
  }

  // ---- final epilogue: acc2 row = tile pixel; warps 0-3 columns 0..31, warps 4-7 columns 32..63 ---------------------
  mbar_wait(bar_mma2, (nblk - 1) & 1);
  tc_fence_after_sync();
  {
    const int q = warp & 3, c = (warp >> 2) * 32;
    const int prow = q * 32 + lane;
    const int py = prow / kB2TX, px = prow - py * kB2TX;
    const int gy = y0 + py, gx = x0 + px;
    const bool ok = gy < p.H && gx < p.W;
    uint32_t rr[32];
    __syncwarp();
    if (c < p.tile_n) {  // warp-uniform
      tmem_ld_32x32b_x32(tmem_base + acc2_col + c + (static_cast<uint32_t>(q * 32) << 16), rr);
      tmem_ld_wait();
      if (ok) {
        __half* orow = p.out + ((static_cast<size_t>(img) * p.H + gy) * p.W + gx) * p.out_ld + c;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          pk[j] = pack_half2(silu_fast(__uint_as_float(rr[2 * j]) + s_b2[c + 2 * j]),
                             silu_fast(__uint_as_float(rr[2 * j + 1]) + s_b2[c + 2 * j + 1]));
#pragma unroll
        for (int h16 = 0; h16 < 2; ++h16) {
          const int cc = c + 16 * h16;
          if (cc + 16 <= p.N) {
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + 16 * h16), "r"(pk[8 * h16]),
                         "r"(pk[8 * h16 + 1]), "r"(pk[8 * h16 + 2]), "r"(pk[8 * h16 + 3]), "r"(pk[8 * h16 + 4]),
                         "r"(pk[8 * h16 + 5]), "r"(pk[8 * h16 + 6]), "r"(pk[8 * h16 + 7])
                         : "memory");
          } else if (cc + 8 <= p.N) {
            *reinterpret_cast<uint4*>(orow + 16 * h16) = make_uint4(pk[8 * h16], pk[8 * h16 + 1], pk[8 * h16 + 2], pk[8 * h16 + 3]);
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

static size_t bneck2_smem_bytes(int k) {
  const size_t halo = static_cast<size_t>(kB2TX + k - 1) * (kB2TY + k - 1) * 128;
  return 1024 + ((halo + 1023) / 1024) * 1024 + 16384 + 4 * 8192 + ((halo + 127) / 128) * 128 + 64 * 4 + 7 * 8 + 16;  // + 4 * mid_pad (biases)
}

template <int K>
static int32_t launch_bneck2(BneckParams& p, int n, cudaStream_t st) {
  const size_t smem = bneck2_smem_bytes(K) + static_cast<size_t>(p.mid_pad) * 4;
  {
    static SmemOptIn opt_in;
    const int32_t rc = smem_opt_in(opt_in, bneck2_kernel<K>, 113 * 1024, "bottleneck(2-CTA shape)");
    if (rc) return rc;
  }
  launch_pdl(bneck2_kernel<K>, dim3(p.tiles_x * p.tiles_y, n), dim3(kB2Threads), smem, st, p);
  return check_launch("bottleneck kernel launch");
}

static size_t bneck_smem_bytes(int k, int mid_pad, int nblk, int tile_n) {
  const size_t halo = static_cast<size_t>(kBnTX + k - 1) * (kBnTY + k - 1) * 128;
  return 1024 + static_cast<size_t>(mid_pad) * 128 + static_cast<size_t>(nblk) * tile_n * 128 + halo + kBnA2Bytes + 2 * halo +
         (static_cast<size_t>(k) * k * mid_pad + 2 * mid_pad + tile_n + 2) * 4 + 16 * 8 + 16;
}

template <int K>
static int32_t launch_bneck(BneckParams& p, size_t smem, cudaStream_t st) {
  {
    static SmemOptIn opt_in;
    const int32_t rc = smem_opt_in(opt_in, bneck_kernel<K>, 227 * 1024, "bottleneck");
    if (rc) return rc;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    sms = 148;
  const int grid = p.tiles < sms ? p.tiles : sms;
  launch_pdl(bneck_kernel<K>, dim3(grid), dim3(kBnThreads), smem, st, p);
  return check_launch("bottleneck kernel launch");
}

}  // namespace mafb200

using namespace mafb200;

// Can mafb200_bottleneck run this shape?  (pure host arithmetic; the engine asks before planning the fused op)
// MAFB200_BNECK_SHAPE: 2 (default) = two CTAs per SM, every warp does every phase, weight panels streamed (any mid <= 512);
// 1 = the persistent warp-specialised CTA per SM with resident panels (mid <= 192).
static int bneck_shape() {
  static const int shape = [] {
    const char* e = getenv("MAFB200_BNECK_SHAPE");
    return (e && e[0] == '1') ? 1 : 2;
  }();
  return shape;
}

extern "C" int32_t mafb200_bottleneck_supported(int32_t c_in, int32_t mid, int32_t c_out, int32_t k) {
  const int max_mid = bneck_shape() == 1 ? kBnMaxMid : 512;
  if ((k != 3 && k != 5) || c_in < 8 || c_in > kBnMaxCin || c_in % 8 || mid < 8 || mid > max_mid || mid % 8 ||
      c_out < 8 || c_out > 64 || c_out % 8)
    return 0;
  if (bneck_shape() == 2) return 1;
  const int mid_pad = round_up(mid, kBnCB), tile_n = round_up(c_out, 16);
  return bneck_smem_bytes(k, mid_pad, mid_pad / kBnCB, tile_n) <= 227 * 1024 ? 1 : 0;
}

static thread_local long long* g_bneck_trace = nullptr;
// Debug hook (not part of the hot path): the next mafb200_bottleneck calls of this thread record clock64 stamps of CTA 0
// into `trace` (device, 3 roles x 64 steps x 4 slots of int64; see BN_TRACE); NULL turns it off.
extern "C" int32_t mafb200_bottleneck_trace(long long* trace) {
  g_bneck_trace = trace;
  return MAF_OK;
}

// dst = SiLU(W2 * SiLU(DW_k(SiLU(W1 * src + b1)) + dw_bias) + b2)     (DepthBottleneckUni, common.py:898-927)
//   w1_packed fp16 [mid_pad][64]  (row = mid channel, zero rows / columns beyond mid / c_in),  b1 fp32 [mid_pad]
//   dw_weight fp32 [k*k][mid_pad] (tap-major, zero beyond mid),                                 dw_bias fp32 [mid_pad]
//   w2_packed fp16 [tile_n][mid_pad] (row = output channel),                                    b2 fp32 [tile_n]
// with mid_pad = round_up(mid, 64), tile_n = round_up(dst->c, 16).
extern "C" int32_t mafb200_bottleneck(const maf_tensor* src, int32_t mid, const void* w1_packed, const float* b1,
                                      const float* dw_weight, const float* dw_bias, int32_t k, const void* w2_packed,
                                      const float* b2, const maf_tensor* dst, void* stream) {
  if (!valid_f16_view(src) || !valid_f16_view(dst)) return fail(MAF_E_ARG, "bottleneck: bad src/dst");
  if (!w1_packed || !b1 || !dw_weight || !dw_bias || !w2_packed || !b2) return fail(MAF_E_ARG, "bottleneck: null weights");
  if (!same_nhw(src, dst)) return fail(MAF_E_ARG, "bottleneck: src/dst n/h/w differ");
  if (!mafb200_bottleneck_supported(src->c, mid, dst->c, k))
    return fail(MAF_E_ARG, "bottleneck: unsupported shape (c_in=%d mid=%d c_out=%d k=%d)", src->c, mid, dst->c, k);
  if (!aligned_f16_view(src)) return fail(MAF_E_ALIGN, "bottleneck: src must be 16-B aligned with c_stride %% 8 == 0");
  if ((reinterpret_cast<uintptr_t>(dst->ptr) & 31) || (dst->c_stride % 16) != 0)
    return fail(MAF_E_ALIGN, "bottleneck: dst must be 32-B aligned with c_stride %% 16 == 0");
  if ((reinterpret_cast<uintptr_t>(w1_packed) & 15) || (reinterpret_cast<uintptr_t>(w2_packed) & 15) ||
      (reinterpret_cast<uintptr_t>(dw_weight) & 7) || (reinterpret_cast<uintptr_t>(dw_bias) & 7))
    return fail(MAF_E_ALIGN, "bottleneck: weight alignment");
  int32_t rc = require_sm100();
  if (rc) return rc;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(MAF_E_ARCH, "cuTensorMapEncodeTiled entry point not available");

  BneckParams p;
  memset(&p, 0, sizeof(p));
  const int mid_pad = round_up(mid, kBnCB), tile_n = round_up(dst->c, 16);
  const int shape = bneck_shape();
  const int tx = shape == 1 ? kBnTX : kB2TX, ty = shape == 1 ? kBnTY : kB2TY;
  {
    const int TW = tx + k - 1, TH = ty + k - 1;
    const cuuint64_t px = static_cast<cuuint64_t>(src->c_stride) * 2;
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(src->c), static_cast<cuuint64_t>(src->w),
                          static_cast<cuuint64_t>(src->h), static_cast<cuuint64_t>(src->n)};
    cuuint64_t strides[3] = {px, px * src->w, px * src->w * src->h};
    cuuint32_t box[4] = {kBnCB, static_cast<cuuint32_t>(TW), static_cast<cuuint32_t>(TH), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, src->ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MAF_E_CUDA, "bottleneck: cuTensorMapEncodeTiled(src) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {kBnCB, static_cast<cuuint64_t>(mid_pad)};
    cuuint64_t strides[1] = {kBnCB * 2};
    cuuint32_t box[2] = {kBnCB, static_cast<cuuint32_t>(mid_pad <= 256 ? mid_pad : 256)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tm_w1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w1_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MAF_E_CUDA, "bottleneck: cuTensorMapEncodeTiled(W1) failed: %d", (int)r);
    box[1] = kBnCB;  // one 64-channel block per load (2-CTA shape)
    r = enc(&p.tm_w1b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w1_packed), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MAF_E_CUDA, "bottleneck: cuTensorMapEncodeTiled(W1 block) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(mid_pad), static_cast<cuuint64_t>(tile_n)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(mid_pad) * 2};
    cuuint32_t box[2] = {kBnCB, static_cast<cuuint32_t>(tile_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tm_w2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w2_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MAF_E_CUDA, "bottleneck: cuTensorMapEncodeTiled(W2) failed: %d", (int)r);
  }
  p.b1 = b1;
  p.dw_w = dw_weight;
  p.dw_b = dw_bias;
  p.b2 = b2;
  p.out = static_cast<__half*>(dst->ptr);
  p.out_ld = dst->c_stride;
  p.H = src->h;
  p.W = src->w;
  p.N = dst->c;
  p.tile_n = tile_n;
  p.mid_pad = mid_pad;
  p.nblk = mid_pad / kBnCB;
  p.tiles_x = ceil_div(src->w, tx);
  p.tiles_y = ceil_div(src->h, ty);
  const long long tiles = static_cast<long long>(src->n) * p.tiles_x * p.tiles_y;
  if (tiles > 0x7fffffff) return fail(MAF_E_ARG, "bottleneck: too many tiles");
  p.tiles = static_cast<int32_t>(tiles);
  p.trace = g_bneck_trace;
  p.idesc1 = umma_idesc_f16(128, kBnCB);
  p.idesc2 = umma_idesc_f16(128, tile_n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (shape == 2) {
    if (src->n > 65535) return fail(MAF_E_ARG, "bottleneck: batch %d > 65535", src->n);
    return k == 3 ? launch_bneck2<3>(p, src->n, st) : launch_bneck2<5>(p, src->n, st);
  }
  const size_t smem = bneck_smem_bytes(k, mid_pad, p.nblk, tile_n);
  return k == 3 ? launch_bneck<3>(p, smem, st) : launch_bneck<5>(p, smem, st);
}
