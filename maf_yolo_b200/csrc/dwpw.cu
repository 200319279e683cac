// K3c — depth-wise k x k (k <= 5) FUSED with the 1x1 conv that consumes it:
//
//     y = act2( W2 * act1( DW_k(x) + b1 ) + b2 )          x: [n,h,w,C] fp16 NHWC,  y: [n,h,w,N]
//
// i.e. DepthBottleneckUni's `conv2 -> SiLU -> one_conv` (yolov6/layers/common.py:915-926: UniRepLKNetBlock deploy
// form, SiLU, Conv 1x1 + SiLU) and Head_DepthUni's `cls_conv -> cls_conv_s` / `reg_conv -> reg_conv_s`
// (common.py:1328-1336: depth-wise, NO activation, Conv 1x1 + SiLU).  The 3c_-wide depth-wise output never goes to
// HBM: it is the most expensive intermediate of a RepHDW block (written by a store-bound kernel, read back by a
// TMA-row-bound one).
//
// One CTA = one 10 x 20 output tile of one image, ALL channels; 256 threads = 8 warps, each warp one 5 x 5 pixel unit
// with lane = channel pair (the FFMA loop of dwconv.cu at CB = 64).  Per 64-channel block:
//   TMA halo tile (14 x 24 x 64 ch, OOB zero fill = padding)  ->  25 x 25 packed FFMA2 per thread  ->  bias + act1  ->
//   fp16 pairs written by hand into the SWIZZLE_128B K-major A tile (row = 25 * warp + pixel of the warp's unit, 16-byte chunk
//   index ^ (row & 7): a warp writes the 128 bytes of ONE row per instruction, conflict free)  ->  fence.proxy.async  ->
//   one thread issues 2 M tiles x 4 tcgen05.mma (128 x N x 16) against that block's slice of W2 (it arrives with the
//   halo tile, two slots), accumulating over the channel blocks in TMEM.
// Epilogue: 8 warps drain the two 128-row accumulators (row = pixel), bias + act2, 256-bit stores.
// Rows 200..255 of the A tile are never written (their accumulator rows are never read).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"

namespace mafb200 {

constexpr int kFwCB = 64;               // channels per block = one 128-byte swizzle row

// Tile shapes: a CTA is (TX/5) x (TY/5) warps, each one 5 x 5 pixel unit.  10 x 20 = 8 warps, 2 M tiles (round 1);
// 10 x 10 = 4 warps, 1 M tile (round 2): half the work per CTA and twice the CTAs per SM, i.e. FOUR instead of two
// independently progressing agents per SM for the same 16 warps.  The kernel alternates an FMA-pipe phase (taps) with
// phases that leave the pipe idle (A-tile write, MMA wait, epilogue, start-up), all CTA-wide in lock step, so the pipe
// is only busy while at least one resident CTA is in its taps: ablation builds (tools/bench_dwpw.py,
// profiles/r02_j_dwpw_ablation.md) showed the halo TMA and the weight loads off the critical path, the taps ~50 % of the
// run time and the FMA pipe 54 % busy (ncu).
template <int TX, int TY>
struct FwTile {
  static constexpr int kWarps = (TX / 5) * (TY / 5);
  static constexpr int kThreads = kWarps * 32;
  static constexpr int kPixels = TX * TY;
  static constexpr int kMTiles = (kPixels + 127) / 128;
  // A tile: one 128-byte row per pixel.  The MMA reads whole 128-row M tiles; rows >= kPixels alias whatever follows the
  // A tile in shared memory (their accumulator rows are never read).
  static constexpr uint32_t kABytes = kMTiles == 2 ? 256 * 128 : ((kPixels * 128 + 1023) / 1024) * 1024;
  static_assert(TX % 5 == 0 && TY % 5 == 0 && kWarps == 4 * kMTiles, "one epilogue warp per 32 accumulator rows");
};

struct DwPwParams {
  CUtensorMap tm_in;   // x as {C, W, H, N}, box {64, TW, TH, 1}, no swizzle
  CUtensorMap tm_w;    // W2 packed [rows][K_packed] fp16, box {64, tile_n}, SWIZZLE_128B (same packing as conv1x1)
  const float* dw_w;   // [k][k][C] fp32 (ops.pack_dw)
  const float* dw_b;   // [C]
  const float* pw_b;   // [tile_n]
  __half* out;
  int32_t out_ld;
  int32_t H, W, C, N, tile_n, tiles_x;
  int32_t act1, act2;
  int32_t tmem_cols;
  uint32_t idesc;
};

template <int K, int kCtas, int kFwTX, int kFwTY>
__global__ void __launch_bounds__(FwTile<kFwTX, kFwTY>::kThreads, kCtas) dwpw_kernel(const __grid_constant__ DwPwParams p) {
  using Tile = FwTile<kFwTX, kFwTY>;
  constexpr int kFwThreads = Tile::kThreads;
  constexpr int P = K / 2;
  constexpr int TW = kFwTX + K - 1, TH = kFwTY + K - 1;
  constexpr uint32_t kHaloBytes = TH * TW * kFwCB * 2;
  constexpr uint32_t kABytes = Tile::kABytes;

  extern __shared__ uint8_t smem_fw_raw[];
  uint8_t* smem = smem_fw_raw + ((1024u - (smem_u32(smem_fw_raw) & 1023u)) & 1023u);
  uint8_t* s_a = smem;                                    // [rows][128 B], SWIZZLE_128B K-major
  uint8_t* s_w = s_a + kABytes;                           // W2: two slots of one k block [tile_n rows][128 B]
  const int kblocks = (p.C + kFwCB - 1) / kFwCB;
  const int b_bytes = p.tile_n * 128;
  __half* s_in = reinterpret_cast<__half*>(s_w + 2 * b_bytes);  // halo tile [TH][TW][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_in) + ((kHaloBytes + 127) / 128) * 128);
  uint64_t* bar_in = bars;       // halo tile + W2 block of a channel block landed (one phase per block)
  uint64_t* bar_mma = bars + 2;  // the MMAs that read the A tile / W2 slot of block cb have completed (one phase per block)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* s_pwb = reinterpret_cast<float*>(tmem_slot + 2);  // [tile_n]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_x = blockIdx.x % p.tiles_x, tile_y = blockIdx.x / p.tiles_x;
  const int img = blockIdx.y;
  const int x0 = tile_x * kFwTX, y0 = tile_y * kFwTY;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tm_in);
    tma_prefetch_desc(&p.tm_w);
    mbar_init(bar_in, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.tile_n; i += kFwThreads) s_pwb[i] = p.pw_b[i];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    pdl_launch_dependents();
    mbar_arrive_expect_tx(bar_in, kHaloBytes + b_bytes);
    tma_load_2d(s_w, &p.tm_w, bar_in, 0, 0);  // weights are constants: may start before the previous kernel ends
    pdl_wait();  // the input tile (and, causally, every output store) follows the previous kernels
    tma_load_tile_4d(s_in, &p.tm_in, bar_in, 0, x0 - P, y0 - P, img);
  }

  // this warp's 5 x 5 unit of the tile; lane = channel pair of the block
  const int uy = warp / (kFwTX / 5), ux = warp % (kFwTX / 5);
  const int oy0 = uy * 5, ox0 = ux * 5;
  const __half* ibase = s_in + (oy0 * TW + ox0) * kFwCB + 2 * lane;

  for (int cb = 0; cb < kblocks; ++cb) {
    const int c0 = cb * kFwCB;
    const bool ch_ok = c0 + 2 * lane < p.C;
    // depth-wise weights of this thread's two channels straight to registers (L2-resident, coalesced)
    float2 wreg[K * K];
#pragma unroll
    for (int t = 0; t < K * K; ++t)
      wreg[t] = ch_ok ? __ldg(reinterpret_cast<const float2*>(p.dw_w + static_cast<size_t>(t) * p.C + c0 + 2 * lane))
                      : make_float2(0.f, 0.f);
    const float2 bv = ch_ok ? __ldg(reinterpret_cast<const float2*>(p.dw_b + c0 + 2 * lane)) : make_float2(0.f, 0.f);

    mbar_wait(bar_in, cb & 1);

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
        win[j] = __half22float2(*reinterpret_cast<const __half2*>(ibase + (d * TW + j) * kFwCB));
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
    __syncthreads();  // every warp is done reading the halo tile of this block
    // the MMAs of the previous block must have finished reading the A tile (about to be overwritten) and their W2
    // slot (the one the next block's weights go to)
    if (cb > 0) mbar_wait(bar_mma, (cb - 1) & 1);
    if (threadIdx.x == 0 && cb + 1 < kblocks) {  // next block's halo tile + W2 block: overlaps the A-tile writes + MMA
      mbar_arrive_expect_tx(bar_in, kHaloBytes + b_bytes);
      tma_load_2d(s_w + ((cb + 1) & 1) * b_bytes, &p.tm_w, bar_in, c0 + kFwCB, 0);
      tma_load_tile_4d(s_in, &p.tm_in, bar_in, c0 + kFwCB, x0 - P, y0 - P, img);
    }

    // bias is already in the accumulator; act1, fp16, SW128 K-major A tile, 4 bytes per lane.  A-tile ROW of a pixel =
    // warp * 25 + (i * 5 + r): any bijection works as long as the epilogue reads the same one, and with this one the swizzle
    // phase (row & 7) = ((i * 5 + r) + warp * 25) & 7 takes one of 8 values that depend on the compile-time (i, r) and on the
    // warp only — the 8 chunk addresses are computed once per thread and every store is base + immediate (was ~5 address
    // instructions per pixel in a kernel that is issue-bound for k = 3: ncu issue slots 67 % busy).
    // Lanes beyond C hold zero weights and a zero bias, so their accumulators are exactly 0 and act1(0) = 0 needs no select.
    {
      const int row0 = warp * 25;
      uint8_t* wbase = s_a + row0 * 128 + ((lane & 3) << 2);
      uint8_t* pre[8];
#pragma unroll
      for (int k7 = 0; k7 < 8; ++k7) pre[k7] = wbase + (((lane >> 2) ^ ((k7 + row0) & 7)) << 4);
      auto write_tile = [&](auto act_c) {  // the activation is a compile-time constant inside: no switch per element
        constexpr int kAct = decltype(act_c)::value;
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            const int c = i * 5 + r;  // compile-time after unrolling
            const float a = apply_act_fast(acc[i][r].x, kAct), b2 = apply_act_fast(acc[i][r].y, kAct);
            *reinterpret_cast<uint32_t*>(pre[c & 7] + c * 128) = pack_half2(a, b2);
          }
      };
      if (p.act1 == ACT_SILU)
        write_tile(std::integral_constant<int, ACT_SILU>{});
      else if (p.act1 == ACT_RELU)
        write_tile(std::integral_constant<int, ACT_RELU>{});
      else
        write_tile(std::integral_constant<int, ACT_NONE>{});
    }
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (threadIdx.x == 0) {
      const uint64_t db = umma_smem_desc_sw128(smem_u32(s_w + (cb & 1) * b_bytes));  // landed with this block's halo tile
#pragma unroll
      for (int mt = 0; mt < Tile::kMTiles; ++mt) {
        const uint64_t da = umma_smem_desc_sw128(smem_u32(s_a + mt * 16384));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tmem_base + mt * p.tile_n, da + 2 * k, db + 2 * k, p.idesc, (cb | k) != 0 ? 1u : 0u);
      }
      tc_commit(bar_mma);
    }
  }

  // ---- epilogue: accumulator row = tile pixel; warps 0-3 -> rows 0..127, warps 4-7 (if any) -> rows 128..255 ------
  mbar_wait(bar_mma, (kblocks - 1) & 1);
  tc_fence_after_sync();
  {
    const int mt = warp >> 2, quarter = warp & 3;
    // accumulator row R = unit (R / 25) pixel (R % 25) (see the A-tile write): unit u sits at (u / (TX / 5), u % (TX / 5))
    const int R = mt * 128 + quarter * 32 + lane;
    const int unit = R / 25, c = R - unit * 25;
    const int py = (unit / (kFwTX / 5)) * 5 + c / 5, px = (unit % (kFwTX / 5)) * 5 + c % 5;
    const int gy = y0 + py, gx = x0 + px;
    const bool ok = R < kFwTX * kFwTY && gy < p.H && gx < p.W;
    __half* orow = p.out + ((static_cast<size_t>(img) * p.H + (ok ? gy : 0)) * p.W + (ok ? gx : 0)) * p.out_ld;
    const uint32_t taddr = tmem_base + mt * p.tile_n + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < p.tile_n; c += 16) {
      uint32_t r[16];
      __syncwarp();
      tmem_ld_32x32b_x16(taddr + c, r);
      tmem_ld_wait();
      if (ok && c < p.N) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          pk[j] = pack_half2(apply_act_fast(__uint_as_float(r[2 * j]) + s_pwb[c + 2 * j], p.act2),
                             apply_act_fast(__uint_as_float(r[2 * j + 1]) + s_pwb[c + 2 * j + 1], p.act2));
        if (c + 16 <= p.N) {
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + c), "r"(pk[0]), "r"(pk[1]),
                       "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                       : "memory");
        } else {  // N % 16 == 8
          *reinterpret_cast<uint4*>(orow + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

template <int K, int kCtas, int TX, int TY>
static int32_t launch_dwpw_as(DwPwParams& p, int n, int tiles, size_t smem, cudaStream_t st) {
  {
    static SmemOptIn opt_in;  // per device (a process-wide flag skipped the opt-in on a second GPU)
    const int32_t rc_attr = smem_opt_in(opt_in, dwpw_kernel<K, kCtas, TX, TY>, 113 * 1024, "dwpw");
    if (rc_attr) return rc_attr;
  }
  launch_pdl(dwpw_kernel<K, kCtas, TX, TY>, dim3(tiles, n), dim3(FwTile<TX, TY>::kThreads), smem, st, p);
  return check_launch("dwpw kernel launch");
}

// Shared memory of one CTA: alignment slack + A tile + two W2 slots + halo tile + barriers + bias.
template <int K, int TX, int TY>
static size_t dwpw_smem(int tile_n) {
  constexpr int TW = TX + K - 1, TH = TY + K - 1;
  return 1024 + FwTile<TX, TY>::kABytes + static_cast<size_t>(2) * tile_n * 128 +
         ((static_cast<size_t>(TH) * TW * kFwCB * 2 + 127) / 128) * 128 + 64 + static_cast<size_t>(tile_n) * 4;
}

static int dwpw_tmem_cols(int m_tiles, int tile_n) {
  int cols = 32;
  while (cols < m_tiles * tile_n) cols <<= 1;
  return cols;
}

// 0: 10 x 20 tiles, 256 threads (round 1).  1 (default): 10 x 10 tiles, 128 threads, twice the CTAs per SM.
static int dwpw_small_tiles() {
  static const int v = [] {
    const char* e = getenv("MAFB200_DWPW_SMALL");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return v;
}

template <int K, int TX, int TY>
static int32_t launch_dwpw(DwPwParams& p, const maf_tensor* src, cudaStream_t st) {
  using Tile = FwTile<TX, TY>;
  EncodeTiledFn enc = encode_tiled_fn();
  {
    constexpr int TW = TX + K - 1, TH = TY + K - 1;
    const cuuint64_t px = static_cast<cuuint64_t>(src->c_stride) * 2;
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(src->c), static_cast<cuuint64_t>(src->w),
                          static_cast<cuuint64_t>(src->h), static_cast<cuuint64_t>(src->n)};
    cuuint64_t strides[3] = {px, px * src->w, px * src->w * src->h};
    cuuint32_t box[4] = {kFwCB, static_cast<cuuint32_t>(TW), static_cast<cuuint32_t>(TH), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, src->ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MAF_E_CUDA, "dwconv_conv1x1: cuTensorMapEncodeTiled(src) failed: %d", (int)r);
  }
  p.tiles_x = ceil_div(src->w, TX);
  p.tmem_cols = dwpw_tmem_cols(Tile::kMTiles, p.tile_n);
  const int tiles = p.tiles_x * ceil_div(src->h, TY);
  const size_t smem = dwpw_smem<K, TX, TY>(p.tile_n);
  if (smem > 113 * 1024) return fail(MAF_E_ARG, "dwpw: %zu B of shared memory needed (C=%d N=%d)", smem, p.C, p.N);
  // resident CTAs per SM: shared memory (228 KB, 1 KB reserved per CTA), TMEM columns (512), registers (launch bounds)
  const int by_smem = static_cast<int>((228 * 1024) / (smem + 1024)), by_tmem = 512 / p.tmem_cols;
  const int fit = by_smem < by_tmem ? by_smem : by_tmem;
  const int n = src->n;
  if constexpr (Tile::kThreads == 128) {
    // k = 5: 124 registers -> 4 CTAs of 128 threads; k = 3: 96 registers -> 5 CTAs
    if constexpr (K == 3) {
      if (fit >= 5) return launch_dwpw_as<K, 5, TX, TY>(p, n, tiles, smem, st);
    }
    if (fit >= 4) return launch_dwpw_as<K, 4, TX, TY>(p, n, tiles, smem, st);
    return launch_dwpw_as<K, 2, TX, TY>(p, n, tiles, smem, st);
  } else {
    if constexpr (K == 3) {
      // k = 3 needs 18 weight + 14 window registers beside the 50 accumulators: it fits 80 registers, and with
      // cout <= 32 three CTAs fit the SM's shared memory and TMEM -> 24 instead of 16 warps to cover the gaps of its
      // short tap loop (round 1: dwpw family 459 -> 449 us).  MAFB200_DWPW_3CTA=0 selects the 2-CTA build.
      static const bool three = [] {
        const char* e = getenv("MAFB200_DWPW_3CTA");
        return !(e && e[0] == '0');
      }();
      if (three && fit >= 3) return launch_dwpw_as<K, 3, TX, TY>(p, n, tiles, smem, st);
    }
    return launch_dwpw_as<K, 2, TX, TY>(p, n, tiles, smem, st);
  }
}

}  // namespace mafb200

using namespace mafb200;

// dst = act2(W2 * act1(DW_k(src) + dw_bias) + pw_bias).  dw_weight fp32 [k][k][C] / dw_bias fp32 [C] as for
// mafb200_dwconv; pw_packed / pw_bias as for mafb200_conv1x1 with ONE source of C channels and cout <= 128
// (one column tile).  k in {3, 5}; C % 8 == 0; dst 32-B aligned with c_stride % 16 == 0.
extern "C" int32_t mafb200_dwconv_conv1x1(const maf_tensor* src, const float* dw_weight, const float* dw_bias, int32_t k,
                                          int32_t act1, const void* pw_packed, const float* pw_bias, int32_t act2,
                                          const maf_tensor* dst, void* stream) {
  if (!valid_f16_view(src) || !valid_f16_view(dst)) return fail(MAF_E_ARG, "dwconv_conv1x1: bad src/dst");
  if (!dw_weight || !dw_bias || !pw_packed || !pw_bias) return fail(MAF_E_ARG, "dwconv_conv1x1: null weights");
  if (!same_nhw(src, dst)) return fail(MAF_E_ARG, "dwconv_conv1x1: src/dst n/h/w differ");
  if (k != 3 && k != 5) return fail(MAF_E_ARG, "dwconv_conv1x1: kernel size %d not in {3,5}", k);
  if (src->c % 8 != 0) return fail(MAF_E_ARG, "dwconv_conv1x1: C must be a multiple of 8 (got %d)", src->c);
  if (!aligned_f16_view(src)) return fail(MAF_E_ALIGN, "dwconv_conv1x1: src must be 16-B aligned with c_stride %% 8 == 0");
  if ((reinterpret_cast<uintptr_t>(dst->ptr) & 31) || (dst->c_stride % 16) != 0 || (dst->c % 8) != 0)
    return fail(MAF_E_ALIGN, "dwconv_conv1x1: dst must be 32-B aligned, c_stride %% 16 == 0, c %% 8 == 0");
  if ((reinterpret_cast<uintptr_t>(dw_weight) & 7) || (reinterpret_cast<uintptr_t>(dw_bias) & 7) ||
      (reinterpret_cast<uintptr_t>(pw_packed) & 15))
    return fail(MAF_E_ALIGN, "dwconv_conv1x1: weight alignment");
  if (act1 < MAF_ACT_NONE || act1 > MAF_ACT_RELU || act2 < MAF_ACT_NONE || act2 > MAF_ACT_SIGMOID)
    return fail(MAF_E_ARG, "dwconv_conv1x1: bad activation");
  if (src->n > 65535) return fail(MAF_E_ARG, "dwconv_conv1x1: batch %d > 65535", src->n);
  int n_tiles = 0, tile_n = 0;
  mafb200_gemm_tiling(dst->c, &n_tiles, &tile_n);
  if (n_tiles != 1) return fail(MAF_E_ARG, "dwconv_conv1x1: cout %d > 128", dst->c);
  int32_t rc = require_sm100();
  if (rc) return rc;

  DwPwParams p;
  memset(&p, 0, sizeof(p));
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(MAF_E_ARCH, "cuTensorMapEncodeTiled entry point not available");
  {
    const int chans[1] = {src->c};
    const int k_packed = mafb200_packed_k_1x1(chans, 1);
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_packed), static_cast<cuuint64_t>(tile_n)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_packed) * 2};
    cuuint32_t box[2] = {kFwCB, static_cast<cuuint32_t>(tile_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(pw_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MAF_E_CUDA, "dwconv_conv1x1: cuTensorMapEncodeTiled(W) failed: %d", (int)r);
  }
  p.dw_w = dw_weight;
  p.dw_b = dw_bias;
  p.pw_b = pw_bias;
  p.out = static_cast<__half*>(dst->ptr);
  p.out_ld = dst->c_stride;
  p.H = src->h;
  p.W = src->w;
  p.C = src->c;
  p.N = dst->c;
  p.tile_n = tile_n;
  p.act1 = act1;
  p.act2 = act2;
  p.idesc = umma_idesc_f16(128, tile_n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dwpw_small_tiles())
    return k == 3 ? launch_dwpw<3, 10, 10>(p, src, st) : launch_dwpw<5, 10, 10>(p, src, st);
  return k == 3 ? launch_dwpw<3, 20, 10>(p, src, st) : launch_dwpw<5, 20, 10>(p, src, st);
}
