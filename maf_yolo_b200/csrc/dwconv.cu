// K3 — depth-wise k x k convolution (k in {3,5,7,9}, stride 1, pad k/2), NHWC fp16 in/out,
// fp32 weights / bias / accumulation, fused bias + {none | SiLU | ReLU}.
//
// This is the deploy form of the reference's "RepHConv": DilatedReparamBlock merged into one
// depth-wise kernel and folded with the outer BN of UniRepLKNetBlock
// (yolov6/layers/common.py:2948-3100), followed by DepthBottleneckUni's SiLU (common.py:915,923)
// or by nothing in Head_DepthUni (common.py:1328,1334).  It is the op for which the reference
// authors wanted a native large-kernel DW implementation (common.py:2601-2609).
//
// Arithmetic intensity is 4.5-37 flop/B -> CUDA cores, not tensor cores; the
// kernel is FFMA-issue-bound, so everything else is kept out of the inner loop:
//   * one TMA tiled load (cp.async.bulk.tensor.4d) brings the (10+k-1) x (20+k-1) x CB halo tile of
//     a 10 x 20 output tile into shared memory; out-of-bounds zero fill IS the conv padding and the
//     channel tail, so there is not a single bounds check or address computation per tap;
//   * weights of the CB channels are staged once per CTA as fp32 [k*k][CB];
//   * a thread owns 2 channels (half2 -> float2) x 5 x 5 outputs: 25 independent float2 accumulators
//     updated with the packed FFMA2 (fma.rn.f32x2: two IEEE FMAs per issue slot, new on sm_100),
//     every shared-memory operand address is (thread base + compile-time immediate);
//   * one tile per CTA with 3-5 CTAs resident per SM measured FASTER than persistent CTAs with a
//     two-stage TMA ring (857 vs 920 us per forward): the inner loop is co-limited by the FFMA pipe and
//     shared-memory bandwidth (fp32 weights: 8 B per lane per tap), so resident warps matter most;
//   * CB = 64 channels per CTA when C % 64 == 0, else 32 (lanes then split into two 5-column strips
//     whose smem rows fall in disjoint banks because the strip width 5 is odd).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"

namespace mafb200 {

// k = 9 at 4 CTAs/SM (128 registers) spills ~80 values; at 3 CTAs/SM ptxas keeps 152 registers and caches weight rows
// across the (d, i) pairs (251 instead of 587 LDS per unit).  A/B knob for the measurement.
#ifndef MAFB200_DW9_MINB
#define MAFB200_DW9_MINB 3
#endif

constexpr int kDwR = 5;     // output columns per thread
constexpr int kDwTY = 5;    // output rows per thread
constexpr int kDwTXB = 20;  // output tile per CTA
constexpr int kDwTYB = 10;

template <int ACT>
__device__ __forceinline__ void dw_store_unit(const float2 (&acc)[kDwTY][kDwR], __half* obase, int row_ld, int out_ld,
                                              int ny, int nx) {
#pragma unroll
  for (int i = 0; i < kDwTY; ++i) {
    if (i < ny) {
      __half* orow = obase + i * row_ld;
#pragma unroll
      for (int r = 0; r < kDwR; ++r) {
        if (r < nx)
          *reinterpret_cast<__half2*>(orow + r * out_ld) =
              __floats2half2_rn(apply_act_fast(acc[i][r].x, ACT), apply_act_fast(acc[i][r].y, ACT));
      }
    }
  }
}

template <int K, int CB, bool kTma>
__global__ void __launch_bounds__(128, (CB == 32 && K == 7) ? 4 : (CB == 32 && K == 9) ? MAFB200_DW9_MINB : 1)
    dwconv_kernel(const __grid_constant__ CUtensorMap tm_in, const __half* __restrict__ in, int in_ld,
                  __half* __restrict__ out, int out_ld, const float* __restrict__ wgt,
                  const float* __restrict__ bias, int H, int W, int C, int act, int tiles_x) {
  constexpr int P = K / 2;
  constexpr int TW = kDwTXB + K - 1;        // halo tile width (pixels)
  constexpr int TH = kDwTYB + K - 1;
  constexpr int PAIRS = CB / 2;             // channel pairs per CTA
  constexpr int XS = 32 / PAIRS;            // 5-column strips a warp covers side by side
  constexpr int UX = kDwTXB / (kDwR * XS);  // warp units along x
  constexpr int UY = kDwTYB / kDwTY;
  constexpr int UNITS = UX * UY;
  constexpr uint32_t kTileBytes = TH * TW * CB * 2;

  extern __shared__ __align__(128) uint8_t smem_dw[];
  __half* s_in = reinterpret_cast<__half*>(smem_dw);                        // [TH][TW][CB]
  float* s_w = reinterpret_cast<float*>(smem_dw + kTileBytes);              // [K*K][CB]   (k >= 7 only)
  float* s_b = s_w + K * K * CB;                                            // [CB]        (k >= 7 only)
  uint64_t* bar = (K <= 5) ? reinterpret_cast<uint64_t*>(smem_dw + kTileBytes) : reinterpret_cast<uint64_t*>(s_b + CB);

  const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
  const int c0 = blockIdx.y * CB;
  const int img = blockIdx.z;
  const int x0 = tile_x * kDwTXB, y0 = tile_y * kDwTYB;

  if (kTma) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
      pdl_launch_dependents();
      pdl_wait();  // the input tile (and, causally through `bar`, every output store) follows the previous kernels
      mbar_arrive_expect_tx(bar, kTileBytes);
      tma_load_tile_4d(s_in, &tm_in, bar, c0, x0 - P, y0 - P, img);
    }
  } else {
    // Halo tile by 16-byte cp.async (zero fill = padding / channel tail).  The TMA box of this tile is one
    // 64-128-byte ROW per halo pixel and the TMA unit sustains only ~0.25 rows/ns per SM
    // (tools/ubench/tma_rate.cu), i.e. 1.3 us per 14x24 tile — as long as the tile's FFMA work itself.
    // Each thread owns fixed (halo x, 8-channel chunk) columns and walks down the rows: addresses advance
    // by constants, ~3 instructions per 16 bytes.
    if (threadIdx.x == 0) pdl_launch_dependents();
    pdl_wait();
    constexpr int CH = CB / 8;          // 16-byte chunks per pixel
    constexpr int COMBOS = TW * CH;     // chunks per halo row
    const uint32_t s_in_a = smem_u32(s_in);
#pragma unroll
    for (int cc = 0; cc < (COMBOS + 127) / 128; ++cc) {
      const int combo = threadIdx.x + cc * 128;
      if (combo < COMBOS) {
        const int hx = combo / CH, j = combo - hx * CH;
        const int gx = x0 - P + hx;
        const bool ok_x = gx >= 0 && gx < W && c0 + 8 * j < C;
        const __half* src = in + ((static_cast<size_t>(img) * H + (y0 - P)) * W + gx) * in_ld + c0 + 8 * j;
        const size_t row_step = static_cast<size_t>(W) * in_ld;
        const uint32_t dst = s_in_a + (hx * CB + 8 * j) * 2;
#pragma unroll
        for (int hy = 0; hy < TH; ++hy) {
          const int gy = y0 - P + hy;
          const bool ok = ok_x && gy >= 0 && gy < H;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + hy * (TW * CB * 2)),
                       "l"(ok ? src : in), "r"(ok ? 16 : 0)
                       : "memory");
          src += row_step;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  constexpr bool kRegW = K <= 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = lane % PAIRS;
  const int xsel = lane / PAIRS;
  const bool ch_ok = c0 + 2 * pair < C;
  // k <= 5: the thread's k*k weights (2 channels) go straight from global memory (L2-resident, coalesced
  // 8 B per lane) into registers while the TMA tile is in flight — no shared-memory staging, no weight
  // LDS in the inner loop.  k = 7 / 9 would need 98 / 162 registers: staged in shared memory instead.
  float2 wreg[kRegW ? K * K : 1];
  float2 bv = make_float2(0.f, 0.f);
  if (kRegW) {
#pragma unroll
    for (int t = 0; t < K * K; ++t)
      wreg[t] = ch_ok ? __ldg(reinterpret_cast<const float2*>(wgt + static_cast<size_t>(t) * C + c0 + 2 * pair))
                      : make_float2(0.f, 0.f);
    if (ch_ok) bv = __ldg(reinterpret_cast<const float2*>(bias + c0 + 2 * pair));
    __syncthreads();  // mbarrier init visible to all threads before they wait on it
  } else {
    // All of the thread's weight loads are issued before the first store (compile-time trip count, 8-byte loads): a
    // rolled 4-byte loop was K*K*CB/128 = 13-20 DEPENDENT L2 round trips in front of every CTA's taps.
    constexpr int kW2 = K * K * CB / 2;
    constexpr int kWIters = (kW2 + 127) / 128;
    float2 wtmp[kWIters];
#pragma unroll
    for (int it = 0; it < kWIters; ++it) {
      const int i = threadIdx.x + it * 128;
      const int t = i / (CB / 2), c = 2 * (i - t * (CB / 2));
      wtmp[it] = (i < kW2 && c0 + c < C) ? __ldg(reinterpret_cast<const float2*>(wgt + static_cast<size_t>(t) * C + c0 + c))
                                         : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < kWIters; ++it) {
      const int i = threadIdx.x + it * 128;
      if (i < kW2) *reinterpret_cast<float2*>(s_w + 2 * i) = wtmp[it];
    }
    for (int i = threadIdx.x; i < CB; i += blockDim.x) s_b[i] = (c0 + i < C) ? __ldg(bias + c0 + i) : 0.0f;
    __syncthreads();
    bv = *reinterpret_cast<const float2*>(s_b + 2 * pair);
  }
  if (kTma) {
    mbar_wait(bar, 0);
  } else {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  const float* wbase = s_w + 2 * pair;

#pragma unroll 1
  for (int u = warp; u < UNITS; u += 4) {
    const int uy = u / UX, ux = u - uy * UX;
    const int oy0 = uy * kDwTY;
    const int ox0 = (ux * XS + xsel) * kDwR;
    const __half* ibase = s_in + (oy0 * TW + ox0) * CB + 2 * pair;

    float2 acc[kDwTY][kDwR];
#pragma unroll
    for (int i = 0; i < kDwTY; ++i)
#pragma unroll
      for (int r = 0; r < kDwR; ++r) acc[i][r] = bv;

    // The (input row d, output row i) nest is expanded by template recursion: `#pragma unroll` left the k = 7 / 9
    // bodies (1225 / 2025 FFMA2) ROLLED over d, with the `ky` range tests, the weight addresses and the window
    // addresses evaluated at run time — 62 % of the executed instructions were not taps (ncu prof33).
    static_for<0, kDwTY + K - 1>([&](auto dc) {
      constexpr int d = decltype(dc)::value;
      float2 win[kDwR + K - 1];
#pragma unroll
      for (int j = 0; j < kDwR + K - 1; ++j)
        win[j] = __half22float2(*reinterpret_cast<const __half2*>(ibase + (d * TW + j) * CB));
      static_for<0, kDwTY>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int ky = d - i;
        if constexpr (ky >= 0 && ky < K) {
          float2 wv[K];
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
            wv[kx] = kRegW ? wreg[ky * K + kx] : *reinterpret_cast<const float2*>(wbase + (ky * K + kx) * CB);
#pragma unroll
          for (int r = 0; r < kDwR; ++r) {
#pragma unroll
            for (int kx = 0; kx < K; ++kx) acc[i][r] = ffma2(win[r + kx], wv[kx], acc[i][r]);
          }
        }
      });
    });

    if (ch_ok) {
      // one 64-bit base per unit, then (row, column) offsets that are small multiples of two run-time strides
      __half* obase = out + ((static_cast<size_t>(img) * H + (y0 + oy0)) * W + (x0 + ox0)) * out_ld + c0 + 2 * pair;
      const int row_ld = W * out_ld;
      const int ny = H - (y0 + oy0), nx = W - (x0 + ox0);  // rows / columns of this unit inside the image
      if (act == ACT_SILU)
        dw_store_unit<ACT_SILU>(acc, obase, row_ld, out_ld, ny, nx);
      else if (act == ACT_RELU)
        dw_store_unit<ACT_RELU>(acc, obase, row_ld, out_ld, ny, nx);
      else
        dw_store_unit<ACT_NONE>(acc, obase, row_ld, out_ld, ny, nx);
    }
  }
}

static bool dw_use_tma() {
  static const bool on = [] {
    const char* v = getenv("MAFB200_DW_TMA");  // default on: the cp.async path measured 5 % slower (722 vs 688 us / forward)
    return !(v && v[0] == '0');
  }();
  return on;
}

template <int K, int CB, bool kTma>
static int32_t launch_dw(const maf_tensor* src, const float* w, const float* bias, int act, const maf_tensor* dst,
                         cudaStream_t st) {
  constexpr int TW = kDwTXB + K - 1, TH = kDwTYB + K - 1;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(MAF_E_ARCH, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap tm;
  const cuuint64_t px = static_cast<cuuint64_t>(src->c_stride) * 2;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(src->c), static_cast<cuuint64_t>(src->w),
                        static_cast<cuuint64_t>(src->h), static_cast<cuuint64_t>(src->n)};
  cuuint64_t strides[3] = {px, px * src->w, px * src->w * src->h};
  cuuint32_t box[4] = {CB, TW, TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, src->ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MAF_E_CUDA, "dwconv: cuTensorMapEncodeTiled(c=%d w=%d h=%d n=%d) failed: %d", src->c, src->w, src->h,
                src->n, (int)r);
  const size_t smem = static_cast<size_t>(TH) * TW * CB * 2 + (K <= 5 ? 0 : static_cast<size_t>(K * K + 1) * CB * 4) + 16;
  {
    static SmemOptIn opt_in;  // per device (a process-wide flag skipped the opt-in on a second GPU)
    const int32_t rc_attr = smem_opt_in(opt_in, dwconv_kernel<K, CB, kTma>, 100 * 1024, "dwconv");
    if (rc_attr) return rc_attr;
  }
  const int tiles_x = ceil_div(src->w, kDwTXB), tiles_y = ceil_div(src->h, kDwTYB);
  dim3 grid(tiles_x * tiles_y, ceil_div(src->c, CB), src->n);
  launch_pdl(dwconv_kernel<K, CB, kTma>, grid, dim3(128), smem, st, tm, static_cast<const __half*>(src->ptr),
             src->c_stride, static_cast<__half*>(dst->ptr), dst->c_stride, w, bias, src->h, src->w, src->c, act, tiles_x);
  return check_launch("dwconv kernel launch");
}

template <int K>
static int32_t dispatch_dw(const maf_tensor* src, const float* w, const float* bias, int act, const maf_tensor* dst,
                           cudaStream_t st) {
  // 64-channel CTAs when they tile C exactly and the halo tile stays small enough for >= 3 CTAs/SM,
  // else 32-channel CTAs (<= 25 % idle lanes in the worst case, C = 72).
  static const int cb64_max_k = [] {
    const char* v = getenv("MAFB200_DW_CB64_MAXK");  // experiment knob
    return v ? atoi(v) : 5;
  }();
  if (dw_use_tma()) {
    if (src->c % 64 == 0 && K <= cb64_max_k) return launch_dw<K, 64, true>(src, w, bias, act, dst, st);
    return launch_dw<K, 32, true>(src, w, bias, act, dst, st);
  }
  if (src->c % 64 == 0 && K <= 5) return launch_dw<K, 64, false>(src, w, bias, act, dst, st);
  return launch_dw<K, 32, false>(src, w, bias, act, dst, st);
}

}  // namespace mafb200

using namespace mafb200;

extern "C" int32_t mafb200_dwconv(const maf_tensor* src, const float* weight, const float* bias, int32_t k,
                                  int32_t act, const maf_tensor* dst, void* stream) {
  if (!valid_f16_view(src) || !valid_f16_view(dst)) return fail(MAF_E_ARG, "dwconv: bad src/dst");
  if (!weight || !bias) return fail(MAF_E_ARG, "dwconv: null weight/bias");
  if (!same_nhw(src, dst) || src->c != dst->c) return fail(MAF_E_ARG, "dwconv: src/dst shape mismatch");
  if ((src->c & 1) || (dst->c_stride & 1) || (reinterpret_cast<uintptr_t>(dst->ptr) & 3))
    return fail(MAF_E_ALIGN, "dwconv: channels and dst stride must be even, dst 4-B aligned");
  if (!aligned_f16_view(src)) return fail(MAF_E_ALIGN, "dwconv: src must be 16-B aligned with c_stride %% 8 == 0 (TMA)");
  if ((reinterpret_cast<uintptr_t>(weight) & 7) || (reinterpret_cast<uintptr_t>(bias) & 7))
    return fail(MAF_E_ALIGN, "dwconv: weight / bias must be 8-B aligned");
  if (src->n > 65535) return fail(MAF_E_ARG, "dwconv: batch %d > 65535", src->n);
  if (act != MAF_ACT_NONE && act != MAF_ACT_SILU && act != MAF_ACT_RELU) return fail(MAF_E_ARG, "dwconv: bad act");
  if (k != 3 && k != 5 && k != 7 && k != 9) return fail(MAF_E_ARG, "dwconv: kernel size %d not in {3,5,7,9}", k);
  int32_t rc = require_sm100();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (k) {
    case 3: return dispatch_dw<3>(src, weight, bias, act, dst, st);
    case 5: return dispatch_dw<5>(src, weight, bias, act, dst, st);
    case 7: return dispatch_dw<7>(src, weight, bias, act, dst, st);
    default: return dispatch_dw<9>(src, weight, bias, act, dst, st);
  }
}
