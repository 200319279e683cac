// K3b — depth-wise k x k convolution (k in {3,5,7,9}, stride 1, pad k/2) on the tensor cores.
//
// Same op as dwconv.cu (the deploy form of UniRepLKNetBlock / DilatedReparamBlock,
// yolov6/layers/common.py:2948-3100, used by DepthBottleneckUni common.py:898-927 and Head_DepthUni
// common.py:1288-1336), but the k*k taps run as warp-level HMMA instead of FFMA chains.  Measured on
// B200 (tools/ubench/pipes.cu): FFMA / FFMA2 / HFMA2 all peak at ~124 FMA/clk/SM, mma.sync m16n8k16 at
// 546 dense TFLOP/s (~940 MAC/clk/SM) — so even a formulation that wastes most of the MMA beats the
// CUDA cores for k >= 5.
//
// Formulation (per channel c, per kernel row ky): a 1-D convolution along x is a product with a banded
// Toeplitz matrix,
//     D[y][x] += sum_j A[y][j] * T_ky[j][x],   A[y][j] = in[y + ky][x0 + j],   T_ky[j][x] = w[c][ky][j - x]
// i.e. one m16n8k16 per (16 output rows, 8 output columns, ky): the 16-wide K window x0 .. x0+15 covers
// the 8 + k - 1 <= 16 inputs the 8 outputs need.  T_ky depends only on (c, ky), so the host packs it once
// as ready-to-use B fragments (fp16, 2 registers per lane): `mafb200_dw_tc_pack`.
//
// The A operand needs x contiguous per channel, activations are NHWC.  Per CTA (16 x 16 outputs x 32 ch):
//   1. 16-byte cp.async brings the (16+k-1) x 24 x 32 halo tile (zero fill = padding), chunks XOR-swizzled;
//      (a TMA tiled load of this box is bound by the TMA unit's per-row rate: one 64-B row per pixel)
//   2. ldmatrix.trans turns it into a channel-planar copy  s_pl[ch][y*24 + x]  (8 pixels x 32 channels
//      per instruction; both the reads and the 32-bit writes are bank-conflict free);
//   3. each warp takes channel PAIRS: per (x block, ky) one ldmatrix.x4 (conflict free: row pitch 48 B)
//      + one HMMA per channel, fp32 accumulators in registers;
//   4. bias + activation on the C fragments, the two channels of a pixel packed to one 32-bit word into an
//      NHWC staging tile (aliases the dead halo tile), then 256-bit coalesced global stores.
#include <string.h>

#include "common.cuh"
#include "host.h"

namespace mafb200 {

constexpr int kTcTile = 16;   // output tile is kTcTile x kTcTile pixels
constexpr int kTcTW = 24;     // halo tile width: 16 + 8 (covers k <= 9; 24 / 8 odd -> conflict-free ldmatrix rows)
constexpr int kTcCB = 32;     // channels per CTA
constexpr int kTcThreads = 256;
constexpr int kTcPixW = 17;   // staging: words per pixel (16 data + 1 pad)
constexpr int kTcRowW = 296;  // staging: words per output row (16 * 17 + 24; = 8 mod 32)

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

template <int K, int kAct>
__global__ void __launch_bounds__(kTcThreads, 3)
    dwconv_tc_kernel(const __half* __restrict__ in, int in_ld, __half* __restrict__ out, int out_ld,
                     const uint2* __restrict__ wtab, const float* __restrict__ bias, int H, int W, int C, int tiles_x) {
  constexpr int P = K / 2;
  constexpr int TH = kTcTile + K - 1;
  constexpr int NPX = TH * kTcTW;          // halo pixels
  constexpr int PP = NPX + 8;              // plane pitch (halves): 8 * odd -> conflict-free transposed writes
  constexpr uint32_t kTileBytes = NPX * kTcCB * 2;
  constexpr uint32_t kInBytes = (kTileBytes + 1023) / 1024 * 1024;
  static_assert(kTcRowW * kTcTile * 4 <= kInBytes, "staging tile must fit in the dead halo tile");
  static_assert((PP / 8) % 2 == 1, "plane pitch must be 8 * odd");

  extern __shared__ uint8_t smem_dw_raw[];
  uint8_t* smem = smem_dw_raw + ((512u - (smem_u32(smem_dw_raw) & 511u)) & 511u);  // 64-B swizzle pattern: 512 B
  uint8_t* s_in = smem;                                            // [NPX][32] fp16, 16-B chunks XOR-swizzled
  __half* s_pl = reinterpret_cast<__half*>(smem + kInBytes);       // [32][PP]
  uint32_t* s_out = reinterpret_cast<uint32_t*>(smem);             // staging [16][kTcRowW] words (aliases s_in)

  const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
  const int c0 = blockIdx.y * kTcCB;
  const int img = blockIdx.z;
  const int x0 = tile_x * kTcTile, y0 = tile_y * kTcTile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  // B fragments (constants) of this warp's first channel pair: in flight while the tile arrives
  const int n_pairs = min(kTcCB, C - c0) >> 1;  // C % 8 == 0
  uint2 bf0[K], bf1[K];
  float bias0 = 0.f, bias1 = 0.f;
  auto load_b = [&](int pair) {
    const uint2* wt = wtab + (static_cast<size_t>(c0 + 2 * pair) * K) * 32 + lane;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      bf0[ky] = __ldg(wt + ky * 32);
      bf1[ky] = __ldg(wt + (K + ky) * 32);
    }
    const float bscale = kAct == ACT_SILU ? 0.5f : 1.0f;
    bias0 = bscale * __ldg(bias + c0 + 2 * pair);
    bias1 = bscale * __ldg(bias + c0 + 2 * pair + 1);
  };
  if (warp < n_pairs) load_b(warp);

  // ---- phase 0: halo tile -> shared memory with 16-byte cp.async (zero fill = conv padding / channel tail).
  // A TMA tiled load of this box is bound by the TMA unit's per-row rate (one 64-byte row per halo pixel,
  // ~4 ns each per SM, tools/ubench/tma_rate.cu: 2.4 TB/s chip-wide at best); the LSU path is not.
  // Threads 0..191 = (row parity, halo x, 8-channel chunk); each iteration covers two halo rows, so every
  // per-thread quantity except the row bound is loop invariant (addresses advance by constants).
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();  // the input (and, causally, every output store) follows the previous kernels
  if (threadIdx.x < 2 * kTcTW * 4) {
    const int rsel = threadIdx.x >= kTcTW * 4 ? 1 : 0;
    const int combo = threadIdx.x - rsel * kTcTW * 4;
    const int hx = combo >> 2, j = combo & 3;
    const int gx = x0 - P + hx;
    const bool ok_x = gx >= 0 && gx < W && c0 + 8 * j < C;
    const int p = rsel * kTcTW + hx;  // halo pixel of iteration 0; (p >> 1) & 3 is the same for p + 48 i
    uint32_t dst = smem_u32(s_in) + p * 64 + ((j ^ ((p >> 1) & 3)) << 4);
    int gy = y0 - P + rsel;
    const __half* src = in + ((static_cast<size_t>(img) * H + gy) * W + gx) * in_ld + c0 + 8 * j;  // may be out of range: only dereferenced when ok
    const size_t src_step = static_cast<size_t>(2) * W * in_ld;
#pragma unroll
    for (int i = 0; i < TH / 2; ++i) {
      const bool ok = ok_x && gy >= 0 && gy < H;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + i * (2 * kTcTW * 64)), "l"(ok ? src : in),
                   "r"(ok ? 16 : 0)
                   : "memory");
      src += src_step;
      gy += 2;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- phase 1: NHWC halo tile -> channel-planar copy ----------------------------------------------
  {
    const int j = lane >> 3, r = lane & 7;  // ldmatrix source row: channel chunk j (8 ch), pixel p0 + r
    const uint32_t s_in_a = smem_u32(s_in);
#pragma unroll 2
    for (int q = warp; q < NPX / 8; q += kTcThreads / 32) {
      const int p = q * 8 + r;
      uint32_t v[4];
      ldmatrix_x4_trans(s_in_a + p * 64 + ((j ^ ((p >> 1) & 3)) << 4), v);
      // lane (g, t) now holds, for chunk jj: channel 8*jj + g, pixels q*8 + 2t, +1
      __half* dst = s_pl + g * PP + q * 8 + 2 * t;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) *reinterpret_cast<uint32_t*>(dst + jj * 8 * PP) = v[jj];
    }
  }
  __syncthreads();  // planar copy complete; the halo tile is dead from here on (staging reuses it)

  // ---- phase 2/3: Toeplitz HMMA per channel pair, epilogue into the staging tile -----------------------
  {
    // ldmatrix.x4 row address of this lane: matrix i = lane >> 3 -> rows (i & 1) * 8 + r, k-half i >> 1
    const int mi = lane >> 3, r = lane & 7;
    const uint32_t a_off = static_cast<uint32_t>((((mi & 1) * 8 + r) * kTcTW + (mi >> 1) * 8) * 2);
    const uint32_t s_pl_a = smem_u32(s_pl);
#pragma unroll 1
    for (int pair = warp; pair < n_pairs; pair += kTcThreads / 32) {
      const uint32_t pa0 = s_pl_a + static_cast<uint32_t>(2 * pair) * (PP * 2) + a_off;
      const uint32_t pa1 = pa0 + PP * 2;
      // four independent accumulator chains: (channel 0 / 1) x (x block 0 / 1)
      float acc[2][2][4];
#pragma unroll
      for (int xb = 0; xb < 2; ++xb)
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[xb][ch][i] = 0.f;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        uint32_t a00[4], a01[4], a10[4], a11[4];
        ldmatrix_x4(pa0 + (ky * kTcTW) * 2, a00);
        ldmatrix_x4(pa1 + (ky * kTcTW) * 2, a01);
        ldmatrix_x4(pa0 + (ky * kTcTW + 8) * 2, a10);
        ldmatrix_x4(pa1 + (ky * kTcTW + 8) * 2, a11);
        mma_16816(acc[0][0], a00, bf0[ky]);
        mma_16816(acc[0][1], a01, bf1[ky]);
        mma_16816(acc[1][0], a10, bf0[ky]);
        mma_16816(acc[1][1], a11, bf1[ky]);
      }
      const float2 hb = make_float2(bias0, bias1);
      // every HMMA that reads this pair's B fragments has been issued: fetch the next pair's now, the
      // epilogue below hides the latency
      const int next = pair + kTcThreads / 32;
      if (next < n_pairs) load_b(next);
      // C fragment: acc[..][0..1] = (row g, cols 2t, 2t+1), acc[..][2..3] = (row g+8, same cols)
#pragma unroll
      for (int xb = 0; xb < 2; ++xb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int y = g + (i >> 1) * 8, x = xb * 8 + 2 * t + (i & 1);
          const float2 v = make_float2(acc[xb][0][i], acc[xb][1][i]);
          float2 o;
          if (kAct == ACT_SILU) {  // hb = 0.5 * bias: h = v/2 + b/2, silu = h + h * tanh(h)
            const float2 h = ffma2(v, make_float2(0.5f, 0.5f), hb);
            o = ffma2(h, make_float2(tanh_approx(h.x), tanh_approx(h.y)), h);
          } else {
            o = fadd2(v, hb);
            if (kAct == ACT_RELU) {
              o.x = fmaxf(o.x, 0.0f);
              o.y = fmaxf(o.y, 0.0f);
            }
          }
          s_out[y * kTcRowW + x * kTcPixW + pair] = pack_half2(o.x, o.y);
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 4: staging tile -> global NHWC, 32 B (16 channels) per thread-task -------------------------
#pragma unroll
  for (int task = threadIdx.x; task < kTcTile * kTcTile * 2; task += kTcThreads) {
    const int pix = task >> 1, hf = task & 1;
    const int y = pix >> 4, x = pix & 15;
    const int gy = y0 + y, gx = x0 + x;
    const int cbase = c0 + hf * 16;
    if (gy >= H || gx >= W || cbase >= C) continue;
    const uint32_t* src = s_out + y * kTcRowW + x * kTcPixW + hf * 8;
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = src[i];
    __half* dst = out + (static_cast<size_t>(img) * H * W + static_cast<size_t>(gy) * W + gx) * out_ld + cbase;
    if (cbase + 16 <= C) {
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                   "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    } else {  // channel tail (C % 8 == 0): one 16-byte store
      *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
    }
  }
}

template <int K, int kAct>
static int32_t launch_dw_tc(const maf_tensor* src, const void* table, const float* bias, const maf_tensor* dst,
                            cudaStream_t st) {
  constexpr int TH = kTcTile + K - 1;
  constexpr int NPX = TH * kTcTW;
  constexpr int PP = NPX + 8;
  const size_t in_bytes = (static_cast<size_t>(NPX) * kTcCB * 2 + 1023) / 1024 * 1024;
  const size_t smem = 512 + in_bytes + static_cast<size_t>(kTcCB) * PP * 2;
  {
    static SmemOptIn opt_in;  // per device (a process-wide flag skipped the opt-in on a second GPU)
    const int32_t rc_attr = smem_opt_in(opt_in, dwconv_tc_kernel<K, kAct>, 100 * 1024, "dwconv_tc");
    if (rc_attr) return rc_attr;
  }
  const int tiles_x = ceil_div(src->w, kTcTile), tiles_y = ceil_div(src->h, kTcTile);
  dim3 grid(tiles_x * tiles_y, ceil_div(src->c, kTcCB), src->n);
  launch_pdl(dwconv_tc_kernel<K, kAct>, grid, dim3(kTcThreads), smem, st, static_cast<const __half*>(src->ptr),
             src->c_stride, static_cast<__half*>(dst->ptr),
             dst->c_stride, static_cast<const uint2*>(table), bias, src->h, src->w, src->c, tiles_x);
  return check_launch("dwconv_tc kernel launch");
}

template <int K>
static int32_t dispatch_dw_tc(const maf_tensor* src, const void* table, const float* bias, int act,
                              const maf_tensor* dst, cudaStream_t st) {
  switch (act) {
    case MAF_ACT_SILU: return launch_dw_tc<K, ACT_SILU>(src, table, bias, dst, st);
    case MAF_ACT_RELU: return launch_dw_tc<K, ACT_RELU>(src, table, bias, dst, st);
    default: return launch_dw_tc<K, ACT_NONE>(src, table, bias, dst, st);
  }
}

}  // namespace mafb200

using namespace mafb200;

extern "C" size_t mafb200_dw_tc_table_bytes(int32_t c, int32_t k) {
  if (c <= 0 || (k != 3 && k != 5 && k != 7 && k != 9)) return 0;
  return static_cast<size_t>(c) * k * 32 * sizeof(uint2);
}

// Host-side packer (pure CPU): weight fp32 [c][k][k] (PyTorch's [C,1,k,k]) -> per (channel, ky) the 32 lanes'
// B fragments of the Toeplitz matrix T_ky[j][n] = w[ky][j - n] (0 <= j - n < k), fp16:
//   lane (g = lane / 4, t = lane % 4):  .x = {T[2t][g], T[2t+1][g]},  .y = {T[2t+8][g], T[2t+9][g]}
extern "C" int32_t mafb200_dw_tc_pack(const float* weight, int32_t c, int32_t k, void* table) {
  if (!weight || !table || c <= 0) return fail(MAF_E_ARG, "dw_tc_pack: bad arguments");
  if (k != 3 && k != 5 && k != 7 && k != 9) return fail(MAF_E_ARG, "dw_tc_pack: kernel size %d not in {3,5,7,9}", k);
  uint32_t* out = static_cast<uint32_t*>(table);
  for (int ch = 0; ch < c; ++ch) {
    for (int ky = 0; ky < k; ++ky) {
      const float* wrow = weight + (static_cast<size_t>(ch) * k + ky) * k;
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        uint32_t regs[2];
        for (int h = 0; h < 2; ++h) {
          uint32_t packed = 0;
          for (int e = 0; e < 2; ++e) {
            const int j = 2 * t + e + 8 * h;
            const int tap = j - g;
            const float v = (tap >= 0 && tap < k) ? wrow[tap] : 0.0f;
            const __half hv = __float2half_rn(v);
            packed |= static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(&hv)) << (16 * e);
          }
          regs[h] = packed;
        }
        uint32_t* dst = out + ((static_cast<size_t>(ch) * k + ky) * 32 + lane) * 2;
        dst[0] = regs[0];
        dst[1] = regs[1];
      }
    }
  }
  return MAF_OK;
}

extern "C" int32_t mafb200_dwconv_tc(const maf_tensor* src, const void* table, const float* bias, int32_t k,
                                     int32_t act, const maf_tensor* dst, void* stream) {
  if (!valid_f16_view(src) || !valid_f16_view(dst)) return fail(MAF_E_ARG, "dwconv_tc: bad src/dst");
  if (!table || !bias) return fail(MAF_E_ARG, "dwconv_tc: null table/bias");
  if (!same_nhw(src, dst) || src->c != dst->c) return fail(MAF_E_ARG, "dwconv_tc: src/dst shape mismatch");
  if (src->c % 8 != 0) return fail(MAF_E_ARG, "dwconv_tc: channels must be a multiple of 8 (got %d); use mafb200_dwconv", src->c);
  if (!aligned_f16_view(src)) return fail(MAF_E_ALIGN, "dwconv_tc: src must be 16-B aligned with c_stride %% 8 == 0");
  if ((reinterpret_cast<uintptr_t>(dst->ptr) & 31) || (dst->c_stride % 16) != 0)
    return fail(MAF_E_ALIGN, "dwconv_tc: dst must be 32-B aligned with c_stride %% 16 == 0 (256-bit stores)");
  if ((reinterpret_cast<uintptr_t>(table) & 7) || (reinterpret_cast<uintptr_t>(bias) & 3))
    return fail(MAF_E_ALIGN, "dwconv_tc: table must be 8-B aligned, bias 4-B aligned");
  if (src->n > 65535) return fail(MAF_E_ARG, "dwconv_tc: batch %d > 65535", src->n);
  if (act != MAF_ACT_NONE && act != MAF_ACT_SILU && act != MAF_ACT_RELU) return fail(MAF_E_ARG, "dwconv_tc: bad act");
  if (k != 3 && k != 5 && k != 7 && k != 9) return fail(MAF_E_ARG, "dwconv_tc: kernel size %d not in {3,5,7,9}", k);
  int32_t rc = require_sm100();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (k) {
    case 3: return dispatch_dw_tc<3>(src, table, bias, act, dst, st);
    case 5: return dispatch_dw_tc<5>(src, table, bias, act, dst, st);
    case 7: return dispatch_dw_tc<7>(src, table, bias, act, dst, st);
    default: return dispatch_dw_tc<9>(src, table, bias, act, dst, st);
  }
}
