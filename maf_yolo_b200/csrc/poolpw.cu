// K5b — 2x2 stride-2 max pool FUSED with the 1x1 conv that consumes it (the first branch of MPRep,
// yolov6/layers/common.py:787-792: `conv1(mp(x))` with MP = nn.MaxPool2d(2, 2), common.py:667-673):
//
//     y = act( W * maxpool2x2(x) + b )          x: [n,h,w,C] fp16 NHWC,  y: [n,h/2,w/2,N]
//
// The pooled map never goes to HBM (it was written by one bandwidth kernel and read back by a GEMM whose whole cost
// is that read).  One CTA = 128 consecutive output pixels (linear over n, h/2, w/2), ALL channels:
//   per 64-channel block: 8 lanes x 16 B cover one pixel's block, 4 source pixels per output pixel (the access pattern
//   of maxpool2x2_kernel, 4 rows per thread = 16 independent 16-B loads in flight)  ->  hmax  ->  hand-swizzled
//   SWIZZLE_128B K-major A tile (one per channel block, all kept: C <= 256)  ->  fence.proxy.async  ->  one thread issues
//   4 tcgen05.mma (128 x N x 16) against that block's slice of W (all blocks fetched by TMA up front: weights are
//   constants, so they are requested before griddepcontrol.wait).
// Epilogue: accumulator row = pixel; the two warp groups split the columns; bias + act, 256-bit stores.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"

namespace mafb200 {

constexpr int kPpRows = 128;
constexpr int kPpCB = 64;
constexpr int kPpThreads = 256;
constexpr int kPpMaxBlocks = 4;

struct PoolPwParams {
  CUtensorMap tm_w;  // W packed [rows][K_packed] fp16, box {64, tile_n}, SWIZZLE_128B (same packing as conv1x1)
  const __half* in;
  const float* bias;  // [tile_n]
  __half* out;
  int32_t in_ld, out_ld;
  int32_t H, W, C, N, tile_n, kblocks;
  int32_t Ho, Wo;
  int64_t M;  // n * Ho * Wo
  int32_t act;
  int32_t tmem_cols;
  uint32_t idesc;
};

__device__ __forceinline__ uint4 pp_hmax8(uint4 a, uint4 b) {
  uint4 r;
  *reinterpret_cast<__half2*>(&r.x) = __hmax2(*reinterpret_cast<__half2*>(&a.x), *reinterpret_cast<__half2*>(&b.x));
  *reinterpret_cast<__half2*>(&r.y) = __hmax2(*reinterpret_cast<__half2*>(&a.y), *reinterpret_cast<__half2*>(&b.y));
  *reinterpret_cast<__half2*>(&r.z) = __hmax2(*reinterpret_cast<__half2*>(&a.z), *reinterpret_cast<__half2*>(&b.z));
  *reinterpret_cast<__half2*>(&r.w) = __hmax2(*reinterpret_cast<__half2*>(&a.w), *reinterpret_cast<__half2*>(&b.w));
  return r;
}

__global__ void __launch_bounds__(kPpThreads) poolpw_kernel(const __grid_constant__ PoolPwParams p) {
  extern __shared__ uint8_t smem_pp_raw[];
  uint8_t* smem = smem_pp_raw + ((1024u - (smem_u32(smem_pp_raw) & 1023u)) & 1023u);
  const int b_bytes = p.tile_n * 128;
  uint8_t* s_a = smem;                                // [kblocks][128 rows][128 B], SWIZZLE_128B K-major
  uint8_t* s_w = s_a + p.kblocks * (kPpRows * 128);   // [kblocks][tile_n rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + p.kblocks * b_bytes);
  uint64_t* bar_w = bars;        // every W block landed
  uint64_t* bar_mma = bars + 1;  // every MMA completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kPpRows;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tm_w);
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.tile_n; i += kPpThreads) s_bias[i] = p.bias[i];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    pdl_launch_dependents();
    mbar_arrive_expect_tx(bar_w, p.kblocks * b_bytes);
    for (int kb = 0; kb < p.kblocks; ++kb) tma_load_2d(s_w + kb * b_bytes, &p.tm_w, bar_w, kb * kPpCB, 0);
  }
  pdl_wait();  // the input (and, causally, every output store) follows the previous kernels

  // thread -> 16-byte chunk (8 channels) of the block, rows rg, rg + 32, rg + 64, rg + 96 of the tile
  const int chunk = threadIdx.x & 7, rg = threadIdx.x >> 3;
  const __half* src[4];
  bool row_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t m = m0 + rg + 32 * j;
    row_ok[j] = m < p.M;
    const int64_t mm = row_ok[j] ? m : 0;
    const int ox = static_cast<int>(mm % p.Wo);
    const int64_t t = mm / p.Wo;
    const int oy = static_cast<int>(t % p.Ho);
    const int64_t b = t / p.Ho;
    src[j] = p.in + ((b * p.H + 2 * oy) * p.W + 2 * ox) * p.in_ld + chunk * 8;
  }
  const size_t down = static_cast<size_t>(p.W) * p.in_ld;

  for (int kb = 0; kb < p.kblocks; ++kb) {
    const bool ch_ok = kb * kPpCB + chunk * 8 < p.C;
    uint4 v[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (ch_ok && row_ok[j]) {
        const __half* q = src[j] + kb * kPpCB;
        v[j][0] = __ldg(reinterpret_cast<const uint4*>(q));
        v[j][1] = __ldg(reinterpret_cast<const uint4*>(q + p.in_ld));
        v[j][2] = __ldg(reinterpret_cast<const uint4*>(q + down));
        v[j][3] = __ldg(reinterpret_cast<const uint4*>(q + down + p.in_ld));
      } else {
        v[j][0] = v[j][1] = v[j][2] = v[j][3] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    uint8_t* a_tile = s_a + kb * (kPpRows * 128);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = rg + 32 * j;
      *reinterpret_cast<uint4*>(a_tile + row * 128 + ((chunk ^ (row & 7)) << 4)) =
          pp_hmax8(pp_hmax8(v[j][0], v[j][1]), pp_hmax8(v[j][2], v[j][3]));
    }
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (threadIdx.x == 0) {
      if (kb == 0) mbar_wait(bar_w, 0);
      const uint64_t da = umma_smem_desc_sw128(smem_u32(a_tile));
      const uint64_t db = umma_smem_desc_sw128(smem_u32(s_w + kb * b_bytes));
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base, da + 2 * k, db + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
      if (kb + 1 == p.kblocks) tc_commit(bar_mma);
    }
  }

  // ---- epilogue: accumulator row = pixel; warps 0-3 take column chunks 0, 2, 4, ...; warps 4-7 chunks 1, 3, ... ------
  mbar_wait(bar_mma, 0);
  tc_fence_after_sync();
  {
    const int quarter = warp & 3, half = warp >> 2;
    const int64_t m = m0 + quarter * 32 + lane;
    const bool ok = m < p.M;
    __half* orow = p.out + (ok ? m : 0) * p.out_ld;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int c = half * 16; c < p.tile_n; c += 32) {
      uint32_t r[16];
      __syncwarp();
      tmem_ld_32x32b_x16(taddr + c, r);
      tmem_ld_wait();
      if (ok && c < p.N) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          pk[j] = pack_half2(apply_act_fast(__uint_as_float(r[2 * j]) + s_bias[c + 2 * j], p.act),
                             apply_act_fast(__uint_as_float(r[2 * j + 1]) + s_bias[c + 2 * j + 1], p.act));
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

}  // namespace mafb200

using namespace mafb200;

// dst = act(W * maxpool2x2(src) + bias).  packed / bias as for mafb200_conv1x1 with ONE source of C channels;
// cout <= 128 (one column tile), C <= 256, C % 8 == 0, even h and w; dst 32-B aligned with c_stride % 16 == 0.
extern "C" int32_t mafb200_maxpool2x2_conv1x1(const maf_tensor* src, const void* packed, const float* bias, int32_t act,
                                              const maf_tensor* dst, void* stream) {
  if (!valid_f16_view(src) || !valid_f16_view(dst)) return fail(MAF_E_ARG, "maxpool2x2_conv1x1: bad src/dst");
  if (!packed || !bias) return fail(MAF_E_ARG, "maxpool2x2_conv1x1: null weights");
  if ((src->h & 1) || (src->w & 1) || dst->n != src->n || dst->h != src->h / 2 || dst->w != src->w / 2)
    return fail(MAF_E_ARG, "maxpool2x2_conv1x1: need even h,w and dst [n,h/2,w/2,cout]");
  if (src->c % 8 != 0 || src->c > kPpMaxBlocks * kPpCB)
    return fail(MAF_E_ARG, "maxpool2x2_conv1x1: C must be a multiple of 8 and <= %d (got %d)", kPpMaxBlocks * kPpCB, src->c);
  if (!aligned_f16_view(src)) return fail(MAF_E_ALIGN, "maxpool2x2_conv1x1: src must be 16-B aligned with c_stride %% 8 == 0");
  if ((reinterpret_cast<uintptr_t>(dst->ptr) & 31) || (dst->c_stride % 16) != 0 || (dst->c % 8) != 0)
    return fail(MAF_E_ALIGN, "maxpool2x2_conv1x1: dst must be 32-B aligned, c_stride %% 16 == 0, c %% 8 == 0");
  if (reinterpret_cast<uintptr_t>(packed) & 15) return fail(MAF_E_ALIGN, "maxpool2x2_conv1x1: weight alignment");
  if (act < MAF_ACT_NONE || act > MAF_ACT_SIGMOID) return fail(MAF_E_ARG, "maxpool2x2_conv1x1: bad activation");
  int n_tiles = 0, tile_n = 0;
  mafb200_gemm_tiling(dst->c, &n_tiles, &tile_n);
  if (n_tiles != 1) return fail(MAF_E_ARG, "maxpool2x2_conv1x1: cout %d > 128", dst->c);
  int32_t rc = require_sm100();
  if (rc) return rc;

  PoolPwParams p;
  memset(&p, 0, sizeof(p));
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(MAF_E_ARCH, "cuTensorMapEncodeTiled entry point not available");
  {
    const int chans[1] = {src->c};
    const int k_packed = mafb200_packed_k_1x1(chans, 1);
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_packed), static_cast<cuuint64_t>(tile_n)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_packed) * 2};
    cuuint32_t box[2] = {kPpCB, static_cast<cuuint32_t>(tile_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MAF_E_CUDA, "maxpool2x2_conv1x1: cuTensorMapEncodeTiled(W) failed: %d", (int)r);
  }
  p.in = static_cast<const __half*>(src->ptr);
  p.bias = bias;
  p.out = static_cast<__half*>(dst->ptr);
  p.in_ld = src->c_stride;
  p.out_ld = dst->c_stride;
  p.H = src->h;
  p.W = src->w;
  p.C = src->c;
  p.N = dst->c;
  p.tile_n = tile_n;
  p.kblocks = ceil_div(src->c, kPpCB);
  p.Ho = dst->h;
  p.Wo = dst->w;
  p.M = static_cast<int64_t>(dst->n) * dst->h * dst->w;
  p.act = act;
  int cols = 32;
  while (cols < tile_n) cols <<= 1;
  p.tmem_cols = cols;
  p.idesc = umma_idesc_f16(128, tile_n);
  const size_t smem = 1024 + static_cast<size_t>(p.kblocks) * (kPpRows * 128 + tile_n * 128) + 32 +
                      static_cast<size_t>(tile_n) * 4;
  {
    static SmemOptIn opt_in;  // per device (a process-wide flag skipped the opt-in on a second GPU)
    const int32_t rc_attr = smem_opt_in(opt_in, poolpw_kernel, 160 * 1024, "maxpool2x2_conv1x1");
    if (rc_attr) return rc_attr;
  }
  const long long ctas = (p.M + kPpRows - 1) / kPpRows;
  launch_pdl(poolpw_kernel, dim3(static_cast<unsigned>(ctas)), dim3(kPpThreads), smem, static_cast<cudaStream_t>(stream), p);
  return check_launch("maxpool2x2_conv1x1 kernel launch");
}
