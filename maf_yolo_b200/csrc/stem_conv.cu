// Layer 0 of every MAF-YOLO variant: RepVGGBlock(3 -> C, 3x3, stride 2) in deploy form
// (yolov6/layers/common.py:206,214-217; configs/yaml/MAF-YOLO-n.yaml:6).
//
// Reads the reference's input tensor as it is — NCHW, fp32/fp16 in [0,1] or raw uint8 (the
// `imgs.float(); imgs /= 255` of yolov6/core/evaler.py:161-163 is folded in) — and writes NHWC fp16.
//
// K = 27 taps is too small for TMA-fed tiles but a CUDA-core version of this layer is FFMA-bound
// (648 FMA per output pixel, measured 170 us at bs32 against a 48 us HBM floor), so the multiply runs on
// the tensor core with a SOFTWARE im2col: each of the CTA's 128 threads gathers the 27 inputs of one
// output pixel, converts to fp16 and writes one row of the A operand directly in the UMMA no-swizzle
// K-major canonical layout (8x16-byte core matrices; K padded to 32); one thread issues two
// tcgen05.mma (128 x Cout x 16); the same thread-per-pixel mapping reads the TMEM row back, adds bias,
// applies the activation and stores Cout contiguous fp16.  The gather itself is 9 coalesced pair loads
// + 9 warp shuffles per thread (taps kx=1,2 are the aligned pixel pair 2*ox, 2*ox+1; tap kx=0 comes from
// the neighbouring lane).
#include "common.cuh"
#include "host.h"

namespace mafb200 {

// (Round 1 read u / 255.0f from a 256-entry shared-memory table `lut`; round 2 multiplies — see Raw<uint8_t>::cvt.  The
// parameter stays in the signature of cvt() so that the three specialisations read alike.)
// Raw loads and their conversion are separate so that ALL loads of a tile can be issued before the first use (and a
// whole tile ahead, see the kernel): Raw<T>::px = one pixel, Raw<T>::pair = the aligned pixel pair (2*ox, 2*ox+1).
template <typename T>
struct Raw;
template <>
struct Raw<float> {
  using px = float;
  using pair = float2;
  static __device__ __forceinline__ px zero_px() { return 0.f; }
  static __device__ __forceinline__ pair zero_pair() { return make_float2(0.f, 0.f); }
  static __device__ __forceinline__ px ld_px(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ pair ld_pair(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ float cvt(px v, const float*) { return v; }
  static __device__ __forceinline__ float2 cvt(pair v, const float*) { return v; }
};
template <>
struct Raw<__half> {
  using px = __half;
  using pair = __half2;
  static __device__ __forceinline__ px zero_px() { return __float2half(0.f); }
  static __device__ __forceinline__ pair zero_pair() { return __floats2half2_rn(0.f, 0.f); }
  static __device__ __forceinline__ px ld_px(const __half* p) { return __ldg(p); }
  static __device__ __forceinline__ pair ld_pair(const __half* p) { return __ldg(reinterpret_cast<const __half2*>(p)); }
  static __device__ __forceinline__ float cvt(px v, const float*) { return __half2float(v); }
  static __device__ __forceinline__ float2 cvt(pair v, const float*) { return __half22float2(v); }
};
template <>
struct Raw<uint8_t> {
  using px = uint8_t;
  using pair = uchar2;
  static __device__ __forceinline__ px zero_px() { return 0; }
  static __device__ __forceinline__ pair zero_pair() { return make_uchar2(0, 0); }
  static __device__ __forceinline__ px ld_px(const uint8_t* p) { return __ldg(p); }
  static __device__ __forceinline__ pair ld_pair(const uint8_t* p) { return __ldg(reinterpret_cast<const uchar2*>(p)); }
  // u * (1 / 255) in fp32 differs from the reference's u / 255 (evaler.py:163) in the last bit for 126 of the 256 values,
  // but the value is rounded to fp16 for the tensor core right away and THAT is identical for all 256
  // (tests/test_host_cpu.py::test_u8_scale_is_exact_in_fp16).  Two instructions per element; the 256-entry table it
  // replaces cost one shared-memory load with ~3-way bank conflicts per element (uint8 input was 4 % slower than fp32).
  // float(u) without the quarter-rate I2F: 0x4B000000 | u is the float 2^23 + u, and fma(2^23 + u, r, -2^23 r) = fl(u r)
  // exactly (-2^23 r is a power-of-two multiple of r, so the only rounding is that of u r): one LOP + one FFMA per element.
  static __device__ __forceinline__ float scale(uint32_t u) {
    constexpr float r = 1.0f / 255.0f;
    return fmaf(__uint_as_float(0x4B000000u | u), r, -8388608.0f * r);
  }
  static __device__ __forceinline__ float cvt(px v, const float*) { return scale(v); }
  static __device__ __forceinline__ float2 cvt(pair v, const float*) { return make_float2(scale(v.x), scale(v.y)); }
};

constexpr int kStemK = 32;                 // 27 taps padded to 2 x UMMA_K
constexpr uint32_t kStemLBO = 128;         // bytes between the two 8-element K chunks of one core-matrix row group
constexpr uint32_t kStemSBO = 4 * 128;     // bytes between 8-row groups (4 K-chunks of 128 B each)

// No-swizzle K-major smem matrix descriptor (layout type 0): start>>4 | LBO>>4 @16 | SBO>>4 @32 | version 1 @46
__device__ __forceinline__ uint64_t umma_smem_desc_noswz(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(kStemLBO >> 4) << 16;
  d |= static_cast<uint64_t>(kStemSBO >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// byte offset of (row r, K-chunk kc) in the canonical layout above
__device__ __forceinline__ uint32_t stem_off(int r, int kc) { return (r >> 3) * kStemSBO + kc * kStemLBO + (r & 7) * 16; }

template <typename T>
__global__ void __launch_bounds__(128, 8) stem_conv_kernel(const T* __restrict__ x, const float* __restrict__ wgt,
                                                         const float* __restrict__ bias, __half* __restrict__ out,
                                                         int n, int h, int w, int cout, int tile_n, int tmem_cols,
                                                         uint32_t idesc, int out_ld, int act, int st256) {
  __shared__ __align__(128) uint8_t s_a[128 * kStemK * 2];   // 8 KB: A operand, 128 pixels x 32 k
  __shared__ __align__(128) uint8_t s_b[64 * kStemK * 2];    // 4 KB: B operand, <= 64 output channels x 32 k
  __shared__ float s_bias[64];
  __shared__ float s_lut[256];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;

  const int t = threadIdx.x;
  const int warp = t >> 5;
  const int ho = h >> 1, wo = w >> 1;
  const long long total = static_cast<long long>(n) * ho * wo;
  const long long n_tiles = (total + 127) / 128;

  if (t == 0) {
    mbar_init(&s_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&s_tmem, tmem_cols);
    tmem_relinquish();
  }
  // B operand: row = output channel, k = (ky*3 + kx)*3 + ci == the flat index of weight[co][ky][kx][ci]
  for (int i = t; i < tile_n * 4; i += 128) {
    const int co = i >> 2, kc = i & 3;
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k0 = kc * 8 + 2 * j;
      const float a = (co < cout && k0 < 27) ? __ldg(wgt + co * 27 + k0) : 0.0f;
      const float b = (co < cout && k0 + 1 < 27) ? __ldg(wgt + co * 27 + k0 + 1) : 0.0f;
      pk[j] = pack_half2(a, b);
    }
    *reinterpret_cast<uint4*>(s_b + stem_off(co, kc)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  if (t < 64) s_bias[t] = t < cout ? __ldg(bias + t) : 0.0f;
  if (sizeof(T) == 1) {
    s_lut[t] = static_cast<float>(t) / 255.0f;
    s_lut[t + 128] = static_cast<float>(t + 128) / 255.0f;
  }

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (t == 0) pdl_launch_dependents();  // after the TMEM allocation (see common.cuh)
  pdl_wait();                           // x / out may belong to kernels still running
  const uint32_t tmem_base = s_tmem;
  const uint64_t da = umma_smem_desc_noswz(smem_u32(s_a));
  const uint64_t db = umma_smem_desc_noswz(smem_u32(s_b));
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  // Persistent CTA: TMEM, barrier and the B operand are set up once; loop over 128-pixel tiles.
  // One output pixel per thread, 27 taps in (ky, kx, ci) order.  Taps kx = 1,2 of output pixel ox are the aligned pair
  // (2*ox, 2*ox+1); tap kx = 0 (pixel 2*ox-1) is the left neighbour lane's second element (warp shuffle) — 9 coalesced
  // vector loads per thread instead of 27 strided scalar ones (w is even, so the pair never crosses the row end).
  // The loads of tile i+1 are issued right after the MMA of tile i, so their latency hides behind the epilogue (ncu
  // had half of the stall samples on the first use of each load group).
  const int lane = t & 31;
  typename Raw<T>::pair pr[9];
  typename Raw<T>::px lf[9];
  bool from_lane = false;
  // Index arithmetic in 32 bits with ONE 64-bit base per tile (ncu: the kernel is issue-bound — issue slots 68-70 % busy,
  // 577 warp instructions per output pixel of which ~40 % were 64-bit address arithmetic, predicates and selects): the
  // nine (ky, ci) rows of a pixel are base + ky * w + ci * h * w elements, offsets that fit 32 bits inside one image.
  const uint32_t hw = static_cast<uint32_t>(h) * static_cast<uint32_t>(w);
  auto issue_loads = [&](long long tile) {
    const long long idx = tile * 128 + t;
    const uint32_t cidx = static_cast<uint32_t>(idx < total ? idx : total - 1);  // total < 2^31 (checked by the host)
    const uint32_t row = cidx / static_cast<uint32_t>(wo);                       // 32-bit divisions: 3x fewer instructions
    const int ox = static_cast<int>(cidx - row * static_cast<uint32_t>(wo));
    const uint32_t b = row / static_cast<uint32_t>(ho);
    const int oy = static_cast<int>(row - b * static_cast<uint32_t>(ho));
    from_lane = lane > 0 && ox > 0;  // lane-1 then holds output pixel ox-1 of the same row
    const bool need_left = lane == 0 && ox > 0;  // ox == 0: the tap is padding
    // rows 2*oy-1, 2*oy, 2*oy+1: only the first can be above the image (oy == 0), only the last below it (h is even)
    const T* base = x + static_cast<size_t>(b) * (3u * hw) + (static_cast<uint32_t>(2 * oy) * static_cast<uint32_t>(w) + 2u * ox);
    const bool top_ok = oy > 0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const bool row_ok = ky != 0 || top_ok;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        // (ky - 1) * w + ci * hw relative to the pixel pair of the centre row; a masked-off row reads the centre row's address
        const int off = (row_ok ? (ky - 1) * w : 0) + static_cast<int>(ci * hw);
        const T* rowp = base + off;
        pr[ky * 3 + ci] = row_ok ? Raw<T>::ld_pair(rowp) : Raw<T>::zero_pair();
        lf[ky * 3 + ci] = (row_ok && need_left) ? Raw<T>::ld_px(rowp - 1) : Raw<T>::zero_px();
      }
    }
  };

  uint32_t phase = 0;
  if (blockIdx.x < n_tiles) issue_loads(blockIdx.x);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, phase ^= 1) {
  const long long idx = tile * 128 + t;
  {
    float v[kStemK];
#pragma unroll
    for (int k = 27; k < kStemK; ++k) v[k] = 0.0f;
#pragma unroll
    for (int j = 0; j < 9; ++j) {  // j = ky * 3 + ci
      const float2 p2 = Raw<T>::cvt(pr[j], s_lut);
      float left = __shfl_up_sync(0xffffffffu, p2.y, 1);
      if (!from_lane) left = Raw<T>::cvt(lf[j], s_lut);
      const int ky = j / 3, ci = j - 3 * ky;
      v[(ky * 3 + 0) * 3 + ci] = left;
      v[(ky * 3 + 1) * 3 + ci] = p2.x;
      v[(ky * 3 + 2) * 3 + ci] = p2.y;
    }
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      *reinterpret_cast<uint4*>(s_a + stem_off(t, kc)) =
          make_uint4(pack_half2(v[8 * kc], v[8 * kc + 1]), pack_half2(v[8 * kc + 2], v[8 * kc + 3]),
                     pack_half2(v[8 * kc + 4], v[8 * kc + 5]), pack_half2(v[8 * kc + 6], v[8 * kc + 7]));
    }
  }
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  if (t == 0) {
#pragma unroll
    for (int k = 0; k < kStemK / 16; ++k) {
      // 16 k-elements = two 8-element chunks = 2 * LBO bytes further: +((2*LBO) >> 4) in the address field
      tc_mma_f16(tmem_base, da + k * ((2 * kStemLBO) >> 4), db + k * ((2 * kStemLBO) >> 4), idesc, k != 0 ? 1u : 0u);
    }
    tc_commit(&s_bar);
  }
  if (tile + gridDim.x < n_tiles) issue_loads(tile + gridDim.x);  // next tile's inputs: in flight during the epilogue
  __syncwarp();
  mbar_wait(&s_bar, phase);
  tc_fence_after_sync();

  // epilogue: thread t == TMEM lane t == output pixel idx
  __half* orow = out + static_cast<size_t>(idx < total ? idx : 0) * out_ld;
#pragma unroll 1
  for (int c = 0; c < tile_n; c += 16) {
    uint32_t r[16];
    __syncwarp();
    tmem_ld_32x32b_x16(taddr + c, r);
    tmem_ld_wait();
    if (idx < total && c < cout) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        pk[j] = pack_half2(apply_act_fast(__uint_as_float(r[2 * j]) + s_bias[c + 2 * j], act),
                           apply_act_fast(__uint_as_float(r[2 * j + 1]) + s_bias[c + 2 * j + 1], act));
      if (st256 && c + 16 <= cout) {
        // 16 valid channels starting on a 32-byte boundary: one whole-sector 256-bit store
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + c), "r"(pk[0]), "r"(pk[1]),
                     "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                     : "memory");
      } else {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int c8 = c + 8 * g;
          if (c8 + 8 <= cout) {
            *reinterpret_cast<uint4*>(orow + c8) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          } else {
            for (int j = 0; j < 8; ++j) {
              if (c8 + j < cout) {
                const uint32_t w2 = pk[4 * g + (j >> 1)];
                orow[c8 + j] = __ushort_as_half(static_cast<unsigned short>((j & 1) ? (w2 >> 16) : (w2 & 0xffffu)));
              }
            }
          }
        }
      }
    }
  }

  // the next tile's gather overwrites s_a and its MMA overwrites the accumulator: all reads must be done
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  }  // tile loop

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mafb200

using namespace mafb200;

extern "C" int32_t mafb200_stem_conv3x3s2(const void* x_nchw, int32_t x_dtype, int32_t n, int32_t h, int32_t w,
                                          const float* weight, const float* bias, int32_t act, const maf_tensor* dst,
                                          void* stream) {
  if (!x_nchw || !weight || !bias) return fail(MAF_E_ARG, "stem_conv: null pointer");
  if (!valid_f16_view(dst)) return fail(MAF_E_ARG, "stem_conv: bad dst");
  if (!aligned_f16_view(dst)) return fail(MAF_E_ALIGN, "stem_conv: dst alignment");
  if (n <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1)) return fail(MAF_E_ARG, "stem_conv: need even h,w (got %dx%d)", h, w);
  if (dst->n != n || dst->h != h / 2 || dst->w != w / 2) return fail(MAF_E_ARG, "stem_conv: dst must be [n,h/2,w/2,c]");
  if (dst->c > 64) return fail(MAF_E_ARG, "stem_conv: cout %d > 64 (MAF-YOLO N/S/M use 24/32/48)", dst->c);
  if (act < MAF_ACT_NONE || act > MAF_ACT_SIGMOID) return fail(MAF_E_ARG, "stem_conv: bad act");
  {
    const uintptr_t need = x_dtype == MAF_F32 ? 7 : (x_dtype == MAF_F16 ? 3 : 1);  // pixel pairs are loaded as one vector
    if (reinterpret_cast<uintptr_t>(x_nchw) & need) return fail(MAF_E_ALIGN, "stem_conv: input pointer must be aligned to two pixels");
  }
  int32_t rc = require_sm100();
  if (rc) return rc;
  const int tile_n = round_up(dst->c, 16);
  int tmem_cols = 32;
  while (tmem_cols < tile_n) tmem_cols <<= 1;
  const uint32_t idesc = umma_idesc_f16(128, tile_n);
  const long long total = static_cast<long long>(n) * (h / 2) * (w / 2);
  if (total >= (1ll << 31)) return fail(MAF_E_ARG, "stem_conv: %lld output pixels (limit 2^31 - 1)", total);
  const long long tiles = (total + 127) / 128;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned blocks = static_cast<unsigned>(tiles < 8ll * sms ? tiles : 8ll * sms);  // persistent, 8 CTAs / SM
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* out = static_cast<__half*>(dst->ptr);
  // 256-bit stores for every full 16-channel group when pixel rows start on 32-byte boundaries
  const int st256 = ((dst->c_stride % 16) == 0 && (reinterpret_cast<uintptr_t>(dst->ptr) & 31) == 0) ? 1 : 0;
  switch (x_dtype) {
    case MAF_F32:
      launch_pdl(stem_conv_kernel<float>, dim3(blocks), dim3(128), 0, st, static_cast<const float*>(x_nchw), weight, bias, out, n, h, w,
                                                      dst->c, tile_n, tmem_cols, idesc, dst->c_stride, act, st256);
      break;
    case MAF_F16:
      launch_pdl(stem_conv_kernel<__half>, dim3(blocks), dim3(128), 0, st, static_cast<const __half*>(x_nchw), weight, bias, out, n, h, w,
                                                       dst->c, tile_n, tmem_cols, idesc, dst->c_stride, act, st256);
      break;
    case MAF_U8:
      launch_pdl(stem_conv_kernel<uint8_t>, dim3(blocks), dim3(128), 0, st, static_cast<const uint8_t*>(x_nchw), weight, bias, out, n, h,
                                                        w, dst->c, tile_n, tmem_cols, idesc, dst->c_stride, act, st256);
      break;
    default:
      return fail(MAF_E_ARG, "stem_conv: unsupported input dtype %d", x_dtype);
  }
  return check_launch("stem_conv kernel launch");
}
