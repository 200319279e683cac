// Layer 0 of every MAF-YOLO variant: RepVGGBlock(3 -> C, 3x3, stride 2) in deploy form
// (yolov6/layers/common.py:206,214-217; configs/yaml/MAF-YOLO-n.yaml:6).
//
// Reads the reference's input tensor as it is — NCHW, fp32/fp16 in [0,1] or raw uint8 (the
// `imgs.float(); imgs /= 255` of yolov6/core/evaler.py:161-163 is folded in) — and writes NHWC fp16.
// K = 27 is far too small for a tensor-core tile, and the layer is pure bandwidth
// (4.9 MB fp32 in + 4.9 MB fp16 out per image at N width), so it runs on CUDA cores:
// one thread per output pixel, 8 output channels at a time, weights broadcast from smem.
#include "common.cuh"
#include "host.h"

namespace mafb200 {

template <typename T>
__device__ __forceinline__ float load_px(const T* p);
template <>
__device__ __forceinline__ float load_px<float>(const float* p) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ float load_px<__half>(const __half* p) {
  return __half2float(__ldg(p));
}
template <>
__device__ __forceinline__ float load_px<uint8_t>(const uint8_t* p) {
  return static_cast<float>(__ldg(p)) / 255.0f;  // same fp32 division the reference performs
}

template <typename T>
__global__ void __launch_bounds__(256) stem_conv_kernel(const T* __restrict__ x, const float* __restrict__ wgt,
                                                         const float* __restrict__ bias, __half* __restrict__ out,
                                                         int n, int h, int w, int cout, int out_ld, int act) {
  extern __shared__ float s_w[];  // [27][cout_pad8] then bias[cout_pad8]
  const int cpad = round_up(cout, 8);
  float* s_b = s_w + 27 * cpad;
  for (int i = threadIdx.x; i < 27 * cpad; i += blockDim.x) {
    const int tap = i / cpad, co = i - tap * cpad;
    // wgt is [co][ky][kx][ci]; tap index here is (ci*9 + ky*3 + kx) to match the load order below
    const int ci = tap / 9, kk = tap - ci * 9;
    s_w[i] = co < cout ? wgt[(co * 9 + kk) * 3 + ci] : 0.0f;
  }
  for (int i = threadIdx.x; i < cpad; i += blockDim.x) s_b[i] = i < cout ? bias[i] : 0.0f;
  __syncthreads();

  const int ho = h >> 1, wo = w >> 1;
  const long long total = static_cast<long long>(n) * ho * wo;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = static_cast<int>(idx % wo);
  const int oy = static_cast<int>((idx / wo) % ho);
  const int b = static_cast<int>(idx / (static_cast<long long>(wo) * ho));

  float v[27];
  const T* xb = x + static_cast<size_t>(b) * 3 * h * w;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy + ky - 1;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = 2 * ox + kx - 1;
        const bool ok = iy >= 0 && iy < h && ix >= 0 && ix < w;
        v[ci * 9 + ky * 3 + kx] = ok ? load_px<T>(xb + (static_cast<size_t>(ci) * h + iy) * w + ix) : 0.0f;
      }
    }
  }

  __half* orow = out + static_cast<size_t>(idx) * out_ld;
  for (int c0 = 0; c0 < cout; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = s_b[c0 + j];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(&s_w[t * cpad + c0]);
      const float4 w1 = *reinterpret_cast<const float4*>(&s_w[t * cpad + c0 + 4]);
      acc[0] = fmaf(v[t], w0.x, acc[0]);
      acc[1] = fmaf(v[t], w0.y, acc[1]);
      acc[2] = fmaf(v[t], w0.z, acc[2]);
      acc[3] = fmaf(v[t], w0.w, acc[3]);
      acc[4] = fmaf(v[t], w1.x, acc[4]);
      acc[5] = fmaf(v[t], w1.y, acc[5]);
      acc[6] = fmaf(v[t], w1.z, acc[6]);
      acc[7] = fmaf(v[t], w1.w, acc[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = apply_act_fast(acc[j], act);
    if (c0 + 8 <= cout) {
      uint4 pk;
      pk.x = pack_half2(acc[0], acc[1]);
      pk.y = pack_half2(acc[2], acc[3]);
      pk.z = pack_half2(acc[4], acc[5]);
      pk.w = pack_half2(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(orow + c0) = pk;
    } else {
      for (int j = 0; j < 8 && c0 + j < cout; ++j) orow[c0 + j] = __float2half_rn(acc[j]);
    }
  }
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
  if (dst->c > 256) return fail(MAF_E_ARG, "stem_conv: cout %d > 256", dst->c);
  if (act < MAF_ACT_NONE || act > MAF_ACT_SIGMOID) return fail(MAF_E_ARG, "stem_conv: bad act");
  int32_t rc = require_sm100();
  if (rc) return rc;
  const int cpad = round_up(dst->c, 8);
  const size_t smem = static_cast<size_t>(28) * cpad * sizeof(float);
  const long long total = static_cast<long long>(n) * (h / 2) * (w / 2);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* out = static_cast<__half*>(dst->ptr);
  switch (x_dtype) {
    case MAF_F32:
      stem_conv_kernel<float><<<blocks, 256, smem, st>>>(static_cast<const float*>(x_nchw), weight, bias, out, n, h, w,
                                                         dst->c, dst->c_stride, act);
      break;
    case MAF_F16:
      stem_conv_kernel<__half><<<blocks, 256, smem, st>>>(static_cast<const __half*>(x_nchw), weight, bias, out, n, h,
                                                          w, dst->c, dst->c_stride, act);
      break;
    case MAF_U8:
      stem_conv_kernel<uint8_t><<<blocks, 256, smem, st>>>(static_cast<const uint8_t*>(x_nchw), weight, bias, out, n,
                                                           h, w, dst->c, dst->c_stride, act);
      break;
    default:
      return fail(MAF_E_ARG, "stem_conv: unsupported input dtype %d", x_dtype);
  }
  return check_launch("stem_conv kernel launch");
}
