// K3 — depth-wise k x k convolution (k in {3,5,7,9}, stride 1, pad k/2), NHWC fp16 in/out,
// fp32 weights / bias / accumulation, fused bias + {none | SiLU}.
//
// This is the deploy form of the reference's "RepHConv": DilatedReparamBlock merged into one
// depth-wise kernel and folded with the outer BN of UniRepLKNetBlock
// (yolov6/layers/common.py:2948-3100), followed by DepthBottleneckUni's SiLU (common.py:915,923)
// or by nothing in Head_DepthUni (common.py:1328,1334).  It is the op for which the reference
// authors wanted a native large-kernel DW implementation (common.py:2601-2609).
//
// Arithmetic intensity is 4.5-37 flop/B -> CUDA cores, not tensor cores.
// Mapping: one thread = 2 channels (one half2 -> float2) x R consecutive output columns x TY
// consecutive output rows.  Lanes of a warp walk consecutive channel pairs, so every global
// load is a fully coalesced 128-B row segment; the (TY+k-1) x (R+k-1) input window is held in
// registers and reused k*k times; neighbouring threads' halos hit L1.
#include "common.cuh"
#include "host.h"

namespace mafb200 {

template <int K, int R, int TY>
__global__ void __launch_bounds__(128)
    dwconv_kernel(const __half* __restrict__ in, int in_ld, __half* __restrict__ out, int out_ld,
                  const float* __restrict__ wgt, const float* __restrict__ bias, int B, int H, int W, int C, int act) {
  constexpr int P = K / 2;
  const int cp = C >> 1;                 // channel pairs
  const int nstrip = ceil_div(W, R);
  const int nty = ceil_div(H, TY);
  const long long total = static_cast<long long>(B) * nty * nstrip * cp;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int pair = static_cast<int>(t % cp);
  long long rest = t / cp;
  const int strip = static_cast<int>(rest % nstrip);
  rest /= nstrip;
  const int ty = static_cast<int>(rest % nty);
  const int b = static_cast<int>(rest / nty);
  const int x0 = strip * R;
  const int y0 = ty * TY;
  const int ch = pair * 2;

  float2 acc[TY][R];
  {
    const float2 bv = *reinterpret_cast<const float2*>(bias + ch);
#pragma unroll
    for (int i = 0; i < TY; ++i)
#pragma unroll
      for (int r = 0; r < R; ++r) acc[i][r] = bv;
  }

  const __half* in_b = in + static_cast<size_t>(b) * H * W * in_ld + ch;
#pragma unroll
  for (int d = 0; d < TY + K - 1; ++d) {
    const int iy = y0 - P + d;
    if (iy < 0 || iy >= H) continue;
    float2 win[R + K - 1];
    const __half* in_row = in_b + static_cast<size_t>(iy) * W * in_ld;
#pragma unroll
    for (int j = 0; j < R + K - 1; ++j) {
      const int ix = x0 - P + j;
      if (ix >= 0 && ix < W) {
        const __half2 hv = __ldg(reinterpret_cast<const __half2*>(in_row + static_cast<size_t>(ix) * in_ld));
        win[j] = __half22float2(hv);
      } else {
        win[j] = make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < TY; ++i) {
      const int ky = d - i;  // compile-time after unrolling
      if (ky < 0 || ky >= K) continue;
      float2 wv[K];
#pragma unroll
      for (int kx = 0; kx < K; ++kx)
        wv[kx] = __ldg(reinterpret_cast<const float2*>(wgt + static_cast<size_t>(ky * K + kx) * C + ch));
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          acc[i][r].x = fmaf(win[r + kx].x, wv[kx].x, acc[i][r].x);
          acc[i][r].y = fmaf(win[r + kx].y, wv[kx].y, acc[i][r].y);
        }
      }
    }
  }

  __half* out_b = out + static_cast<size_t>(b) * H * W * out_ld + ch;
#pragma unroll
  for (int i = 0; i < TY; ++i) {
    const int y = y0 + i;
    if (y >= H) continue;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int x = x0 + r;
      if (x >= W) continue;
      const float a = apply_act_fast(acc[i][r].x, act);
      const float c = apply_act_fast(acc[i][r].y, act);
      *reinterpret_cast<__half2*>(out_b + (static_cast<size_t>(y) * W + x) * out_ld) = __floats2half2_rn(a, c);
    }
  }
}

template <int K, int R, int TY>
static int32_t launch_dw(const maf_tensor* src, const float* w, const float* bias, int act, const maf_tensor* dst,
                         cudaStream_t st) {
  const long long total = static_cast<long long>(src->n) * ceil_div(src->h, TY) * ceil_div(src->w, R) * (src->c / 2);
  const unsigned blocks = static_cast<unsigned>((total + 127) / 128);
  dwconv_kernel<K, R, TY><<<blocks, 128, 0, st>>>(static_cast<const __half*>(src->ptr), src->c_stride,
                                                  static_cast<__half*>(dst->ptr), dst->c_stride, w, bias, src->n,
                                                  src->h, src->w, src->c, act);
  return check_launch("dwconv kernel launch");
}

template <int K>
static int32_t dispatch_dw(const maf_tensor* src, const float* w, const float* bias, int act, const maf_tensor* dst,
                           cudaStream_t st) {
  // Column strip R: 4 divides every map width of the 640x640 pyramid (160/80/40/20) and keeps the
  // register window (R+K-1) small; 5-wide strips for widths that are multiples of 5 but not 4.
  if (src->w % 4 == 0 || src->w % 5 != 0) return launch_dw<K, 4, 4>(src, w, bias, act, dst, st);
  return launch_dw<K, 5, 4>(src, w, bias, act, dst, st);
}

}  // namespace mafb200

using namespace mafb200;

extern "C" int32_t mafb200_dwconv(const maf_tensor* src, const float* weight, const float* bias, int32_t k,
                                  int32_t act, const maf_tensor* dst, void* stream) {
  if (!valid_f16_view(src) || !valid_f16_view(dst)) return fail(MAF_E_ARG, "dwconv: bad src/dst");
  if (!weight || !bias) return fail(MAF_E_ARG, "dwconv: null weight/bias");
  if (!same_nhw(src, dst) || src->c != dst->c) return fail(MAF_E_ARG, "dwconv: src/dst shape mismatch");
  if ((src->c & 1) || (src->c_stride & 1) || (dst->c_stride & 1) || (reinterpret_cast<uintptr_t>(src->ptr) & 3) ||
      (reinterpret_cast<uintptr_t>(dst->ptr) & 3) || (reinterpret_cast<uintptr_t>(weight) & 7) ||
      (reinterpret_cast<uintptr_t>(bias) & 7))
    return fail(MAF_E_ALIGN, "dwconv: channels / strides must be even, pointers 4-B (data) and 8-B (weights) aligned");
  if (act != MAF_ACT_NONE && act != MAF_ACT_SILU && act != MAF_ACT_RELU) return fail(MAF_E_ARG, "dwconv: bad act");
  int32_t rc = require_sm100();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (k) {
    case 3: return dispatch_dw<3>(src, weight, bias, act, dst, st);
    case 5: return dispatch_dw<5>(src, weight, bias, act, dst, st);
    case 7: return dispatch_dw<7>(src, weight, bias, act, dst, st);
    case 9: return dispatch_dw<9>(src, weight, bias, act, dst, st);
    default: return fail(MAF_E_ARG, "dwconv: kernel size %d not in {3,5,7,9}", k);
  }
}
