// K0 — image pre-processing on the device (the step right before the hot path).
//
// One kernel replaces, per image, the reference's CPU chain
//   letterbox            yolov6/data/data_augment.py:53-83   cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder(114)
//   Inferer.precess_image yolov6/core/inferer.py:168-178     HWC -> CHW, BGR -> RGB
// and writes the uint8 NCHW tensor the stem kernel reads (the `/255` of inferer.py:176 / evaler.py:163 is
// folded into the stem conv).  The resize reproduces OpenCV's 8-bit INTER_LINEAR bit for bit
// (modules/imgproc/src/resize.cpp: 11-bit fixed-point coefficients, HResizeLinear then VResizeLinear's
// ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2): x coefficients are zeroed at the borders, y source rows are
// clipped with their weights kept — checked against cv2.resize in tests/test_preprocess_cpu.py.
// The geometry (resized size, top/left border) is computed by the host exactly as letterbox() does.
#include "common.cuh"
#include "host.h"

namespace mafb200 {

struct LetterboxParams {
  const uint8_t* src;  // [src_h][src_w][3] HWC (BGR)
  uint8_t* dst;        // [3][dst_h][dst_w] CHW; channel c of dst = channel (swap_rb ? 2 - c : c) of src
  int32_t src_h, src_w, src_pitch;
  int32_t dst_h, dst_w;
  int32_t new_h, new_w;  // resized (un-padded) size
  int32_t top, left;
  int32_t fill, swap_rb;
};

__device__ __forceinline__ void lin_coef(int d, double scale, int sn, bool zero_at_border, int& s0, int& s1, int& a0,
                                         int& a1) {
  float f = static_cast<float>((d + 0.5) * scale - 0.5);
  int s = static_cast<int>(floorf(f));
  f -= static_cast<float>(s);
  if (zero_at_border) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= sn - 1) { f = 0.f; s = sn - 1; }
  }
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));  // saturate_cast<short>: round half to even
  a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
  s1 = min(max(s + 1, 0), sn - 1);
  s0 = min(max(s, 0), sn - 1);
}

__global__ void __launch_bounds__(256) letterbox_kernel(const LetterboxParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= p.dst_w) return;
  const int rx = x - p.left, ry = y - p.top;
  int v[3] = {p.fill, p.fill, p.fill};
  if (rx >= 0 && rx < p.new_w && ry >= 0 && ry < p.new_h) {
    if (p.new_w == p.src_w && p.new_h == p.src_h) {  // data_augment.py:73: no resize when the size already fits
      const uint8_t* s = p.src + static_cast<size_t>(ry) * p.src_pitch + rx * 3;
      v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
    } else {
      int sx0, sx1, ax0, ax1, sy0, sy1, by0, by1;
      lin_coef(rx, static_cast<double>(p.src_w) / p.new_w, p.src_w, true, sx0, sx1, ax0, ax1);
      lin_coef(ry, static_cast<double>(p.src_h) / p.new_h, p.src_h, false, sy0, sy1, by0, by1);
      const uint8_t* r0 = p.src + static_cast<size_t>(sy0) * p.src_pitch;
      const uint8_t* r1 = p.src + static_cast<size_t>(sy1) * p.src_pitch;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = r0[sx0 * 3 + c] * ax0 + r0[sx1 * 3 + c] * ax1;
        const int h1 = r1[sx0 * 3 + c] * ax0 + r1[sx1 * 3 + c] * ax1;
        const int o = (((by0 * (h0 >> 4)) >> 16) + ((by1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v[c] = min(max(o, 0), 255);
      }
    }
  }
  const size_t plane = static_cast<size_t>(p.dst_h) * p.dst_w;
  uint8_t* d = p.dst + static_cast<size_t>(y) * p.dst_w + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) d[c * plane] = static_cast<uint8_t>(v[p.swap_rb ? 2 - c : c]);
}

}  // namespace mafb200

using namespace mafb200;

extern "C" int32_t mafb200_letterbox_u8(const void* src_hwc, int32_t src_h, int32_t src_w, int32_t src_pitch_bytes,
                                        void* dst_chw, int32_t dst_h, int32_t dst_w, int32_t new_h, int32_t new_w,
                                        int32_t top, int32_t left, int32_t fill, int32_t swap_rb, void* stream) {
  if (!src_hwc || !dst_chw) return fail(MAF_E_ARG, "letterbox: null pointer");
  if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || new_h <= 0 || new_w <= 0)
    return fail(MAF_E_ARG, "letterbox: bad sizes src %dx%d dst %dx%d new %dx%d", src_h, src_w, dst_h, dst_w, new_h, new_w);
  if (src_pitch_bytes < 3 * src_w) return fail(MAF_E_ARG, "letterbox: src pitch %d < 3 * width", src_pitch_bytes);
  if (top < 0 || left < 0 || top + new_h > dst_h || left + new_w > dst_w)
    return fail(MAF_E_ARG, "letterbox: resized image (%dx%d at %d,%d) does not fit %dx%d", new_h, new_w, top, left, dst_h, dst_w);
  if (fill < 0 || fill > 255) return fail(MAF_E_ARG, "letterbox: fill %d", fill);
  if (dst_h > 65535) return fail(MAF_E_ARG, "letterbox: dst_h %d > 65535", dst_h);
  int32_t rc = require_sm100();
  if (rc) return rc;
  LetterboxParams p;
  p.src = static_cast<const uint8_t*>(src_hwc);
  p.dst = static_cast<uint8_t*>(dst_chw);
  p.src_h = src_h; p.src_w = src_w; p.src_pitch = src_pitch_bytes;
  p.dst_h = dst_h; p.dst_w = dst_w; p.new_h = new_h; p.new_w = new_w;
  p.top = top; p.left = left; p.fill = fill; p.swap_rb = swap_rb != 0;
  launch_pdl(letterbox_kernel, dim3(ceil_div(dst_w, 256), dst_h), dim3(256), 0, static_cast<cudaStream_t>(stream), p);
  return check_launch("letterbox kernel launch");
}
