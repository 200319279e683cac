// K9 — post-NMS box rescaling on the device (the step right after the hot path).
//
// Replaces, for a whole padded batch [B, max_det, 6] at once and without host round trips:
//   Evaler.scale_coords           yolov6/core/evaler.py:391-418   (x - pad) / gain, clamp to the original image
//   Evaler.box_convert + top-left yolov6/core/evaler.py:382-390,428-429  xyxy -> (x_tl, y_tl, w, h) for COCO json
//   Inferer.rescale               yolov6/core/inferer.py:181-195  (the same arithmetic with one ratio)
// plus the category-id lookup ids[int(cls)] of evaler.py:433.  All arithmetic is fp32 in the reference's
// operation order with explicit round-to-nearest intrinsics (no FMA contraction), so the result is
// bit-identical to torch on the CPU; `recip_mul` selects torch-CUDA's scalar-division form (x * (1/gain)).
// The reference's python loop does one .tolist() / .item() per detection (evaler.py:430-441).
#include "common.cuh"
#include "host.h"

namespace mafb200 {

struct ScaleParams {
  const float* det;
  const int32_t* count;
  const float* params;  // [B][6] = gain_x, gain_y, pad_x, pad_y, w0, h0
  const int32_t* category_ids;
  float* out;
  int32_t* out_cat;
  int32_t batch, max_det, nc, mode, recip_mul;
};

__device__ __forceinline__ float scale_one(float v, float pad, float gain, float inv_gain, float hi, int recip_mul) {
  const float s = __fsub_rn(v, pad);
  const float q = recip_mul ? __fmul_rn(s, inv_gain) : __fdiv_rn(s, gain);
  return fminf(fmaxf(q, 0.0f), hi);  // torch clamp_(0, hi): max then min
}

__global__ void __launch_bounds__(128) scale_detections_kernel(const ScaleParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.max_det) return;
  const size_t row = (static_cast<size_t>(b) * p.max_det + i) * 6;
  float o[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int cat = -1;
  if (i < p.count[b]) {
    const float* pr = p.params + b * 6;
    const float gx = pr[0], gy = pr[1], px = pr[2], py = pr[3], w0 = pr[4], h0 = pr[5];
    const float igx = __fdiv_rn(1.0f, gx), igy = __fdiv_rn(1.0f, gy);
    const float* d = p.det + row;
    const float x1 = scale_one(d[0], px, gx, igx, w0, p.recip_mul);
    const float y1 = scale_one(d[1], py, gy, igy, h0, p.recip_mul);
    const float x2 = scale_one(d[2], px, gx, igx, w0, p.recip_mul);
    const float y2 = scale_one(d[3], py, gy, igy, h0, p.recip_mul);
    if (p.mode == 0) {
      o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
    } else {
      // box_convert (evaler.py:385-388) then bboxes[:, :2] -= bboxes[:, 2:] / 2 (evaler.py:429)
      const float cx = __fdiv_rn(__fadd_rn(x1, x2), 2.0f), cy = __fdiv_rn(__fadd_rn(y1, y2), 2.0f);
      const float w = __fsub_rn(x2, x1), h = __fsub_rn(y2, y1);
      o[0] = __fsub_rn(cx, __fdiv_rn(w, 2.0f));
      o[1] = __fsub_rn(cy, __fdiv_rn(h, 2.0f));
      o[2] = w;
      o[3] = h;
    }
    o[4] = d[4];
    o[5] = d[5];
    const int c = static_cast<int>(d[5]);
    cat = (p.category_ids != nullptr && c >= 0 && c < p.nc) ? p.category_ids[c] : c;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) p.out[row + k] = o[k];
  if (p.out_cat != nullptr) p.out_cat[static_cast<size_t>(b) * p.max_det + i] = cat;
}

}  // namespace mafb200

using namespace mafb200;

extern "C" int32_t mafb200_scale_detections(const float* det, const int32_t* count, int32_t batch, int32_t max_det,
                                            const float* params, const int32_t* category_ids, int32_t nc, int32_t mode,
                                            int32_t recip_mul, float* out, int32_t* out_cat, void* stream) {
  if (!det || !count || !params || !out) return fail(MAF_E_ARG, "scale_detections: null pointer");
  if (batch <= 0 || max_det <= 0) return fail(MAF_E_ARG, "scale_detections: bad shape B=%d max_det=%d", batch, max_det);
  if (batch > 65535) return fail(MAF_E_ARG, "scale_detections: batch %d > 65535", batch);
  if (mode != 0 && mode != 1) return fail(MAF_E_ARG, "scale_detections: mode must be 0 (xyxy) or 1 (COCO xywh)");
  if (category_ids != nullptr && nc <= 0) return fail(MAF_E_ARG, "scale_detections: nc=%d with a category table", nc);
  int32_t rc = require_sm100();
  if (rc) return rc;
  ScaleParams p;
  p.det = det;
  p.count = count;
  p.params = params;
  p.category_ids = category_ids;
  p.out = out;
  p.out_cat = out_cat;
  p.batch = batch;
  p.max_det = max_det;
  p.nc = nc;
  p.mode = mode;
  p.recip_mul = recip_mul != 0;
  launch_pdl(scale_detections_kernel, dim3(ceil_div(max_det, 128), batch), dim3(128), 0,
             static_cast<cudaStream_t>(stream), p);
  return check_launch("scale_detections kernel launch");
}
