// K7 — detect-head decode: class sigmoid + DFL integral + anchor/grid-offset decode in ONE kernel.
//
// Restates, per pyramid level, the eval branch of Detect_yaml.forward
// (yolov6/models/yolo.py:355-396): softmax over the reg_max+1 bins of each box side, expectation
// with proj = 0..reg_max (yolo.py:328-330,377-378), anchor points (x+0.5, y+0.5) in grid units
// (yolov6/assigners/anchor_generator.py:11-25), dist2bbox 'xywh' (yolov6/utils/general.py:29-40),
// times the level stride (yolo.py:389), objectness column = 1 (yolo.py:393), plus the class
// sigmoid of Head_DepthUni.forward (yolov6/layers/common.py:1332).
//
// A CTA handles 64 consecutive anchors of one (image, level): 4 lanes per anchor integrate the
// four sides (17 fp16 logits each, fp32 math) and exchange l/t/r/b with warp shuffles; then the
// whole CTA streams the 64 x (5+nc) fp32 output rows — contiguous in `pred` — with coalesced
// stores, applying the sigmoid to the class logits on the way.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "host.h"
#include "nms_filter.cuh"

namespace mafb200 {

constexpr int kMaxLevels = 8;
constexpr int kAnchorsPerCta = 64;

struct DecodeParams {
  const __half* cls[kMaxLevels];
  const __half* reg[kMaxLevels];
  int32_t cls_ld[kMaxLevels], reg_ld[kMaxLevels];
  int32_t h[kMaxLevels], w[kMaxLevels];
  int32_t anchor_off[kMaxLevels];  // first anchor index of the level in the concatenated list
  int32_t cta_off[kMaxLevels + 1]; // first CTA (per image) of the level
  float stride[kMaxLevels];
  int32_t n_levels, nc, bins, total_anchors, cls_is_prob;
  float* pred;   // [B, A, 5+nc] or nullptr (detect mode: never materialised)
  // detect mode (mafb200_head_decode_detect): the threshold / compaction pass of the NMS fused in
  float* boxes;  // [B, A, 4] (cx, cy, w, h) or nullptr
  int32_t emit, multi_label;
  float conf;
  float skip_below;  // detect mode without pred: a row whose largest raw class value is below this cannot pass `> conf`
  const uint8_t* class_filter;
  int32_t* ncand;
  unsigned long long* keys;
  long long cap_pow2;
};

constexpr int kRowBuf = 160;  // floats per warp: one output row (5 + nc <= 160 in detect mode)

// kDetect = false: the reference-shaped decode (writes pred); true: detect mode (candidate filter fused in).  Two
// instantiations so that the plain path carries neither the extra shared memory nor the extra branches.
template <bool kDetect>
__global__ void __launch_bounds__(256) head_decode_kernel(const __grid_constant__ DecodeParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_box[kAnchorsPerCta][4];
  __shared__ float s_row[kDetect ? 8 : 1][kDetect ? kRowBuf : 1];  // detect mode: the row a warp just decoded, read back by the filter
  __shared__ __align__(16) __half s_reg[kAnchorsPerCta * 264];  // up to 4*(63+1)=256 (+pad) halves per anchor
  const int b = blockIdx.y;
  int lvl = 0;
  while (lvl + 1 < p.n_levels && static_cast<int>(blockIdx.x) >= p.cta_off[lvl + 1]) ++lvl;
  const int L = p.h[lvl] * p.w[lvl];
  const int a0 = (blockIdx.x - p.cta_off[lvl]) * kAnchorsPerCta;
  const int na = min(kAnchorsPerCta, L - a0);
  const int no = 5 + p.nc;
  const int reg_ld = p.reg_ld[lvl];
  const int nreg = 4 * p.bins;

  const __half* cls0 = p.cls[lvl] + (static_cast<size_t>(b) * L + a0) * p.cls_ld[lvl];
  const bool can_skip = kDetect && p.emit && p.pred == nullptr;
  // detect mode: ~99 % of the rows hold no candidate.  Decide that for all 64 rows of the CTA at once on the raw
  // logits (sigmoid is monotonic; `skip_below` sits a safety margin under logit(conf)): 4 lanes per row, every load of
  // the CTA in flight together, instead of one dependent row at a time per warp.
  __shared__ uint8_t s_skip[kDetect ? kAnchorsPerCta : 1];
  const bool fast_skip = can_skip && (p.nc & 15) == 0 && (p.cls_ld[lvl] & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(cls0) & 7) == 0;
  if (fast_skip) {
    const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
    float mz = -INFINITY;
    if (r < na) {
      const int per = p.nc >> 2;  // halves per lane, a multiple of 4
      const uint2* src = reinterpret_cast<const uint2*>(cls0 + static_cast<size_t>(r) * p.cls_ld[lvl] + q * per);
      for (int i = 0; i < (per >> 2); ++i) {
        const uint2 u = __ldg(src + i);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        mz = fmaxf(mz, fmaxf(fmaxf(a.x, a.y), fmaxf(c.x, c.y)));
      }
    }
    mz = fmaxf(mz, __shfl_xor_sync(0xffffffffu, mz, 1));
    mz = fmaxf(mz, __shfl_xor_sync(0xffffffffu, mz, 2));
    if (q == 0 && r < na) s_skip[r] = mz < p.skip_below ? 1 : 0;
    // a CTA without a single candidate row (about half of them at conf 0.03) is done: its boxes are never read
    if (__syncthreads_or(q == 0 && r < na && s_skip[r] == 0) == 0) return;
  }

  // ---- phase 0: stage the reg rows of the CTA's anchors (rows are contiguous: coalesced 16-B loads) ----
  const __half* reg_src = p.reg[lvl] + (static_cast<size_t>(b) * L + a0) * reg_ld;
  const bool fast = (reg_ld & 7) == 0 && (reinterpret_cast<uintptr_t>(reg_src) & 15) == 0 && reg_ld <= 264;
  const int pitch = fast ? reg_ld : nreg;  // smem row pitch (halves)
  {
    const __half* src = reg_src;
    if (fast) {
      const int nvec = na * reg_ld / 8;
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(s_reg);
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) d4[i] = __ldg(s4 + i);
    } else {
      for (int i = threadIdx.x; i < na * nreg; i += blockDim.x) {
        const int al = i / nreg, j = i - al * nreg;
        s_reg[al * pitch + j] = __ldg(src + static_cast<size_t>(al) * reg_ld + j);
      }
    }
  }
  __syncthreads();

  // ---- phase 1: DFL expectation per (anchor, side); 4 consecutive lanes = one anchor ------------
  {
    const int al = threadIdx.x >> 2, side = threadIdx.x & 3;
    const int a = a0 + al;
    float dist = 0.f;
    if (al < na) {
      const __half* r = s_reg + al * pitch + side * p.bins;
      float mx = -INFINITY;
      for (int i = 0; i < p.bins; ++i) mx = fmaxf(mx, __half2float(r[i]));
      float s = 0.f, e = 0.f;
      for (int i = 0; i < p.bins; ++i) {
        const float ex = __expf(__half2float(r[i]) - mx);
        s += ex;
        e = fmaf(static_cast<float>(i), ex, e);
      }
      dist = e / s;
    }
    // gather l,t,r,b of this anchor from the 4-lane group
    const unsigned base = (threadIdx.x & 31) & ~3u;
    const float dl = __shfl_sync(0xffffffffu, dist, base + 0);
    const float dt = __shfl_sync(0xffffffffu, dist, base + 1);
    const float dr = __shfl_sync(0xffffffffu, dist, base + 2);
    const float db = __shfl_sync(0xffffffffu, dist, base + 3);
    if (al < na && side == 0) {
      const int gy = a / p.w[lvl], gx = a - gy * p.w[lvl];
      const float ax = static_cast<float>(gx) + 0.5f, ay = static_cast<float>(gy) + 0.5f;
      const float x1 = ax - dl, y1 = ay - dt, x2 = ax + dr, y2 = ay + db;
      const float st = p.stride[lvl];
      s_box[al][0] = ((x1 + x2) / 2.0f) * st;
      s_box[al][1] = ((y1 + y2) / 2.0f) * st;
      s_box[al][2] = (x2 - x1) * st;
      s_box[al][3] = (y2 - y1) * st;
    }
  }
  __syncthreads();

  // ---- phase 2: one warp per output row (anchor), lanes stride the 5+nc columns: contiguous fp32
  //      stores, contiguous fp16 class-logit loads, no integer division ---------------------------------
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row0 = static_cast<size_t>(b) * p.total_anchors + p.anchor_off[lvl] + a0;
  float* dst0 = p.pred != nullptr ? p.pred + row0 * no : nullptr;
  unsigned long long* keys_b = (kDetect && p.emit) ? p.keys + static_cast<size_t>(b) * p.cap_pow2 : nullptr;
  for (int al = warp; al < na; al += 8) {
    float* dst = dst0 != nullptr ? dst0 + static_cast<size_t>(al) * no : nullptr;
    const __half* cls = cls0 + static_cast<size_t>(al) * p.cls_ld[lvl];
    if (fast_skip && s_skip[al]) continue;
    if (kDetect && p.boxes != nullptr && lane == 0)
      *reinterpret_cast<float4*>(p.boxes + (row0 + al) * 4) = make_float4(s_box[al][0], s_box[al][1], s_box[al][2], s_box[al][3]);
    if (!fast_skip && can_skip) {
      float mz = -INFINITY;
      for (int c = lane; c < p.nc; c += 32) mz = fmaxf(mz, __half2float(__ldg(cls + c)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mz = fmaxf(mz, __shfl_xor_sync(0xffffffffu, mz, o));
      if (mz < p.skip_below) continue;
    }
    for (int j = lane; j < no; j += 32) {
      float v;
      if (j < 4) {
        v = s_box[al][j];
      } else if (j == 4) {
        v = 1.0f;
      } else {
        const float z = __half2float(__ldg(cls + (j - 5)));
        v = p.cls_is_prob ? z : __fdividef(1.0f, 1.0f + __expf(-z));
      }
      if (dst != nullptr) dst[j] = v;
      if (kDetect && p.emit) s_row[warp][j] = v;
    }
    if (kDetect && p.emit) {
      // the same fp32 values the reference-shaped path would read back from `pred`, so the candidate set is identical
      __syncwarp();
      nms_filter_row(s_row[warp], p.anchor_off[lvl] + a0 + al, p.nc, p.conf, p.multi_label, p.class_filter, &p.ncand[b],
                     keys_b, p.cap_pow2, lane);
      __syncwarp();
    }
  }
}

}  // namespace mafb200

using namespace mafb200;

static int32_t decode_common(const maf_tensor* cls_logits, const maf_tensor* reg, const float* strides, int32_t n_levels,
                             int32_t reg_max, int32_t cls_is_prob, DecodeParams& p, void* stream);

extern "C" int32_t mafb200_head_decode(const maf_tensor* cls_logits, const maf_tensor* reg, const float* strides,
                                       int32_t n_levels, int32_t reg_max, int32_t cls_is_prob, float* pred,
                                       void* stream) {
  if (!pred) return fail(MAF_E_ARG, "head_decode: null pointer");
  DecodeParams p;
  memset(&p, 0, sizeof(p));
  p.pred = pred;
  return decode_common(cls_logits, reg, strides, n_levels, reg_max, cls_is_prob, p, stream);
}

// Decode + the threshold / compaction pass of non_max_suppression (yolov6/utils/nms.py:48-84) in one kernel: writes
// the (cx, cy, w, h) boxes [B, A, 4] and the unordered candidate keys + per-image counts into `workspace` (layout of
// mafb200_nms), for mafb200_nms_select.  `pred` may be NULL: the [B, A, 5+nc] tensor is then never written.
extern "C" int32_t mafb200_head_decode_detect(const maf_tensor* cls_logits, const maf_tensor* reg, const float* strides,
                                              int32_t n_levels, int32_t reg_max, int32_t cls_is_prob, float* pred,
                                              float* boxes, double conf_thres, int32_t multi_label,
                                              const uint8_t* class_filter, void* workspace, size_t workspace_bytes,
                                              void* stream) {
  if (!boxes || !workspace || !cls_logits) return fail(MAF_E_ARG, "head_decode_detect: null pointer");
  if (!(conf_thres >= 0.0 && conf_thres <= 1.0))
    return fail(MAF_E_ARG, "head_decode_detect: conf_thres must be in [0,1], got %g", conf_thres);
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(MAF_E_ALIGN, "head_decode_detect: workspace must be 256-B aligned");
  if (reinterpret_cast<uintptr_t>(boxes) & 15) return fail(MAF_E_ALIGN, "head_decode_detect: boxes must be 16-B aligned");
  if (n_levels < 1 || n_levels > kMaxLevels) return fail(MAF_E_ARG, "head_decode_detect: n_levels=%d", n_levels);
  const int nc = cls_logits[0].c, n = cls_logits[0].n;
  if (5 + nc > kRowBuf) return fail(MAF_E_ARG, "head_decode_detect: nc=%d > %d", nc, kRowBuf - 5);
  long long anchors = 0;
  for (int l = 0; l < n_levels; ++l) anchors += static_cast<long long>(cls_logits[l].h) * cls_logits[l].w;
  if (anchors * nc > 0x7fffffffll) return fail(MAF_E_ARG, "head_decode_detect: anchors*nc overflows int32");
  if (workspace_bytes < mafb200_nms_workspace_bytes(n, static_cast<int32_t>(anchors), nc))
    return fail(MAF_E_WORKSPACE, "head_decode_detect: workspace %zu < required %zu", workspace_bytes,
                mafb200_nms_workspace_bytes(n, static_cast<int32_t>(anchors), nc));
  DecodeParams p;
  memset(&p, 0, sizeof(p));
  p.pred = pred;
  p.boxes = boxes;
  p.emit = 1;
  p.multi_label = (multi_label != 0 && nc > 1) ? 1 : 0;  // nms.py:57
  p.conf = static_cast<float>(conf_thres);
  // raw-value bound under which sigmoid(z) (or z itself when the input already holds probabilities) cannot exceed conf
  if (cls_is_prob) p.skip_below = static_cast<float>(conf_thres) - 1e-3f;
  else if (conf_thres <= 0.0) p.skip_below = -INFINITY;
  else if (conf_thres >= 1.0) p.skip_below = 30.0f;
  else p.skip_below = static_cast<float>(log(conf_thres / (1.0 - conf_thres)) - 0.05);
  p.class_filter = class_filter;
  const size_t hdr = ((static_cast<size_t>(n) * 4 + 255) / 256) * 256;
  p.ncand = static_cast<int32_t*>(workspace);
  p.keys = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) + hdr);
  long long cap = 1;
  while (cap < anchors * nc) cap <<= 1;
  p.cap_pow2 = cap;
  cudaError_t e = cudaMemsetAsync(p.ncand, 0, static_cast<size_t>(n) * 4, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(MAF_E_CUDA, "head_decode_detect: cudaMemsetAsync: %s", cudaGetErrorString(e));
  return decode_common(cls_logits, reg, strides, n_levels, reg_max, cls_is_prob, p, stream);
}

static int32_t decode_common(const maf_tensor* cls_logits, const maf_tensor* reg, const float* strides, int32_t n_levels,
                             int32_t reg_max, int32_t cls_is_prob, DecodeParams& p, void* stream) {
  if (!cls_logits || !reg || !strides) return fail(MAF_E_ARG, "head_decode: null pointer");
  if (n_levels < 1 || n_levels > kMaxLevels) return fail(MAF_E_ARG, "head_decode: n_levels=%d (1..%d)", n_levels, kMaxLevels);
  if (reg_max < 1 || reg_max > 63) return fail(MAF_E_ARG, "head_decode: reg_max=%d (1..63)", reg_max);
  const int nc = cls_logits[0].c, n = cls_logits[0].n;
  int anchors = 0, ctas = 0;
  for (int l = 0; l < n_levels; ++l) {
    const maf_tensor* c = &cls_logits[l];
    const maf_tensor* r = &reg[l];
    if (!valid_f16_view(c) || !valid_f16_view(r)) return fail(MAF_E_ARG, "head_decode: bad tensor at level %d", l);
    if (c->c != nc || c->n != n || !same_nhw(c, r) || r->c != 4 * (reg_max + 1))
      return fail(MAF_E_ARG, "head_decode: level %d shape mismatch (cls c=%d reg c=%d)", l, c->c, r->c);
    p.cls[l] = static_cast<const __half*>(c->ptr);
    p.reg[l] = static_cast<const __half*>(r->ptr);
    p.cls_ld[l] = c->c_stride;
    p.reg_ld[l] = r->c_stride;
    p.h[l] = c->h;
    p.w[l] = c->w;
    p.stride[l] = strides[l];
    p.anchor_off[l] = anchors;
    p.cta_off[l] = ctas;
    anchors += c->h * c->w;
    ctas += ceil_div(c->h * c->w, kAnchorsPerCta);
  }
  p.cta_off[n_levels] = ctas;
  p.n_levels = n_levels;
  p.nc = nc;
  p.bins = reg_max + 1;
  p.total_anchors = anchors;
  p.cls_is_prob = cls_is_prob != 0;
  if (n > 65535) return fail(MAF_E_ARG, "head_decode: batch %d > 65535", n);
  int32_t rc = require_sm100();
  if (rc) return rc;
  if (p.emit)  // follows a memset node, not a kernel: plain stream-ordered launch
    launch_pdl<false>(head_decode_kernel<true>, dim3(ctas, n), dim3(256), 0, static_cast<cudaStream_t>(stream), p);
  else
    launch_pdl(head_decode_kernel<false>, dim3(ctas, n), dim3(256), 0, static_cast<cudaStream_t>(stream), p);
  return check_launch("head_decode kernel launch");
}
