// K10 — training-side kernels: label assignment + varifocal / GIoU / DFL loss, forward value AND the
// gradients w.r.t. the head outputs, for the reference's `ComputeLoss.__call__` (yolov6/models/loss.py:56-162): formal_assigner =
// TaskAlignedAssigner(topk=13, alpha=1, beta=6) (yolov6/assigners/tal_assigner.py:22-151, assigner_utils.py:25-89) and, while
// epoch_num < warmup_epoch, warmup_assigner = ATSSAssigner(9) (atss_assigner.py:17-161, iou2d_calculator.py:201-242);
// figure_iou.py:27-65, general.py:29-49.
//
// The reference materialises [B, G, 8400] float64 tensors five times over (IoU, metric, in-box mask, top-k one-hot with
// 8400 classes, ...: the "OOM RuntimeError ... CPU mode" branch of loss.py:95-133) and loops over the targets in
// python / numpy on the host.  Here nothing of size B x G x A exists except one byte map:
//
//   targets_kernel   [T,6] fp32 rows -> padded [B,G,5] float64 (class, xyxy pixels), original order kept
//   decode_kernel    DFL softmax-expectation -> boxes [B,A,4] fp32 in stride units      (4 lanes per anchor)
//   atss_select      (warm-up) one warp per (image, box): 9 closest anchor centres per level, mean + std IoU threshold -> pos
//   topk_kernel      (formal) one CTA per (image, box): IoU / in-box test / alignment metric of all A anchors into shared memory,
//                    13 rounds of block arg-max -> byte map pos[B,G,A]
//   resolve_kernel   one thread per anchor: boxes that claim it; more than one -> the box with the highest IoU over ALL
//                    boxes; atomicMax of the per-box normalisers (non-negative float64 as uint64)
//   norm_kernel      target score + class of every anchor; the last block to finish folds target_scores_sum
//   vfl_kernel       varifocal loss over B x A x nc + d/d pred_scores (4 classes per thread)
//   box_kernel       GIoU (forward-mode dual numbers) + DFL of the foreground anchors + d/d pred_distri; the last block to
//                    finish folds the partial sums in a fixed order: loss = 1.0 cls + 2.5 iou + 0.5 dfl
// Seven launches + one memset per call; sums are deterministic (fixed-order folds, no floating-point atomics).
//
// Arithmetic follows the reference's dtypes: its target tensor is FLOAT64 (built with numpy, never cast), so the IoUs, the
// metric, the normalised target scores and the loss sums are float64; the predictions, BCE and cross-entropy are fp32.
// Comparison-deciding float64 expressions use round-to-nearest intrinsics in the reference's operation order (no FMA
// contraction).  iou^6 (three multiplications), exp and log are within 2 ulp of torch's, which can move a top-13 cut only
// between metrics that are equal to ~1e-16 relative.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "host.h"

namespace mafb200 {

constexpr int kRegBins = 17;   // reg_max + 1
constexpr int kTopK = 13;
constexpr double kTalEps = 1e-9;
constexpr double kIouLossEps = 1e-10;

struct LossParams {
  const float* pred_scores;  // [B,A,nc] probabilities
  const float* pred_distri;  // [B,A,68] logits
  const float* targets;      // [T,6]
  const float* boxes_in;     // optional [B,A,4] stride units (skips decode_kernel)
  int32_t T, B, A, nc, G, img;
  int32_t n0, n1, n2;        // cells per side of the three levels (strides 8, 16, 32)
  // workspace
  double* gt;                // [B,G,5]
  uint8_t* mask_gt;          // [B,G]
  uint8_t* pos;              // [B,G,A]
  float* boxes;              // [B,A,4] stride units
  int32_t* gt_idx;           // [B,A]
  uint8_t* fg;               // [B,A]
  double* align_a;           // [B,A]
  double* ovl_a;             // [B,A]
  double* norm;              // [B,A]
  unsigned long long* pos_align;  // [B,G] float64 bits
  unsigned long long* pos_ovl;    // [B,G]
  double* partial;           // [4][kMaxPartials]: tss, cls, iou, dfl
  double* scalars;           // [8]: loss, 2.5 iou, 0.5 dfl, cls, tss, num_fg, target overflow, (spare)
  int32_t* counters;         // [4]: num_fg, target overflow, blocks done (norm stage), blocks done (box stage)
  int32_t* tlabel;           // [B,A] class of the assigned box, -1 on background anchors
  int32_t atss;              // 1: warm-up (ATSS) assigner instead of the task-aligned one
  int32_t n_vfl;             // blocks of the varifocal kernel (its partial sums are folded by the box kernel's last block)
  float* grad_scores;        // optional [B,A,nc]
  float* grad_distri;        // optional [B,A,68]
};
constexpr int kMaxPartials = 16384;  // blocks per reduction stage (the box kernel is one thread per (anchor, side): B x A <= 1M)

// ---- anchors: level-major, row-major inside a level (anchor_generator.py:29-52) ----------------------------------
struct Anchor {
  float px, py, stride;  // pixel centre, stride
};
__device__ __forceinline__ Anchor anchor_of(const LossParams& p, int a) {
  int n = p.n0, s = 8;
  if (a >= p.n0 * p.n0) {
    a -= p.n0 * p.n0;
    n = p.n1;
    s = 16;
    if (a >= p.n1 * p.n1) {
      a -= p.n1 * p.n1;
      n = p.n2;
      s = 32;
    }
  }
  const int y = a / n, x = a - y * n;
  Anchor r;
  r.stride = static_cast<float>(s);
  r.px = __fmul_rn(static_cast<float>(x) + 0.5f, r.stride);
  r.py = __fmul_rn(static_cast<float>(y) + 0.5f, r.stride);
  return r;
}

// ---- float64 helpers in the reference's operation order ------------------------------------------------------------
__device__ __forceinline__ double clip0(double v) { return v < 0.0 ? 0.0 : v; }

// assigner_utils.py:72-89: box1 = ground truth, box2 = prediction (pixels)
// The prediction's own area is an expression of fp32 tensors only and therefore fp32 in the reference; everything that
// touches the float64 ground truth is float64.
__device__ __forceinline__ double iou_gt_pred(const double* g, float fx1, float fy1, float fx2, float fy2) {
  const double px1 = fx1, py1 = fy1, px2 = fx2, py2 = fy2;
  const double x1 = fmax(g[0], px1), y1 = fmax(g[1], py1), x2 = fmin(g[2], px2), y2 = fmin(g[3], py2);
  const double overlap = __dmul_rn(clip0(__dsub_rn(x2, x1)), clip0(__dsub_rn(y2, y1)));
  const double area1 = __dmul_rn(clip0(__dsub_rn(g[2], g[0])), clip0(__dsub_rn(g[3], g[1])));
  const double area2 = __fmul_rn(fmaxf(__fsub_rn(fx2, fx1), 0.0f), fmaxf(__fsub_rn(fy2, fy1), 0.0f));
  const double uni = __dadd_rn(__dsub_rn(__dadd_rn(area1, area2), overlap), kTalEps);
  return __ddiv_rn(overlap, uni);
}

// iou2d_calculator.py:201-242 (mode 'iou'): ground truth (float64) against an ANCHOR box (fp32, a 5-stride square around the
// anchor point: anchor_generator.py:32-41) — the ATSS assigner's overlaps.  No clamp on the areas, union floored at 1e-6.
__device__ __forceinline__ double iou_gt_anchor(const double* g, float acx, float acy, float stride) {
  const float half = __fmul_rn(__fmul_rn(5.0f, stride), 0.5f);
  const float fx1 = __fsub_rn(acx, half), fy1 = __fsub_rn(acy, half), fx2 = __fadd_rn(acx, half), fy2 = __fadd_rn(acy, half);
  const double area1 = __dmul_rn(__dsub_rn(g[2], g[0]), __dsub_rn(g[3], g[1]));
  const double area2 = __fmul_rn(__fsub_rn(fx2, fx1), __fsub_rn(fy2, fy1));
  const double x1 = fmax(g[0], static_cast<double>(fx1)), y1 = fmax(g[1], static_cast<double>(fy1));
  const double x2 = fmin(g[2], static_cast<double>(fx2)), y2 = fmin(g[3], static_cast<double>(fy2));
  const double overlap = __dmul_rn(clip0(__dsub_rn(x2, x1)), clip0(__dsub_rn(y2, y1)));
  const double uni = fmax(__dsub_rn(__dadd_rn(area1, area2), overlap), 1e-6);
  return __ddiv_rn(overlap, uni);
}

// overlaps.pow(6.0) (tal_assigner.py:107) as three float64 multiplications: <= 2 ulp from a correctly rounded pow, which is
// what CUDA's pow() guarantees too, at a hundredth of its cost (pow was 45 of the top-k kernel's 57 us)
__device__ __forceinline__ double pow6(double x) {
  const double x2 = __dmul_rn(x, x);
  return __dmul_rn(__dmul_rn(x2, x2), x2);
}

// prediction box of anchor a in pixels: fp32 (stride units) x fp32 stride, as `pred_bboxes.detach() * stride_tensor`
__device__ __forceinline__ void pred_box_px(const LossParams& p, int b, int a, float stride, float* o) {
  const float4 v = *reinterpret_cast<const float4*>(p.boxes + (static_cast<size_t>(b) * p.A + a) * 4);
  o[0] = __fmul_rn(v.x, stride);
  o[1] = __fmul_rn(v.y, stride);
  o[2] = __fmul_rn(v.z, stride);
  o[3] = __fmul_rn(v.w, stride);
}

// ---- 1. targets ------------------------------------------------------------------------------------------------------
// loss.py:164-172.  Row order inside an image = order in `targets` (the assigner's tie-breaks depend on it): the slot of
// a row is the number of earlier rows of the same image (T is a few hundred at most).
__global__ void __launch_bounds__(256) loss_targets_kernel(const LossParams p) {
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < p.B * p.G; i += blockDim.x) {
    double* g = p.gt + static_cast<size_t>(i) * 5;
    g[0] = -1.0;
    g[1] = g[2] = g[3] = g[4] = 0.0;
    p.mask_gt[i] = 0;
    p.pos_align[i] = 0ull;
    p.pos_ovl[i] = 0ull;
  }
  if (threadIdx.x < 4) p.counters[threadIdx.x] = 0;
  __syncthreads();
  const double scale = static_cast<double>(p.img);
  __shared__ int16_t s_img[4096];  // image index of every row (the slot count below is O(T^2) reads)
  const bool staged = p.T <= 4096 && p.B <= 32767;  // int16 image indices
  if (staged)
    for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
      const int b = static_cast<int>(p.targets[static_cast<size_t>(t) * 6]);
      s_img[t] = static_cast<int16_t>(b < 0 || b >= p.B ? -1 : b);
    }
  __syncthreads();
  for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
    const float* r = p.targets + static_cast<size_t>(t) * 6;
    const int b = static_cast<int>(r[0]);
    if (b < 0 || b >= p.B) {
      atomicAdd(p.counters + 1, 1);
      continue;
    }
    int slot = 0;
    if (staged) {
      for (int u = 0; u < t; ++u) slot += s_img[u] == b;
    } else {
      for (int u = 0; u < t; ++u) slot += static_cast<int>(p.targets[static_cast<size_t>(u) * 6]) == b;
    }
    if (slot >= p.G) {
      atomicAdd(p.counters + 1, 1);
      continue;
    }
    const double cx = __dmul_rn(static_cast<double>(r[2]), scale), cy = __dmul_rn(static_cast<double>(r[3]), scale);
    const double w = __dmul_rn(static_cast<double>(r[4]), scale), h = __dmul_rn(static_cast<double>(r[5]), scale);
    const double x1 = __dsub_rn(cx, __dmul_rn(w, 0.5)), y1 = __dsub_rn(cy, __dmul_rn(h, 0.5));  // general.py:52-58
    const double x2 = __dadd_rn(x1, w), y2 = __dadd_rn(y1, h);
    double* g = p.gt + (static_cast<size_t>(b) * p.G + slot) * 5;
    g[0] = static_cast<double>(r[1]);
    g[1] = x1;
    g[2] = y1;
    g[3] = x2;
    g[4] = y2;
    p.mask_gt[b * p.G + slot] = __dadd_rn(__dadd_rn(__dadd_rn(x1, y1), x2), y2) > 0.0;  // loss.py:76
  }
}

// ---- 2. box decode ----------------------------------------------------------------------------------------------------
// loss.py:174-178: one lane per (anchor, side): softmax over 17 logits, expectation, point -/+ distance (stride units)
__device__ __forceinline__ float dfl_expect(const float* logit, float* prob /* [17] or nullptr */, float* lse_out) {
  float v[kRegBins];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < kRegBins; ++k) {
    v[k] = logit[k];
    m = fmaxf(m, v[k]);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kRegBins; ++k) {
    v[k] = expf(v[k] - m);
    s += v[k];
  }
  float e = 0.f;
#pragma unroll
  for (int k = 0; k < kRegBins; ++k) {
    const float q = v[k] / s;
    if (prob) prob[k] = q;
    e += q * static_cast<float>(k);
  }
  if (lse_out) *lse_out = m + logf(s);
  return e;
}

__global__ void __launch_bounds__(256) loss_decode_kernel(const LossParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, a, side)
  if (i >= static_cast<size_t>(p.B) * p.A * 4) return;
  const int side = static_cast<int>(i & 3);
  const int a = static_cast<int>((i >> 2) % p.A);
  const float e = dfl_expect(p.pred_distri + i * kRegBins, nullptr, nullptr);
  const Anchor an = anchor_of(p, a);
  const float c = (side & 1) ? an.py / an.stride : an.px / an.stride;  // anchor_points / stride_tensor (exact)
  p.boxes[i] = side < 2 ? c - e : c + e;
  if (p.grad_distri) {
    // zero d loss / d pred_distri here (68 floats = 17 float4 per anchor, spread over its four lanes): the box kernel
    // then stores foreground rows only, ~1 % of the anchors
    float4* g4 = reinterpret_cast<float4*>(p.grad_distri + (i >> 2) * (4 * kRegBins));
    for (int q = side; q < kRegBins; q += 4) g4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---- 3. top-13 anchors per box ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tal_topk_kernel(const LossParams p) {
  extern __shared__ double s_metric[];  // [A] candidate metrics (67 KB at 8400 anchors: three CTAs per SM, one wave for 16 x 20 boxes)
  pdl_launch_dependents();
  pdl_wait();
  const int g = blockIdx.x, b = blockIdx.y;
  if (!p.mask_gt[b * p.G + g]) return;  // padding rows vote for anchor 0 thirteen times and are dropped (tal_assigner.py:121-127)
  const double* gt = p.gt + (static_cast<size_t>(b) * p.G + g) * 5;
  const double gb[4] = {gt[1], gt[2], gt[3], gt[4]};
  int label = static_cast<int>(static_cast<long long>(gt[0]));
  label = label < 0 ? 0 : (label >= p.nc ? p.nc - 1 : label);  // memory safety only: real boxes carry valid classes
  const float* score = p.pred_scores + static_cast<size_t>(b) * p.A * p.nc + label;
  // Anchors whose centre can lie inside the box form one rectangle of cells per level: enumerate those (one cell of
  // margin; the exact float64 test of assigner_utils.py:25-45 decides) instead of all A anchors — ~300 instead of 8400
  // for a 100-pixel box, for the metric pass AND for each of the 13 selection rounds.
  int n_l[3] = {p.n0, p.n1, p.n2}, x_lo[3], y_lo[3], nx[3], cnt[4];
  cnt[0] = 0;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const double st = static_cast<double>(8 << l);
    int xl = static_cast<int>(floor(gb[0] / st - 0.5)) - 1, xh = static_cast<int>(ceil(gb[2] / st - 0.5)) + 1;
    int yl = static_cast<int>(floor(gb[1] / st - 0.5)) - 1, yh = static_cast<int>(ceil(gb[3] / st - 0.5)) + 1;
    xl = xl < 0 ? 0 : xl;
    yl = yl < 0 ? 0 : yl;
    xh = xh > n_l[l] - 1 ? n_l[l] - 1 : xh;
    yh = yh > n_l[l] - 1 ? n_l[l] - 1 : yh;
    x_lo[l] = xl;
    y_lo[l] = yl;
    nx[l] = xh >= xl ? xh - xl + 1 : 0;
    const int ny = yh >= yl ? yh - yl + 1 : 0;
    cnt[l + 1] = cnt[l] + nx[l] * ny;
  }
  const int n_cand = cnt[3];
  auto cand_anchor = [&](int i) {  // candidate number -> anchor index (level-major, row-major)
    const int l = i < cnt[1] ? 0 : (i < cnt[2] ? 1 : 2);
    const int local = i - cnt[l];
    const int cy = y_lo[l] + local / nx[l], cx = x_lo[l] + local % nx[l];
    return (l > 0 ? p.n0 * p.n0 : 0) + (l > 1 ? p.n1 * p.n1 : 0) + cy * n_l[l] + cx;
  };
  // four candidates per thread and pass: all global loads (box, score) of the four are issued before the first float64
  // division consumes one (the pass is bound by cold L2 / DRAM round trips, not by arithmetic)
  for (int i0 = threadIdx.x; i0 < n_cand; i0 += 4 * blockDim.x) {
    bool inside[4];
    float4 box4[4];
    float sc4[4], stride4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      inside[u] = false;
      if (i < n_cand) {
        const int a = cand_anchor(i);
        const Anchor an = anchor_of(p, a);
        const double ax = an.px, ay = an.py;
        // centre strictly inside the box (assigner_utils.py:25-45)
        const double dmin = fmin(fmin(__dsub_rn(ax, gb[0]), __dsub_rn(ay, gb[1])), fmin(__dsub_rn(gb[2], ax), __dsub_rn(gb[3], ay)));
        stride4[u] = an.stride;
        inside[u] = dmin > kTalEps;
        if (inside[u]) {
          box4[u] = *reinterpret_cast<const float4*>(p.boxes + (static_cast<size_t>(b) * p.A + a) * 4);
          sc4[u] = score[static_cast<size_t>(a) * p.nc];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < n_cand) {
        double m = 0.0;
        if (inside[u]) {  // pred_bboxes.detach() * stride_tensor in fp32, then the float64 IoU
          const double iou = iou_gt_pred(gb, __fmul_rn(box4[u].x, stride4[u]), __fmul_rn(box4[u].y, stride4[u]),
                                         __fmul_rn(box4[u].z, stride4[u]), __fmul_rn(box4[u].w, stride4[u]));
          m = __dmul_rn(static_cast<double>(sc4[u]), pow6(iou));  // tal_assigner.py:107
        }
        s_metric[i] = m;
      }
    }
  }
  __syncthreads();
  // 13 rounds of block-wide arg-max over the candidates (ties go to the earlier candidate).  A variant in which every warp
  // first extracts the top 13 of its own slice with shuffles only and warp 0 merges the 8 short lists measured slower
  // (37 vs 33 us for 16 x 20 boxes): the rounds are short, the kernel's time is the largest box's CTA.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_win;
  uint8_t* pos = p.pos + (static_cast<size_t>(b) * p.G + g) * p.A;
  for (int round = 0; round < kTopK; ++round) {
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n_cand; i += blockDim.x) {
      const double v = s_metric[i];
      if (v > best) {  // strictly greater: the earliest candidate wins among equals
        best = v;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_val[warp] = best;
      s_idx[warp] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double v = s_val[0];
      int wi = s_idx[0];
      for (int w = 1; w < 8; ++w)
        if (s_val[w] > v || (s_val[w] == v && s_idx[w] < wi)) {
          v = s_val[w];
          wi = s_idx[w];
        }
      // a metric of exactly 0 is an anchor outside the box (or a zero score): mask_in_gts drops it whichever of the
      // tied zeros torch.topk would have returned
      if (v > 0.0) {
        pos[cand_anchor(wi)] = 1;
        s_metric[wi] = -1.0;
        s_win = 1;
      } else {
        s_win = 0;
      }
    }
    __syncthreads();
    if (!s_win) break;
  }
}

// ---- 3b. ATSS warm-up assigner (atss_assigner.py:44-71): candidates of one box -------------------------------------------------
// One warp per (image, box).  Per level the 9 anchor centres closest to the box centre (they always lie in the 9 x 9 window of
// cells around the centre's cell, shifted into the grid at the borders); threshold = mean + unbiased std of the 27 candidates'
// anchor-box IoUs; positives = candidates above it whose centre lies strictly inside the box.
__global__ void __launch_bounds__(256) atss_select_kernel(const LossParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pair >= p.B * p.G) return;
  if (!p.mask_gt[pair]) return;
  const int b = pair / p.G;
  const double* gt = p.gt + static_cast<size_t>(pair) * 5;
  const double gb[4] = {gt[1], gt[2], gt[3], gt[4]};
  const double gcx = __ddiv_rn(__dadd_rn(gb[0], gb[2]), 2.0), gcy = __ddiv_rn(__dadd_rn(gb[1], gb[3]), 2.0);  // assigner_utils.py:14-16
  uint8_t* pos = p.pos + static_cast<size_t>(pair) * p.A;
  const int n_l[3] = {p.n0, p.n1, p.n2};
  // this lane's selected candidate of each level (lanes 0..8 hold the 9 picks of a level): anchor index and IoU
  int sel_a[3] = {-1, -1, -1};
  double sel_iou[3] = {0.0, 0.0, 0.0};
  double sum = 0.0;
  int n_sel = 0;
  int level_off = 0;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const int n = n_l[l];
    const float st = static_cast<float>(8 << l);
    const int win = n < 9 ? n : 9;
    int cx0 = static_cast<int>(floor(gcx / static_cast<double>(st))), cy0 = static_cast<int>(floor(gcy / static_cast<double>(st)));
    cx0 = cx0 < 0 ? 0 : (cx0 > n - 1 ? n - 1 : cx0);
    cy0 = cy0 < 0 ? 0 : (cy0 > n - 1 ? n - 1 : cy0);
    int xs = cx0 - 4, ys = cy0 - 4;
    xs = xs < 0 ? 0 : (xs > n - win ? n - win : xs);
    ys = ys < 0 ? 0 : (ys > n - win ? n - win : ys);
    // distances of the window's cells (3 per lane), assigner_utils.py:17-21: float64 centre - fp32 anchor centre
    double d[3];
    int ai[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int c = lane + 32 * u;
      d[u] = INFINITY;
      ai[u] = 0x7fffffff;
      if (c < win * win) {
        const int cy = ys + c / win, cx = xs + c % win;
        const float acx = __fmul_rn(static_cast<float>(cx) + 0.5f, st), acy = __fmul_rn(static_cast<float>(cy) + 0.5f, st);
        const double dx = __dsub_rn(gcx, static_cast<double>(acx)), dy = __dsub_rn(gcy, static_cast<double>(acy));
        d[u] = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        ai[u] = level_off + cy * n + cx;
      }
    }
    const int k = win * win < 9 ? win * win : 9;
    for (int r = 0; r < k; ++r) {
      double best = d[0];
      int bi = ai[0];
#pragma unroll
      for (int u = 1; u < 3; ++u)
        if (d[u] < best || (d[u] == best && ai[u] < bi)) {
          best = d[u];
          bi = ai[u];
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < best || (ov == best && oi < bi)) {
          best = ov;
          bi = oi;
        }
      }
#pragma unroll
      for (int u = 0; u < 3; ++u)
        if (ai[u] == bi) d[u] = INFINITY;  // taken
      if (lane == r) sel_a[l] = bi;
    }
    if (sel_a[l] >= 0) {
      const int local = sel_a[l] - level_off;
      const int cy = local / n, cx = local - cy * n;
      sel_iou[l] = iou_gt_anchor(gb, __fmul_rn(static_cast<float>(cx) + 0.5f, st), __fmul_rn(static_cast<float>(cy) + 0.5f, st), st);
      sum += sel_iou[l];
      ++n_sel;
    }
    level_off += n * n;
  }
  // mean + unbiased std over all candidates of the box (atss_assigner.py:136-139)
  double tot = sum;
  int cnt = n_sel;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  const double mean = tot / static_cast<double>(cnt);
  double ss = 0.0;
#pragma unroll
  for (int l = 0; l < 3; ++l)
    if (sel_a[l] >= 0) ss += (sel_iou[l] - mean) * (sel_iou[l] - mean);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const double thr = mean + sqrt(ss / static_cast<double>(cnt - 1));
  level_off = 0;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    if (sel_a[l] >= 0 && sel_iou[l] > thr) {
      const Anchor an = anchor_of(p, sel_a[l]);
      const double ax = an.px, ay = an.py;
      const double dmin = fmin(fmin(__dsub_rn(ax, gb[0]), __dsub_rn(ay, gb[1])), fmin(__dsub_rn(gb[2], ax), __dsub_rn(gb[3], ay)));
      if (dmin > kTalEps) pos[sel_a[l]] = 1;
    }
  }
  (void)b;
}

// ---- 4. one box per anchor ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tal_resolve_kernel(const LossParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= p.A) return;
  const size_t ba = static_cast<size_t>(b) * p.A + a;
  int count = 0, first = 0;
  for (int g = 0; g < p.G; ++g) {
    const int v = p.pos[(static_cast<size_t>(b) * p.G + g) * p.A + a];
    if (v && count == 0) first = g;
    count += v;
  }
  if (count == 0) {
    p.gt_idx[ba] = 0;
    p.fg[ba] = 0;
    p.align_a[ba] = 0.0;
    p.ovl_a[ba] = 0.0;
    return;
  }
  const Anchor an = anchor_of(p, a);
  float pb[4];
  pred_box_px(p, b, a, an.stride, pb);
  int gs = first;
  double iou;
  if (count > 1) {  // assigner_utils.py:60-66: arg-max of the IoU over ALL boxes of the image (first maximum)
    double best = -1.0;
    for (int g = 0; g < p.G; ++g) {
      const double* gt = p.gt + (static_cast<size_t>(b) * p.G + g) * 5;
      // the overlaps the assigner works with: prediction boxes (task-aligned) or anchor boxes (ATSS)
      const double v = p.atss ? iou_gt_anchor(gt + 1, an.px, an.py, an.stride) : iou_gt_pred(gt + 1, pb[0], pb[1], pb[2], pb[3]);
      if (v > best) {
        best = v;
        gs = g;
      }
    }
    iou = p.atss ? iou_gt_pred(p.gt + (static_cast<size_t>(b) * p.G + gs) * 5 + 1, pb[0], pb[1], pb[2], pb[3]) : best;
  } else {
    iou = iou_gt_pred(p.gt + (static_cast<size_t>(b) * p.G + gs) * 5 + 1, pb[0], pb[1], pb[2], pb[3]);
  }
  const double* gt = p.gt + (static_cast<size_t>(b) * p.G + gs) * 5;
  long long lab = static_cast<long long>(gt[0]);
  if (lab < 0) lab += p.nc;  // pd_scores[..., -1] of a padding row (tal_assigner.py:101-104); cannot win with IoU 0
  if (lab < 0 || lab >= p.nc) lab = 0;
  const double sc = p.pred_scores[ba * p.nc + lab];
  const double align = __dmul_rn(sc, pow6(iou));
  p.gt_idx[ba] = gs;
  p.fg[ba] = 1;
  p.align_a[ba] = align;
  p.ovl_a[ba] = iou;
  atomicMax(p.pos_align + b * p.G + gs, static_cast<unsigned long long>(__double_as_longlong(align)));
  atomicMax(p.pos_ovl + b * p.G + gs, static_cast<unsigned long long>(__double_as_longlong(iou)));
  atomicAdd(p.counters, 1);
}

// fixed-order block sum of one double per thread (256 threads) -> thread 0
__device__ __forceinline__ double block_sum_256(double v, double* s_red /* [8] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; ++w) t += s_red[w];
  __syncthreads();
  return t;
}

// ---- 5. normalised target scores (tal_assigner.py:66-72), class of every anchor, target_scores_sum -------------------------
// sums partial[0..n) in a fixed order
__device__ __forceinline__ double reduce_partials(const double* part, int n, double* s_red) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += part[i];
  return block_sum_256(acc, s_red);
}

// true in exactly one block: the one that finishes last (its view of the other blocks' partial sums is complete)
__device__ __forceinline__ bool last_block_done(int32_t* counter, int n_blocks) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1) == n_blocks - 1;
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

__global__ void __launch_bounds__(256) tal_norm_kernel(const LossParams p) {
  __shared__ double s_red[8];
  pdl_launch_dependents();
  pdl_wait();
  const size_t n = static_cast<size_t>(p.B) * p.A;
  double acc = 0.0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    double v = 0.0;
    int tl = -1;
    if (p.fg[i]) {
      const int b = static_cast<int>(i / p.A);
      const int gs = p.gt_idx[i];
      const double pa = __longlong_as_double(static_cast<long long>(p.pos_align[b * p.G + gs]));
      const double po = __longlong_as_double(static_cast<long long>(p.pos_ovl[b * p.G + gs]));
      v = __ddiv_rn(__dmul_rn(p.align_a[i], po), __dadd_rn(pa, kTalEps));
      // ATSS soft label (atss_assigner.py:82-85): IoU of the assigned box with the predicted box, multiplied IN PLACE into
      // the fp32 one-hot, i.e. rounded to fp32
      if (p.atss) v = static_cast<double>(static_cast<float>(p.ovl_a[i]));
      const long long lab = static_cast<long long>(p.gt[(static_cast<size_t>(b) * p.G + gs) * 5]);
      tl = lab < 0 ? 0 : static_cast<int>(lab);  // tal_assigner.py:143
    }
    p.norm[i] = v;
    p.tlabel[i] = tl;
    acc += v;
  }
  const double t = block_sum_256(acc, s_red);
  if (threadIdx.x == 0) p.partial[blockIdx.x] = t;
  if (last_block_done(p.counters + 2, gridDim.x)) {  // fixed-order sum whichever block comes last
    const double tss = reduce_partials(p.partial, gridDim.x, s_red);
    if (threadIdx.x == 0) p.scalars[4] = tss;
  }
}

// ---- 6. varifocal loss (loss.py:181-192) + gradient ---------------------------------------------------------------------------
struct VflOut {
  double term;
  float grad;
};
// one class probability `pr`; `t` = normalised target score if this is the anchor's assigned class, else nullptr semantics
__device__ __forceinline__ VflOut vfl_one(float pr, bool is_t, double t, double inv_tss) {
  const float l1 = fmaxf(log1pf(-pr), -100.0f);                // BCE's clamp
  const float den = fmaxf((1.0f - pr) * pr, 1e-12f);           // BCE backward's clamp
  VflOut o;
  double grad = 0.0;
  if (is_t) {
    const float l2 = fmaxf(logf(pr), -100.0f);
    const float t32 = static_cast<float>(t);
    const float bce = (t32 - 1.0f) * l1 - t32 * l2;
    o.term = static_cast<double>(bce) * t;                     // weight = gt_score on the assigned class
    grad = t * static_cast<double>((pr - t32) / den);
  } else {
    const float w = 0.75f * (pr * pr);                         // alpha * p^gamma on every other class
    const float bce = -l1;                                     // (0 - 1) * l1 - 0 * l2
    o.term = static_cast<double>(bce) * static_cast<double>(w);
    o.grad = (w * (pr / den) + bce * (1.5f * pr)) * static_cast<float>(inv_tss);  // fp32 like the reference's backward
    return o;
  }
  o.grad = static_cast<float>(grad * inv_tss);                 // loss weight 'class' = 1.0
  return o;
}

template <bool kVec4>
__global__ void __launch_bounds__(256) loss_vfl_kernel(const LossParams p) {
  __shared__ double s_red[8];
  pdl_launch_dependents();
  pdl_wait();
  const double inv_tss = 1.0 / p.scalars[4];
  double acc = 0.0;
  if (kVec4) {
    const unsigned per_row = p.nc >> 2;
    const size_t n4 = static_cast<size_t>(p.B) * p.A * per_row;
    const float4* in4 = reinterpret_cast<const float4*>(p.pred_scores);
    float4* out4 = reinterpret_cast<float4*>(p.grad_scores);
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const size_t row = i / per_row;
      const int c0 = static_cast<int>(i - row * per_row) << 2;
      const float4 v = in4[i];
      const int tl = p.tlabel[row];
      const double t = tl >= 0 ? p.norm[row] : 0.0;
      const VflOut a0 = vfl_one(v.x, tl == c0, t, inv_tss), a1 = vfl_one(v.y, tl == c0 + 1, t, inv_tss);
      const VflOut a2 = vfl_one(v.z, tl == c0 + 2, t, inv_tss), a3 = vfl_one(v.w, tl == c0 + 3, t, inv_tss);
      acc += (a0.term + a1.term) + (a2.term + a3.term);
      if (out4) out4[i] = make_float4(a0.grad, a1.grad, a2.grad, a3.grad);
    }
  } else {
    const size_t n = static_cast<size_t>(p.B) * p.A * p.nc;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const size_t row = i / p.nc;
      const int c = static_cast<int>(i - row * p.nc);
      const int tl = p.tlabel[row];
      const VflOut a = vfl_one(p.pred_scores[i], tl == c, tl >= 0 ? p.norm[row] : 0.0, inv_tss);
      acc += a.term;
      if (p.grad_scores) p.grad_scores[i] = a.grad;
    }
  }
  const double t = block_sum_256(acc, s_red);
  if (threadIdx.x == 0) p.partial[kMaxPartials + blockIdx.x] = t;
}

// ---- 7. GIoU + DFL of the foreground anchors (loss.py:195-254) + gradient -----------------------------------------------------
// forward-mode dual number: value + derivative w.r.t. ONE chosen coordinate of the predicted box
struct Dual {
  double v, d;
};
__device__ __forceinline__ Dual dk(double v) { return {v, 0.0}; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) { return {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; }
__device__ __forceinline__ Dual dmin(Dual a, Dual b) { return a.v <= b.v ? a : b; }
__device__ __forceinline__ Dual dmax(Dual a, Dual b) { return a.v >= b.v ? a : b; }
__device__ __forceinline__ Dual dclamp0(Dual a) { return a.v >= 0.0 ? a : Dual{0.0, 0.0}; }

__global__ void __launch_bounds__(256) loss_box_kernel(const LossParams p) {
  __shared__ double s_red[8];
  pdl_launch_dependents();
  pdl_wait();
  const double tss = p.scalars[4];
  const size_t n = static_cast<size_t>(p.B) * p.A * 4;
  double iou_term = 0.0, dfl_term = 0.0;
  // (anchor, side) items in blocks of 256, dealt round-robin to the CTAs; every thread of a CTA runs the same number of
  // rounds (the quad shuffles need whole warps); a warp without a foreground lane skips the round after one byte load
  const size_t chunks = (n + 255) / 256;
  for (size_t chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
  const size_t i = chunk * 256 + threadIdx.x;  // (b, a, side); a quad = one anchor
  const bool in_range = i < n;
  const size_t row = in_range ? (i >> 2) : 0;
  const int side = static_cast<int>(i & 3);
  const bool fg = in_range && p.fg[row];
  if (!__any_sync(0xffffffffu, fg)) continue;
  // quads never straddle a warp and n % 4 == 0, so the shuffles below see whole anchors; lanes of background anchors run
  // the arithmetic on zeros and store nothing but zero gradients
  float prob[kRegBins];
  float lse = 0.f, e = 0.f;
  const float* logit = p.pred_distri + (in_range ? i : 0) * kRegBins;
  if (fg) e = dfl_expect(logit, prob, &lse);
  const int b = static_cast<int>(row / p.A), a = static_cast<int>(row - static_cast<size_t>(b) * p.A);
  const Anchor an = anchor_of(p, a);
  const float cs = (side & 1) ? an.py / an.stride : an.px / an.stride;
  double my_coord = 0.0, tgt_coord = 0.0, bw = 0.0;
  if (fg) {
    my_coord = p.boxes[i];  // the box the assigner saw (decode_kernel, or the caller's)
    const double* gt = p.gt + (static_cast<size_t>(b) * p.G + p.gt_idx[row]) * 5;
    tgt_coord = __ddiv_rn(gt[1 + side], static_cast<double>(an.stride));  // target_bboxes /= stride_tensor (loss.py:141)
    bw = p.norm[row];                                                     // target_scores.sum(-1)
  }
  const unsigned quad_base = (threadIdx.x & 31) & ~3u;
  double pbx[4], tbx[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    pbx[s] = __shfl_sync(0xffffffffu, my_coord, quad_base + s);
    tbx[s] = __shfl_sync(0xffffffffu, tgt_coord, quad_base + s);
  }
  double g_coord = 0.0;  // d giou_loss / d (this lane's coordinate)
  double dfl_ce = 0.0;   // this side's left/right cross-entropy
  double wl = 0.0, wr = 0.0;
  int tl = 0;
  if (fg) {
    // figure_iou.py:37-65, box1 = prediction, box2 = target; the derivative is taken w.r.t. coordinate `side` of box1
    Dual x1 = dk(pbx[0]), y1 = dk(pbx[1]), x2 = dk(pbx[2]), y2 = dk(pbx[3]);
    (side == 0 ? x1 : side == 1 ? y1 : side == 2 ? x2 : y2).d = 1.0;
    const Dual u1 = dk(tbx[0]), v1 = dk(tbx[1]), u2 = dk(tbx[2]), v2 = dk(tbx[3]);
    const Dual eps = dk(kIouLossEps);
    const Dual inter = dclamp0(dmin(x2, u2) - dmax(x1, u1)) * dclamp0(dmin(y2, v2) - dmax(y1, v1));
    // w1, h1 (+ eps) and w1 * h1 are expressions of the fp32 prediction only: fp32 in the reference (the derivatives are
    // taken of the exact expression)
    const float fx1 = static_cast<float>(pbx[0]), fy1 = static_cast<float>(pbx[1]), fx2 = static_cast<float>(pbx[2]), fy2 = static_cast<float>(pbx[3]);
    const float w1f = __fsub_rn(fx2, fx1), h1f = __fadd_rn(__fsub_rn(fy2, fy1), 1e-10f);
    const Dual w1 = {static_cast<double>(w1f), (x2 - x1).d}, h1 = {static_cast<double>(h1f), (y2 - y1).d};
    Dual a1 = w1 * h1;
    a1.v = static_cast<double>(__fmul_rn(w1f, h1f));
    const Dual w2 = u2 - u1, h2 = v2 - v1 + eps;
    const Dual uni = a1 + w2 * h2 - inter + eps;
    const Dual iou = inter / uni;
    const Dual cw = dmax(x2, u2) - dmin(x1, u1), ch = dmax(y2, v2) - dmin(y1, v1);
    const Dual hull = cw * ch + eps;
    const Dual loss = dk(1.0) - (iou - (hull - uni) / hull);
    g_coord = loss.d;
    if (side == 0) iou_term += loss.v * bw;  // one lane per anchor contributes the value
    // DFL (loss.py:243-254): target distance of this side, clipped to [0, reg_max - 0.01] (general.py:43-49)
    const double csd = static_cast<double>(cs);
    double tgt = side < 2 ? csd - tbx[side] : tbx[side] - csd;
    tgt = fmin(fmax(tgt, 0.0), 16.0 - 0.01);
    tl = static_cast<int>(tgt);
    wl = static_cast<double>(static_cast<float>(tl + 1)) - tgt;
    wr = 1.0 - wl;
    const float ce_l = lse - logit[tl], ce_r = lse - logit[tl + 1];  // F.cross_entropy, fp32
    dfl_ce = static_cast<double>(ce_l) * wl + static_cast<double>(ce_r) * wr;
  }
  // mean over the four sides
  double ce4 = dfl_ce;
  ce4 += __shfl_xor_sync(0xffffffffu, ce4, 1);
  ce4 += __shfl_xor_sync(0xffffffffu, ce4, 2);
  if (fg && side == 0) dfl_term += (ce4 / 4.0) * bw;
  if (p.grad_distri && in_range) {
    float* go = p.grad_distri + i * kRegBins;
    if (fg) {
      // coordinate = point -/+ expectation; d expectation / d logit_j = p_j (j - E)
      const double g_dist = (side < 2 ? -g_coord : g_coord) * bw * 2.5 / tss;
      const double g_dfl = 0.25 * bw * 0.5 / tss;
#pragma unroll
      for (int j = 0; j < kRegBins; ++j) {
        const double pj = prob[j];
        double gj = g_dist * pj * (static_cast<double>(j) - static_cast<double>(e));
        gj += g_dfl * (pj - (j == tl ? wl : 0.0) - (j == tl + 1 ? wr : 0.0));
        go[j] = static_cast<float>(gj);
      }
    }  // background rows were zeroed by decode_kernel (or by the host's memset when the caller supplied the boxes)
  }
  }  // chunk
  const double ti = block_sum_256(iou_term, s_red);
  const double td = block_sum_256(dfl_term, s_red);
  if (threadIdx.x == 0) {
    p.partial[2 * kMaxPartials + blockIdx.x] = ti;
    p.partial[3 * kMaxPartials + blockIdx.x] = td;
  }
  // ---- totals (loss.py:145-162) by whichever block finishes last, always in the same order ---------------------------------
  if (last_block_done(p.counters + 3, gridDim.x)) {
    const double cls = reduce_partials(p.partial + kMaxPartials, p.n_vfl, s_red);
    const double iou = reduce_partials(p.partial + 2 * kMaxPartials, gridDim.x, s_red);
    const double dfl = reduce_partials(p.partial + 3 * kMaxPartials, gridDim.x, s_red);
    if (threadIdx.x == 0) {
      const int num_fg = p.counters[0];
      const double l_cls = cls / tss;                      // inf / nan without a single foreground anchor, as the reference
      const double l_iou = num_fg > 0 ? iou / tss : 0.0;   // loss.py:205,238-240
      const double l_dfl = num_fg > 0 ? dfl / tss : 0.0;
      p.scalars[0] = 1.0 * l_cls + 2.5 * l_iou + 0.5 * l_dfl;
      p.scalars[1] = 2.5 * l_iou;
      p.scalars[2] = 0.5 * l_dfl;
      p.scalars[3] = 1.0 * l_cls;
      p.scalars[5] = static_cast<double>(num_fg);
      p.scalars[6] = static_cast<double>(p.counters[1]);
      p.scalars[7] = 0.0;
    }
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct LossLayout {
  size_t gt, mask_gt, pos, boxes, gt_idx, fg, align_a, ovl_a, norm, tlabel, pos_align, pos_ovl, partial, scalars, counters, total;
};
static LossLayout loss_layout(size_t B, size_t A, size_t G) {
  LossLayout l;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o = align_up(o + bytes, 256);
    return at;
  };
  l.gt = take(B * G * 5 * 8);
  l.mask_gt = take(B * G);
  l.pos = take(B * G * A);
  l.boxes = take(B * A * 4 * 4);
  l.gt_idx = take(B * A * 4);
  l.fg = take(B * A);
  l.align_a = take(B * A * 8);
  l.ovl_a = take(B * A * 8);
  l.norm = take(B * A * 8);
  l.tlabel = take(B * A * 4);
  l.pos_align = take(B * G * 8);
  l.pos_ovl = take(B * G * 8);
  l.partial = take(static_cast<size_t>(4) * kMaxPartials * 8);
  l.scalars = take(8 * 8);
  l.counters = take(4 * 4);
  l.total = o;
  return l;
}

}  // namespace mafb200

using namespace mafb200;

extern "C" size_t mafb200_loss_workspace_bytes(int32_t batch, int32_t anchors, int32_t gt_cap) {
  if (batch <= 0 || anchors <= 0 || gt_cap < 0) return 0;
  return loss_layout(batch, anchors, gt_cap > 0 ? gt_cap : 1).total;
}

extern "C" int32_t mafb200_detect_loss(const float* pred_scores, const float* pred_distri, const float* targets,
                                       int32_t num_targets, int32_t batch, int32_t img_size, int32_t num_classes,
                                       int32_t gt_cap, const float* boxes_override, void* workspace, size_t workspace_bytes,
                                       double* scalars_out, float* grad_scores, float* grad_distri, int32_t* out_gt_idx,
                                       uint8_t* out_fg, double* out_target_score, int32_t assigner, void* stream) {
  if (!pred_scores || !pred_distri || !workspace || !scalars_out) return fail(MAF_E_ARG, "detect_loss: null pointer");
  if (assigner != MAF_ASSIGN_TAL && assigner != MAF_ASSIGN_ATSS) return fail(MAF_E_ARG, "detect_loss: assigner %d", assigner);
  if (num_targets < 0 || (num_targets > 0 && !targets)) return fail(MAF_E_ARG, "detect_loss: bad targets");
  if (batch <= 0 || batch > 65535 || num_classes <= 0 || img_size <= 0 || img_size % 32 != 0)
    return fail(MAF_E_ARG, "detect_loss: batch %d / classes %d / image size %d (must be a multiple of 32)", batch, num_classes, img_size);
  if (gt_cap < 0 || gt_cap > 65535) return fail(MAF_E_ARG, "detect_loss: gt_cap %d", gt_cap);
  const int n0 = img_size / 8, n1 = img_size / 16, n2 = img_size / 32;
  const int A = n0 * n0 + n1 * n1 + n2 * n2;
  const int G = gt_cap > 0 ? gt_cap : 1;  // gt_cap 0: one padding row per image (no foreground; the reference's early return)
  const size_t topk_smem = static_cast<size_t>(A) * 8;  // one float64 metric per candidate (<= A)
  if (topk_smem > 200 * 1024) return fail(MAF_E_ARG, "detect_loss: %d anchors need %zu B of shared memory", A, topk_smem);
  if ((reinterpret_cast<uintptr_t>(pred_scores) | reinterpret_cast<uintptr_t>(pred_distri) | reinterpret_cast<uintptr_t>(workspace)) & 15)
    return fail(MAF_E_ALIGN, "detect_loss: predictions / workspace must be 16-B aligned");
  const LossLayout l = loss_layout(batch, A, G);
  if (workspace_bytes < l.total) return fail(MAF_E_WORKSPACE, "detect_loss: workspace %zu B < %zu B", workspace_bytes, l.total);
  int32_t rc = require_sm100();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);

  LossParams p;
  memset(&p, 0, sizeof(p));
  p.pred_scores = pred_scores;
  p.pred_distri = pred_distri;
  p.targets = targets;
  p.boxes_in = boxes_override;
  p.T = num_targets;
  p.B = batch;
  p.A = A;
  p.nc = num_classes;
  p.G = G;
  p.img = img_size;
  p.n0 = n0;
  p.n1 = n1;
  p.n2 = n2;
  p.gt = reinterpret_cast<double*>(ws + l.gt);
  p.mask_gt = ws + l.mask_gt;
  p.pos = ws + l.pos;
  p.boxes = boxes_override ? const_cast<float*>(boxes_override) : reinterpret_cast<float*>(ws + l.boxes);
  p.gt_idx = out_gt_idx ? out_gt_idx : reinterpret_cast<int32_t*>(ws + l.gt_idx);
  p.fg = out_fg ? out_fg : ws + l.fg;
  p.align_a = reinterpret_cast<double*>(ws + l.align_a);
  p.ovl_a = reinterpret_cast<double*>(ws + l.ovl_a);
  p.norm = out_target_score ? out_target_score : reinterpret_cast<double*>(ws + l.norm);
  p.pos_align = reinterpret_cast<unsigned long long*>(ws + l.pos_align);
  p.pos_ovl = reinterpret_cast<unsigned long long*>(ws + l.pos_ovl);
  p.partial = reinterpret_cast<double*>(ws + l.partial);
  p.scalars = scalars_out;
  p.counters = reinterpret_cast<int32_t*>(ws + l.counters);
  p.tlabel = reinterpret_cast<int32_t*>(ws + l.tlabel);
  p.grad_scores = grad_scores;
  p.grad_distri = grad_distri;
  p.atss = assigner == MAF_ASSIGN_ATSS;
  if (boxes_override && (reinterpret_cast<uintptr_t>(boxes_override) & 15)) return fail(MAF_E_ALIGN, "detect_loss: boxes must be 16-B aligned");
  if ((reinterpret_cast<uintptr_t>(grad_scores) | reinterpret_cast<uintptr_t>(grad_distri)) & 15)
    return fail(MAF_E_ALIGN, "detect_loss: gradients must be 16-B aligned");

  if (cudaMemsetAsync(p.pos, 0, static_cast<size_t>(batch) * G * A, st) != cudaSuccess)
    return fail(MAF_E_CUDA, "detect_loss: cudaMemsetAsync failed");
  if (boxes_override && grad_distri &&  // decode_kernel (skipped then) is what zeroes the background rows
      cudaMemsetAsync(grad_distri, 0, static_cast<size_t>(batch) * A * 4 * kRegBins * sizeof(float), st) != cudaSuccess)
    return fail(MAF_E_CUDA, "detect_loss: cudaMemsetAsync failed");
  {
    static SmemOptIn opt_in;
    rc = smem_opt_in(opt_in, tal_topk_kernel, 200 * 1024, "tal_topk");
    if (rc) return rc;
  }
  const size_t rows = static_cast<size_t>(batch) * A;
  auto blocks_for = [](size_t n, int per_thread) {
    size_t b = (n + 256ull * per_thread - 1) / (256ull * per_thread);
    if (b < 1) b = 1;
    if (b > kMaxPartials) b = kMaxPartials;
    return static_cast<int>(b);
  };
  const int n_norm = blocks_for(rows, 4), n_vfl = blocks_for(rows * num_classes, 16);
  p.n_vfl = n_vfl;
  size_t box_blocks = (rows * 4 + 255) / 256;  // chunks of 256 (anchor, side) items, dealt round-robin
  if (box_blocks > 1184) box_blocks = 1184;     // 8 CTAs on each of 148 SMs
  launch_pdl(loss_targets_kernel, dim3(1), dim3(256), 0, st, p);
  if (!boxes_override) launch_pdl(loss_decode_kernel, dim3(static_cast<unsigned>((rows * 4 + 255) / 256)), dim3(256), 0, st, p);
  if (p.atss)
    launch_pdl(atss_select_kernel, dim3((batch * G + 7) / 8), dim3(256), 0, st, p);
  else
    launch_pdl(tal_topk_kernel, dim3(G, batch), dim3(256), topk_smem, st, p);
  launch_pdl(tal_resolve_kernel, dim3((A + 255) / 256, batch), dim3(256), 0, st, p);
  launch_pdl(tal_norm_kernel, dim3(n_norm), dim3(256), 0, st, p);
  if (num_classes % 4 == 0)
    launch_pdl(loss_vfl_kernel<true>, dim3(n_vfl), dim3(256), 0, st, p);
  else
    launch_pdl(loss_vfl_kernel<false>, dim3(n_vfl), dim3(256), 0, st, p);
  launch_pdl(loss_box_kernel, dim3(static_cast<unsigned>(box_blocks)), dim3(256), 0, st, p);
  return check_launch("detect_loss kernels");
}
