// K8 — batched NMS with the exact semantics of the reference's non_max_suppression
// (yolov6/utils/nms.py:31-105) + torchvision.ops.nms, without host round trips:
//
//   nms.py:48   candidates: obj > conf  AND  max_c cls > conf          (fp32 compares)
//   nms.py:69   score = cls * obj
//   nms.py:72   box = xywh2xyxy (x1 = cx - w/2 ...)                    (fp32)
//   nms.py:75-80 multi_label: every (anchor, class) with score > conf, in (anchor asc, class asc)
//               order; else the best class per anchor (first maximum), kept if score > conf
//   nms.py:83-84 optional class filter
//   nms.py:90-91 more than max_nms candidates -> keep the max_nms best scores
//   nms.py:94-95 boxes + class * 4096 (0 if agnostic)                  (fp32 add, rounding kept)
//   nms.py:96   torchvision nms: stable sort by score descending, greedy suppression where
//               inter / (area_i + area_j - inter) > iou_thres (strict, IEEE division,
//               threshold compared in double as torchvision's CPU kernel does)
//   nms.py:97-100 first max_det survivors, in score order
//   (nms.py:101-103, the 10 s wall-clock bail-out, is deliberately NOT reproduced.)
//
// Kernel 1 (compaction): one warp per anchor row; candidates are emitted UNORDERED (warp-
//   aggregated atomics) as 64-bit keys  (~score_bits << 32) | (anchor*nc + class): ascending key
//   order == (score desc, anchor asc, class asc) == the reference's stable sort of its ordered list.
// Kernel 2 (one CTA per image): bitonic sort of the keys (shared memory when they fit, global
//   workspace otherwise), then greedy NMS in score order over 256-candidate chunks: every
//   candidate is tested against the boxes kept so far (<= max_det of them, in smem), the chunk's
//   own 256x256 IoU bit matrix resolves intra-chunk dependencies, and the loop stops as soon as
//   max_det boxes are kept (later boxes cannot change the first max_det results).
// All float ops that decide a comparison use explicit round-to-nearest intrinsics so the
// compiler cannot contract them into FMAs.
#include <string.h>

#include "common.cuh"
#include "host.h"
#include "nms_filter.cuh"

namespace mafb200 {

constexpr int kSortSmemKeys = 8192;  // 64 KB of keys
constexpr int kChunk = 256;
constexpr int kMaxDetCap = 2048;
constexpr float kMaxWh = 4096.0f;    // nms.py:54

struct NmsParams {
  const float* pred;
  const float* box;     // (cx, cy, w, h) of anchor a of image b at box[(b * A + a) * box_stride]: pred itself
  int32_t box_stride;   // (stride 5 + nc) or the compact [B, A, 4] array of the fused decode (stride 4)
  int32_t B, A, nc;
  float conf;
  double iou;
  int32_t multi_label, agnostic;
  const uint8_t* class_filter;
  int32_t max_det, max_nms;
  float* det;
  int32_t* count;
  long long det_stride;    // floats between the detections of consecutive images (0 -> max_det * 6)
  long long count_stride;  // int32 elements between consecutive counts (0 -> 1)
  int32_t* ncand;       // [B]
  unsigned long long* keys;  // [B][cap_pow2]
  long long cap_pow2;
};

// ---- kernel 1: threshold + compaction ----------------------------------------------------------
// A CTA stages kCompactRows consecutive prediction rows (one contiguous span of pred) in shared memory
// with coalesced 16-byte loads — all of them in flight before the first use — then each warp filters
// rows from there.  (One warp per row straight from global memory ran at 1.7 TB/s: three dependent,
// unaligned 128-byte loads per row and nothing else in flight.)
constexpr int kCompactRows = 64;

__global__ void __launch_bounds__(256) nms_compact_kernel(const NmsParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float s_rows[];  // [kCompactRows][5 + nc]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int a0 = blockIdx.x * kCompactRows;
  const int rows = min(kCompactRows, p.A - a0);
  const int no = 5 + p.nc;
  const float* src = p.pred + (static_cast<size_t>(b) * p.A + a0) * no;
  const int nflt = rows * no;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int n4 = nflt >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    float4* dst4 = reinterpret_cast<float4*>(s_rows);
#pragma unroll 6
    for (int i = threadIdx.x; i < n4; i += 256) dst4[i] = __ldg(src4 + i);
    for (int i = (n4 << 2) + threadIdx.x; i < nflt; i += 256) s_rows[i] = __ldg(src + i);
  } else {
#pragma unroll 8
    for (int i = threadIdx.x; i < nflt; i += 256) s_rows[i] = __ldg(src + i);
  }
  __syncthreads();

  unsigned long long* keys = p.keys + static_cast<size_t>(b) * p.cap_pow2;
  for (int r = warp; r < rows; r += 8)
    nms_filter_row(s_rows + r * no, a0 + r, p.nc, p.conf, p.multi_label, p.class_filter, &p.ncand[b], keys, p.cap_pow2, lane);
}

// ---- bitonic sort helpers ------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort(unsigned long long* k, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int i = 2 * t - (t & (stride - 1));
        const int j = i + stride;
        const bool up = (i & size) == 0;
        const unsigned long long a = k[i], c = k[j];
        if ((a > c) == up) {
          k[i] = c;
          k[j] = a;
        }
      }
    }
  }
  __syncthreads();
}

// Intersection area exactly as torchvision computes it (fp32, no FMA contraction).
__device__ __forceinline__ float box_inter(float ax1, float ay1, float ax2, float ay2, float bx1, float by1, float bx2,
                                           float by2) {
  const float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1);
  const float xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1));
  const float h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  return __fmul_rn(w, h);
}
// inter / (area_a + area_b - inter) > thr with IEEE division, threshold compared in double (torchvision CPU).
// inter == 0 gives 0 (or NaN for two empty boxes): never > thr for thr in [0, 1], so callers test
// `inter > 0` first and only pay the division for boxes that actually overlap — with class-offset boxes
// (nms.py:94-95) that is a few percent of all pairs.
__device__ __forceinline__ bool iou_gt_inter(float inter, float aarea, float barea, double thr) {
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
  return static_cast<double>(ovr) > thr;
}

// ---- kernel 2: per-image sort + greedy NMS ---------------------------------------------------------
__global__ void __launch_bounds__(1024) nms_select_kernel(const NmsParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ unsigned long long s_dyn[];
  unsigned long long* s_keys = s_dyn;                                          // [kSortSmemKeys]
  float* k_box = reinterpret_cast<float*>(s_keys + kSortSmemKeys);             // kept: [max_det][5]
  float* c_box = k_box + static_cast<size_t>(round_up(p.max_det, 2)) * 5;      // chunk: [kChunk][5] (8-B aligned)
  unsigned long long* c_mask = reinterpret_cast<unsigned long long*>(c_box + kChunk * 5);  // [kChunk][4]
  int* c_alive = reinterpret_cast<int*>(c_mask + kChunk * 4);                  // [kChunk]
  int* c_kept = c_alive + kChunk;                                              // [kChunk] chunk-local indices kept
  unsigned* c_alive_bits = reinterpret_cast<unsigned*>(c_kept + kChunk);       // [kChunk / 32]
  __shared__ int s_nk;
  __shared__ int s_done;

  const int b = blockIdx.x;
  const int no = 5 + p.nc;
  long long n = p.ncand[b];
  if (n > p.cap_pow2) n = p.cap_pow2;
  float* det = p.det + static_cast<size_t>(b) * (p.det_stride ? p.det_stride : static_cast<long long>(p.max_det) * 6);
  int32_t* count_b = p.count + static_cast<size_t>(b) * (p.count_stride ? p.count_stride : 1);
  for (int i = threadIdx.x; i < p.max_det * 6; i += blockDim.x) det[i] = 0.0f;
  if (n == 0) {
    if (threadIdx.x == 0) *count_b = 0;
    return;
  }

  // ---- sort -----------------------------------------------------------------------------------
  // n <= 8192: bitonic sort in shared memory.  Larger: greedy NMS only ever needs the best-scoring prefix of
  // the list (it stops at max_det survivors), so first radix-SELECT the best <= 8192 keys (two 11-bit
  // histogram levels over the score bits; bins in key order, so the selected set is exactly a prefix of the
  // sorted list), sort and process those in shared memory, and fall back to the full global-memory sort only
  // if they run out before max_det boxes are kept.  Results are identical either way.
  unsigned long long* gkeys = p.keys + static_cast<size_t>(b) * p.cap_pow2;
  const int n_eff = static_cast<int>(n < p.max_nms ? n : p.max_nms);
  const float* pred_b = p.box + static_cast<size_t>(b) * p.A * p.box_stride;
  __shared__ int s_sel[4];  // [0] level-1 boundary bin, [1] level-2 boundary bin, [2] selected count, [3] fill cursor

  for (int attempt = (n > kSortSmemKeys ? 0 : 1); attempt < 2; ++attempt) {
  unsigned long long* keys;
  int n_run = n_eff;
  if (n <= kSortSmemKeys) {
    int P = 1;
    while (P < n) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_keys[i] = i < n ? gkeys[i] : ~0ull;
    keys = s_keys;
    bitonic_sort(keys, P);
  } else if (attempt == 0) {
    int* hist = reinterpret_cast<int*>(c_mask);  // 2048 bins (the chunk mask is not in use yet)
    const int budget = kSortSmemKeys;
    int prefix = 0;
    for (int level = 0; level < 2; ++level) {
      for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const int b1 = level == 1 ? s_sel[0] : 0;
      for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned hi = static_cast<unsigned>(gkeys[i] >> 32);
        if (level == 0) atomicAdd(&hist[hi >> 21], 1);
        else if (static_cast<int>(hi >> 21) == b1) atomicAdd(&hist[(hi >> 10) & 2047], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {  // first bin whose inclusion would exceed the budget (few bins are non-empty)
        int acc = prefix, bin = 0;
        for (; bin < 2048; ++bin) {
          if (acc + hist[bin] > budget) break;
          acc += hist[bin];
        }
        s_sel[level] = bin;
        s_sel[2] = acc;
      }
      __syncthreads();
      prefix = s_sel[2];
      if (s_sel[level] >= 2048) break;  // everything fits (cannot happen for n > budget at level 0)
    }
    const int m = s_sel[2];
    if (m < min(n_eff, 2 * p.max_det)) continue;  // too few selectable (massive score ties): full sort
    const int b1 = s_sel[0], b2 = s_sel[1];
    if (threadIdx.x == 0) s_sel[3] = 0;
    __syncthreads();
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long key = gkeys[i];
      const unsigned hi = static_cast<unsigned>(key >> 32);
      const int k1 = static_cast<int>(hi >> 21), k2 = static_cast<int>((hi >> 10) & 2047);
      if (k1 < b1 || (k1 == b1 && k2 < b2)) s_keys[atomicAdd(&s_sel[3], 1)] = key;
    }
    int P = 1;
    while (P < m) P <<= 1;
    __syncthreads();
    for (int i = m + threadIdx.x; i < P; i += blockDim.x) s_keys[i] = ~0ull;
    keys = s_keys;
    bitonic_sort(keys, P);
    n_run = m < n_eff ? m : n_eff;
  } else {
    int P = 1;
    while (P < n) P <<= 1;
    for (long long i = n + threadIdx.x; i < P; i += blockDim.x) gkeys[i] = ~0ull;
    keys = gkeys;
    bitonic_sort(keys, P);
  }

  __syncthreads();
  if (threadIdx.x == 0) {
    s_nk = 0;
    s_done = 0;
  }
  __syncthreads();

  for (int base = 0; base < n_run; base += kChunk) {
    const int cn = min(kChunk, n_run - base);
    const int nk0 = s_nk;
    // (1) materialise the chunk's class-offset boxes
    if (threadIdx.x < cn) {
      const unsigned idx = static_cast<unsigned>(keys[base + threadIdx.x] & 0xffffffffull);
      const int a = idx / p.nc, c = idx - a * p.nc;
      const float* row = pred_b + static_cast<size_t>(a) * p.box_stride;
      const float cx = row[0], cy = row[1], w = row[2], h = row[3];
      const float off = p.agnostic ? 0.0f : __fmul_rn(static_cast<float>(c), kMaxWh);
      const float x1 = __fadd_rn(__fsub_rn(cx, __fdiv_rn(w, 2.0f)), off);
      const float y1 = __fadd_rn(__fsub_rn(cy, __fdiv_rn(h, 2.0f)), off);
      const float x2 = __fadd_rn(__fadd_rn(cx, __fdiv_rn(w, 2.0f)), off);
      const float y2 = __fadd_rn(__fadd_rn(cy, __fdiv_rn(h, 2.0f)), off);
      float* cb = c_box + threadIdx.x * 5;
      cb[0] = x1;
      cb[1] = y1;
      cb[2] = x2;
      cb[3] = y2;
      cb[4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
      c_alive[threadIdx.x] = 1;
    }
    for (int i = threadIdx.x; i < kChunk * 4; i += blockDim.x) c_mask[i] = 0ull;
    __syncthreads();
    // (2) suppression by boxes kept in earlier chunks: 4 threads per candidate stride the kept list
    {
      const int ci = threadIdx.x >> 2, part = threadIdx.x & 3;
      if (ci < cn) {
        const float* cb = c_box + ci * 5;
        const float b0 = cb[0], b1 = cb[1], b2 = cb[2], b3 = cb[3], b4 = cb[4];
        bool dead = false;
        for (int k = part; k < nk0 && !dead; k += 4) {
          const float* kb = k_box + k * 5;
          const float inter = box_inter(kb[0], kb[1], kb[2], kb[3], b0, b1, b2, b3);
          if (inter > 0.0f) dead = iou_gt_inter(inter, kb[4], b4, p.iou);
        }
        if (dead) c_alive[ci] = 0;
      }
    }
    __syncthreads();
    // alive bitmap of the chunk (one ballot per warp of the first 256 threads)
    if (threadIdx.x < kChunk) {
      const unsigned bal = __ballot_sync(0xffffffffu, threadIdx.x < cn && c_alive[threadIdx.x] != 0);
      if ((threadIdx.x & 31) == 0) c_alive_bits[threadIdx.x >> 5] = bal;
    }
    // (3) intra-chunk IoU bit matrix: bit j of row i set if i suppresses j (j > i).  Rows of candidates
    //     already dead and 64-column words entirely at or below the diagonal are skipped.
    for (int t = threadIdx.x; t < kChunk * 4; t += blockDim.x) {
      const int i = t >> 2, wq = t & 3;
      const int j0 = wq * 64;
      if (i >= cn || !c_alive[i] || j0 + 63 <= i) continue;  // c_mask was zeroed in (1)
      const float* ib = c_box + i * 5;
      const float i0 = ib[0], i1 = ib[1], i2 = ib[2], i3 = ib[3], i4 = ib[4];
      const int jlo = max(j0, i + 1), jhi = min(j0 + 64, cn);
      // pass 1: which boxes overlap at all (no division); pass 2: exact IoU test for those only
      unsigned long long cand = 0ull;
      for (int j = jlo; j < jhi; ++j) {
        const float* jb = c_box + j * 5;
        if (box_inter(i0, i1, i2, i3, jb[0], jb[1], jb[2], jb[3]) > 0.0f) cand |= 1ull << (j - j0);
      }
      unsigned long long bits = 0ull;
      while (cand) {
        const int bpos = __ffsll(static_cast<long long>(cand)) - 1;
        cand &= cand - 1ull;
        const float* jb = c_box + (j0 + bpos) * 5;
        const float inter = box_inter(i0, i1, i2, i3, jb[0], jb[1], jb[2], jb[3]);
        if (iou_gt_inter(inter, i4, jb[4], p.iou)) bits |= 1ull << bpos;
      }
      c_mask[t] = bits;
    }
    __syncthreads();
    // (4) sequential resolve by one thread: walks only the SET bits of (alive & ~removed), so the cost is
    //     per kept box (<= max_det per image), shared memory and registers only
    if (threadIdx.x == 0) {
      unsigned long long rem[4] = {0ull, 0ull, 0ull, 0ull};
      int nk = nk0;
#pragma unroll
      for (int wi = 0; wi < 4; ++wi) {
        const unsigned long long alive =
            static_cast<unsigned long long>(c_alive_bits[2 * wi]) | (static_cast<unsigned long long>(c_alive_bits[2 * wi + 1]) << 32);
        unsigned long long todo = alive;
        while (nk < p.max_det) {
          const unsigned long long avail = todo & ~rem[wi];
          if (avail == 0ull) break;
          const int bpos = __ffsll(static_cast<long long>(avail)) - 1;
          const int i = wi * 64 + bpos;
          todo &= ~((2ull << bpos) - 1ull);  // everything up to and including bpos is decided
          rem[0] |= c_mask[i * 4 + 0];
          rem[1] |= c_mask[i * 4 + 1];
          rem[2] |= c_mask[i * 4 + 2];
          rem[3] |= c_mask[i * 4 + 3];
          c_kept[nk - nk0] = i;
          ++nk;
        }
      }
      s_nk = nk;
      s_done = nk >= p.max_det ? 1 : 0;
    }
    __syncthreads();
    // (5) parallel: append the kept boxes to the kept list and write their output rows
    {
      const int nk1 = s_nk;
      for (int t = threadIdx.x; t < nk1 - nk0; t += blockDim.x) {
        const int i = c_kept[t];
        const float* cb = c_box + i * 5;
        float* kb = k_box + (nk0 + t) * 5;
        kb[0] = cb[0];
        kb[1] = cb[1];
        kb[2] = cb[2];
        kb[3] = cb[3];
        kb[4] = cb[4];
        // output row: un-offset box, score, class (nms.py:76-80,100)
        const unsigned long long key = keys[base + i];
        const unsigned idx = static_cast<unsigned>(key & 0xffffffffull);
        const int a = idx / p.nc, c = idx - a * p.nc;
        const float* row = pred_b + static_cast<size_t>(a) * p.box_stride;
        const float cx = row[0], cy = row[1], w = row[2], h = row[3];
        float* o = det + static_cast<size_t>(nk0 + t) * 6;
        o[0] = __fsub_rn(cx, __fdiv_rn(w, 2.0f));
        o[1] = __fsub_rn(cy, __fdiv_rn(h, 2.0f));
        o[2] = __fadd_rn(cx, __fdiv_rn(w, 2.0f));
        o[3] = __fadd_rn(cy, __fdiv_rn(h, 2.0f));
        o[4] = __uint_as_float(~static_cast<unsigned>(key >> 32));
        o[5] = static_cast<float>(c);
      }
    }
    __syncthreads();
    if (s_done) break;
  }
  // a prefix run that kept max_det boxes, or that covered every candidate, is final
  if (n_run == n_eff || s_nk >= p.max_det) break;
  __syncthreads();
  }  // attempt
  if (threadIdx.x == 0) *count_b = s_nk;
}

static long long pow2_ceil(long long v) {
  long long p = 1;
  while (p < v) p <<= 1;
  return p;
}

static int32_t launch_select(const NmsParams& p, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(kSortSmemKeys) * 8 + static_cast<size_t>(round_up(p.max_det, 2)) * 5 * 4 +
                      kChunk * 5 * 4 + kChunk * 4 * 8 + kChunk * 4 * 2 + (kChunk / 32) * 4;
  {
    static SmemOptIn opt_in;  // per device (a process-wide flag skipped the opt-in on a second GPU)
    const int32_t rc_attr = smem_opt_in(opt_in, nms_select_kernel, 160 * 1024, "nms_select");
    if (rc_attr) return rc_attr;
  }
  launch_pdl<false>(nms_select_kernel, dim3(p.B), dim3(1024), smem, st, p);
  return check_launch("nms_select kernel launch");
}

}  // namespace mafb200

using namespace mafb200;

extern "C" size_t mafb200_nms_workspace_bytes(int32_t batch, int32_t anchors, int32_t nc) {
  if (batch <= 0 || anchors <= 0 || nc <= 0) return 0;
  const long long cap = pow2_ceil(static_cast<long long>(anchors) * nc);
  return 256 + static_cast<size_t>(batch) * cap * sizeof(unsigned long long) +
         ((static_cast<size_t>(batch) * 4 + 255) / 256) * 256;
}

extern "C" int32_t mafb200_nms(const float* pred, int32_t batch, int32_t anchors, int32_t nc, double conf_thres,
                               double iou_thres, int32_t multi_label, int32_t agnostic, const uint8_t* class_filter,
                               int32_t max_det, int32_t max_nms, float* det, int32_t* count, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (!pred || !det || !count || !workspace) return fail(MAF_E_ARG, "nms: null pointer");
  if (batch <= 0 || anchors <= 0 || nc <= 0) return fail(MAF_E_ARG, "nms: bad shape B=%d A=%d nc=%d", batch, anchors, nc);
  // same parameter checks as the reference's asserts (nms.py:50-51)
  if (!(conf_thres >= 0.0 && conf_thres <= 1.0)) return fail(MAF_E_ARG, "nms: conf_thres must be in [0,1], got %g", conf_thres);
  if (!(iou_thres >= 0.0 && iou_thres <= 1.0)) return fail(MAF_E_ARG, "nms: iou_thres must be in [0,1], got %g", iou_thres);
  if (max_det <= 0 || max_det > kMaxDetCap) return fail(MAF_E_ARG, "nms: max_det=%d (1..%d)", max_det, kMaxDetCap);
  if (max_nms <= 0) return fail(MAF_E_ARG, "nms: max_nms=%d", max_nms);
  if (static_cast<long long>(anchors) * nc > 0x7fffffffll) return fail(MAF_E_ARG, "nms: anchors*nc overflows int32");
  if (batch > 65535) return fail(MAF_E_ARG, "nms: batch %d > 65535", batch);
  if (workspace_bytes < mafb200_nms_workspace_bytes(batch, anchors, nc))
    return fail(MAF_E_WORKSPACE, "nms: workspace %zu < required %zu", workspace_bytes,
                mafb200_nms_workspace_bytes(batch, anchors, nc));
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(MAF_E_ALIGN, "nms: workspace must be 256-B aligned");
  int32_t rc = require_sm100();
  if (rc) return rc;

  NmsParams p;
  memset(&p, 0, sizeof(p));
  p.pred = pred;
  p.box = pred;
  p.box_stride = 5 + nc;
  p.B = batch;
  p.A = anchors;
  p.nc = nc;
  p.conf = static_cast<float>(conf_thres);  // torch compares an fp32 tensor with the scalar cast to fp32
  p.iou = iou_thres;
  p.multi_label = (multi_label != 0 && nc > 1) ? 1 : 0;  // nms.py:57
  p.agnostic = agnostic != 0;
  p.class_filter = class_filter;
  p.max_det = max_det;
  p.max_nms = max_nms;
  p.det = det;
  p.count = count;
  const size_t hdr = ((static_cast<size_t>(batch) * 4 + 255) / 256) * 256;
  p.ncand = static_cast<int32_t*>(workspace);
  p.keys = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) + hdr);
  p.cap_pow2 = pow2_ceil(static_cast<long long>(anchors) * nc);

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(p.ncand, 0, static_cast<size_t>(batch) * 4, st);
  if (e != cudaSuccess) return fail(MAF_E_CUDA, "nms: cudaMemsetAsync: %s", cudaGetErrorString(e));
  // follows a memset node, not a kernel: plain stream-ordered launch
  const size_t csmem = static_cast<size_t>(kCompactRows) * (5 + nc) * sizeof(float);
  if (csmem > 200 * 1024) return fail(MAF_E_ARG, "nms: nc=%d too large for the compaction tile", nc);
  if (csmem > 48 * 1024) {
    static SmemOptIn opt_in;  // per device, grows with nc
    rc = smem_opt_in(opt_in, nms_compact_kernel, static_cast<int>(csmem), "nms_compact");
    if (rc) return rc;
  }
  launch_pdl<false>(nms_compact_kernel, dim3(ceil_div(anchors, kCompactRows), batch), dim3(256), csmem, st, p);
  rc = check_launch("nms_compact kernel launch");
  if (rc) return rc;

  return launch_select(p, st);
}

// Second half of mafb200_nms alone: sort + greedy NMS over candidates that are ALREADY in `workspace` (written by
// mafb200_head_decode_detect, which fuses the decode with the threshold / compaction pass so that the
// [B, A, 5+nc] prediction tensor is never materialised on the serving path).  `boxes`: fp32 (cx, cy, w, h) of
// anchor a of image b at boxes[(b * anchors + a) * box_stride].
static int32_t nms_select_impl(const float* boxes, int32_t box_stride, int32_t batch, int32_t anchors, int32_t nc,
                               double iou_thres, int32_t agnostic, int32_t max_det, int32_t max_nms, float* det,
                               long long det_stride, int32_t* count, long long count_stride, void* workspace,
                               size_t workspace_bytes, void* stream);

extern "C" int32_t mafb200_nms_select(const float* boxes, int32_t box_stride, int32_t batch, int32_t anchors, int32_t nc,
                                      double iou_thres, int32_t agnostic, int32_t max_det, int32_t max_nms, float* det,
                                      int32_t* count, void* workspace, size_t workspace_bytes, void* stream) {
  return nms_select_impl(boxes, box_stride, batch, anchors, nc, iou_thres, agnostic, max_det, max_nms, det, 0, count, 0,
                         workspace, workspace_bytes, stream);
}

// Same, writing the PACKED row layout of the detection all-gather (maf_yolo_b200/dist.py): image b's detections at
// packed + b * row_floats (max_det * 6 floats) and its count, as int32 bits, in the float right behind them — so the
// rank's slice of the gather buffer is written in place and ONE collective moves both, without any copy kernel.
extern "C" int32_t mafb200_nms_select_packed(const float* boxes, int32_t box_stride, int32_t batch, int32_t anchors,
                                             int32_t nc, double iou_thres, int32_t agnostic, int32_t max_det,
                                             int32_t max_nms, float* packed, int32_t row_floats, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  if (!packed || row_floats < max_det * 6 + 1 || max_det <= 0)
    return fail(MAF_E_ARG, "nms_select_packed: row_floats=%d < max_det * 6 + 1", row_floats);
  return nms_select_impl(boxes, box_stride, batch, anchors, nc, iou_thres, agnostic, max_det, max_nms, packed, row_floats,
                         reinterpret_cast<int32_t*>(packed) + static_cast<size_t>(max_det) * 6, row_floats, workspace,
                         workspace_bytes, stream);
}

static int32_t nms_select_impl(const float* boxes, int32_t box_stride, int32_t batch, int32_t anchors, int32_t nc,
                               double iou_thres, int32_t agnostic, int32_t max_det, int32_t max_nms, float* det,
                               long long det_stride, int32_t* count, long long count_stride, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (!boxes || !det || !count || !workspace) return fail(MAF_E_ARG, "nms_select: null pointer");
  if (batch <= 0 || anchors <= 0 || nc <= 0 || box_stride < 4)
    return fail(MAF_E_ARG, "nms_select: bad shape B=%d A=%d nc=%d stride=%d", batch, anchors, nc, box_stride);
  if (!(iou_thres >= 0.0 && iou_thres <= 1.0)) return fail(MAF_E_ARG, "nms_select: iou_thres must be in [0,1], got %g", iou_thres);
  if (max_det <= 0 || max_det > kMaxDetCap) return fail(MAF_E_ARG, "nms_select: max_det=%d (1..%d)", max_det, kMaxDetCap);
  if (max_nms <= 0) return fail(MAF_E_ARG, "nms_select: max_nms=%d", max_nms);
  if (batch > 65535) return fail(MAF_E_ARG, "nms_select: batch %d > 65535", batch);
  if (workspace_bytes < mafb200_nms_workspace_bytes(batch, anchors, nc))
    return fail(MAF_E_WORKSPACE, "nms_select: workspace %zu < required %zu", workspace_bytes,
                mafb200_nms_workspace_bytes(batch, anchors, nc));
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(MAF_E_ALIGN, "nms_select: workspace must be 256-B aligned");
  int32_t rc = require_sm100();
  if (rc) return rc;
  NmsParams p;
  memset(&p, 0, sizeof(p));
  p.box = boxes;
  p.box_stride = box_stride;
  p.B = batch;
  p.A = anchors;
  p.nc = nc;
  p.iou = iou_thres;
  p.agnostic = agnostic != 0;
  p.max_det = max_det;
  p.max_nms = max_nms;
  p.det = det;
  p.count = count;
  p.det_stride = det_stride;
  p.count_stride = count_stride;
  const size_t hdr = ((static_cast<size_t>(batch) * 4 + 255) / 256) * 256;
  p.ncand = static_cast<int32_t*>(workspace);
  p.keys = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) + hdr);
  p.cap_pow2 = pow2_ceil(static_cast<long long>(anchors) * nc);
  return launch_select(p, static_cast<cudaStream_t>(stream));
}
