// K1 / K2 — convolution as (implicit) GEMM on tcgen05 tensor cores.
//
//   D[M, N] = act( A[M, K] * W[N, K]^T + bias )       fp16 operands, fp32 accumulate in TMEM
//
// K1 (1x1 conv, yolov6/layers/common.py:29-50): A = NHWC activations, M = n*h*w pixels, K = Cin.
//     Up to four sources are walked back-to-back along K, which is the reference's
//     Concat -> Conv(1x1) (the MAFPN fusion stages, configs/yaml/MAF-YOLO-n.yaml:17-42, and the
//     concat inside RepHDW / MPRep / SPPF, common.py:944,791,129) without materialising the concat.
// K2 (3x3 stride-2 pad-1 conv, common.py:76-83,206): same core, the A tile of tap (ky,kx) is
//     fetched by an im2col-mode TMA descriptor (traversal stride 2, bounding-box corners -1/-1,
//     zero fill = the padding), K = 9 * Cin.
//
// Persistent, warp-specialised CTAs (2 per SM); a CTA owns ONE column tile (n) and loops over row
// tiles (m) of 128 pixels:
//   warp 0 / lane 0 : TMA producer.  When the whole weight panel of the column tile fits
//                     (K_packed * tile_n * 2 B <= 48 KB — every layer of the N variant but four) it is
//                     loaded ONCE and stays resident; the ring then carries only A tiles (16 KB each),
//                     so the producer runs 3-6 row tiles ahead of the tensor core.  Otherwise A and W
//                     k-blocks stream together through the ring.
//   warp 1 / lane 0 : tcgen05.mma issuer (UMMA 128 x tile_n x 16, 4 per k-block) into one of TWO TMEM
//                     accumulators; commits free the ring slot / publish the accumulator.
//   warps 2..9      : epilogue of tile i while tile i+1 is loaded and multiplied: tcgen05.ld 32x32b.x32
//                     (one output row per thread) -> +bias -> activation (fast-math SiLU) -> fp16 ->
//                     4 x 16-byte global stores per 32 columns (64 contiguous bytes per thread), plus
//                     the x2 nearest-upsampled copy for the two layers that feed an nn.Upsample.
//                     (A shared-memory staged TMA tensor store was measured SLOWER: one-row-per-thread
//                     writes into a dense row-major staging tile are 8-32-way bank conflicted.)
// K is tiny here (1-12 blocks of 64) and N <= 128 per tile: the kernel is HBM/epilogue-bound, not
// tensor-bound, so the design keeps loads deep in flight and writes whole 32-byte sectors.
//
// Measured and rejected for the 3x3 producer (profiles/r01_notes.md): (1) a software im2col by four LSU
// gather warps (8 lanes x 16 B per pixel row, SW128 layout written by hand): 2.6x slower than the im2col
// TMA (1132 vs 432 us per forward) because each slot is a dependent load -> st.shared -> fence round trip;
// (2) a 1-CTA/SM 216 KB shape keeping the 54-72 KB weight panels of the small-Cin layers resident: slower.
// The im2col TMA itself costs ~0.35-0.55 us per 128-pixel box almost independently of the channel count
// (one L2 request per pixel row), which is what bounds the 3x3 layers today.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"

namespace mafb200 {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 32 * (2 + kEpiWarps);  // TMA warp + MMA warp + epilogue warps

struct GemmParams {
  CUtensorMap tmA[MAF_MAX_SRC];
  CUtensorMap tmW;
  const float* bias;
  __half* out;
  __half* out2;  // x2-upsampled copy or nullptr
  int32_t M, N, tile_n;
  int32_t out_ld, out2_ld;
  int32_t out_h, out_w;  // spatial size of the output map (for out2 and im2col tile origin)
  int32_t nsrc;
  int32_t kblocks[MAF_MAX_SRC];  // 1x1: 64-wide K blocks of each source; 3x3: kblocks[0] = blocks per tap
  int32_t total_kb;
  int32_t act;
  int32_t stages;      // ring slots
  int32_t w_resident;  // 1: weight panel loaded once, ring slots hold A only
  int32_t tmem_cols;   // total TMEM columns allocated: two accumulator buffers
  int32_t n_tiles;
  int32_t m_stride;    // row-tile step of a CTA = gridDim.x / n_tiles
  int32_t st256;       // 1: every 16-column group of every output row starts 32-B aligned -> 256-bit stores
  int32_t pair;        // 3x3 s2 over PIXEL PAIRS (Cin <= 32): 6 taps (ky, pair offset) instead of 9 (ky, kx)
  uint32_t idesc;
  // ---- K7: head-prediction epilogues (mafb200_head_pred; kEpi 1 = DFL box decode, 2 = class sigmoid / filter) ----
  float* pred;                 // [B, A, no] fp32 or nullptr
  int32_t cls_off;             // first class column inside a row of `pred` (5 for the eval tensor, 0 for train-form scores)
  int32_t raw_reg;             // kEpi 1: store the 68 raw DFL logits (train-form pred_distri) instead of decoding them
  float* boxes;                // [B, A, 4] fp32 or nullptr
  const maf_detect_cfg* cfg;   // device memory; non-null = emit NMS candidates (kEpi 2)
  int32_t* ncand;              // [B]
  unsigned long long* keys;    // [B][cap_pow2]
  long long cap_pow2;
  int32_t lvl_L, lvl_w;        // anchors (= pixels) per image of this level, map width
  int32_t anchor_off, total_anchors, no, nc;
  float lvl_stride;
};

// 256-bit global store (SASS STG.E.ENL2.256): one full 32-byte sector per instruction
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// bias + activation + fp16 pack of 32 accumulator columns, specialised per activation so the element loop
// has no per-element dispatch, on the packed fp32x2 pipe (FFMA2 / FADD2: two IEEE results per issue slot,
// bit-identical to the scalar forms).  SiLU = h + h*tanh(h) with h = 0.5*acc + 0.5*bias: per PAIR of
// elements FFMA2, 2 x MUFU.TANH, FFMA2, F2FP — 2.5 issue slots per element instead of 4.5.
template <int kAct>
__device__ __forceinline__ void epi_pack32(const uint32_t (&r)[32], const float* __restrict__ bias,
                                           const float* __restrict__ half_bias, uint32_t (&pk)[16]) {
  const float2 half2c = make_float2(0.5f, 0.5f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 acc = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
    float2 v;
    if (kAct == ACT_SILU) {
      const float2 h = ffma2(acc, half2c, *reinterpret_cast<const float2*>(half_bias + 2 * j));
      v = ffma2(h, make_float2(tanh_approx(h.x), tanh_approx(h.y)), h);
    } else {
      v = fadd2(acc, *reinterpret_cast<const float2*>(bias + 2 * j));
      if (kAct == ACT_RELU) {
        v.x = fmaxf(v.x, 0.0f);
        v.y = fmaxf(v.y, 0.0f);
      } else if (kAct == ACT_SIGMOID) {
        v.x = __fdividef(1.0f, 1.0f + __expf(-v.x));
        v.y = __fdividef(1.0f, 1.0f + __expf(-v.y));
      }
    }
    pk[j] = pack_half2(v.x, v.y);
  }
}

struct GemmSmem {
  uint8_t* wpanel;
  uint8_t* ring;
  uint64_t *full_bar, *empty_bar, *tmem_full_bar, *tmem_empty_bar, *w_bar;
  float* bias;
  int slot_bytes, b_bytes;
  bool w_res;
};

// ---- MMA issuer (one elected lane) -------------------------------------------------------------
__device__ __forceinline__ void gemm_mma_loop(const GemmParams& p, const GemmSmem& sm, uint32_t tmem_base,
                                              uint32_t acc_cols, int mt0, int mt_step, int m_tiles) {
  const int total_kb = p.total_kb, stages = p.stages;
  if (sm.w_res) {
    mbar_wait(sm.w_bar, 0);
    tc_fence_after_sync();
  }
  int kb = 0;
  int it = 0;
  for (int mt = mt0; mt < m_tiles; mt += mt_step, ++it) {
    const int as = it & 1;
    mbar_wait(&sm.tmem_empty_bar[as], ((it >> 1) & 1) ^ 1);  // epilogue drained this accumulator
    tc_fence_after_sync();
    const uint32_t tmem_d = tmem_base + as * acc_cols;
    for (int k2 = 0; k2 < total_kb; ++k2, ++kb) {
      const int slot = kb % stages;
      const uint32_t phase = (kb / stages) & 1;
      mbar_wait(&sm.full_bar[slot], phase);
      tc_fence_after_sync();
      const uint32_t sa = smem_u32(sm.ring + slot * sm.slot_bytes);
      const uint32_t sb = sm.w_res ? smem_u32(sm.wpanel + k2 * sm.b_bytes) : sa + kABytes;
      const uint64_t da = umma_smem_desc_sw128(sa);
      const uint64_t db = umma_smem_desc_sw128(sb);
#pragma unroll
      for (int k = 0; k < kBlockK / 16; ++k) {
        // advance 16 fp16 = 32 B inside the 128-B swizzle row: +2 in the (addr >> 4) field
        tc_mma_f16(tmem_d, da + 2 * k, db + 2 * k, p.idesc, (k2 | k) != 0 ? 1u : 0u);
      }
      tc_commit(&sm.empty_bar[slot]);
    }
    tc_commit(&sm.tmem_full_bar[as]);
  }
}

// ---- epilogue (8 warps; `ew` = 0..7 index of this warp among them) ---------------------------------
__device__ __forceinline__ void gemm_epilogue_loop(const GemmParams& p, const GemmSmem& sm, uint32_t tmem_base,
                                                   uint32_t acc_cols, int mt0, int mt_step, int m_tiles, int n0,
                                                   int warp, int lane, int ew) {
  const float* s_bias = sm.bias;
  const float* s_hbias = sm.bias + p.tile_n + 32;  // 0.5 * bias (SiLU path); +32: groups may read past tile_n
  const int quarter = warp & 3;     // TMEM lanes this warp may access: 32 * (warp id % 4)
  const int col_group = ew >> 2;    // even / odd 32-column groups
  const int row = quarter * 32 + lane;
  const int act = p.act;
  const size_t up_dx = p.out2_ld;
  const size_t up_dy = static_cast<size_t>(2) * p.out_w * p.out2_ld;
  int it = 0;
  for (int mt = mt0; mt < m_tiles; mt += mt_step, ++it) {
    const int as = it & 1;
    const int m = mt * kBlockM + row;
    const bool row_ok = m < p.M;
    mbar_wait(&sm.tmem_full_bar[as], (it >> 1) & 1);
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base + as * acc_cols + (static_cast<uint32_t>(quarter * 32) << 16);
    __half* orow = p.out + static_cast<size_t>(row_ok ? m : 0) * p.out_ld;
    __half* urow = nullptr;
    if (p.out2 != nullptr && row_ok) {
      const int hw = p.out_h * p.out_w;
      const int img = m / hw;
      const int rem = m - img * hw;
      const int y = rem / p.out_w;
      const int x = rem - y * p.out_w;
      urow = p.out2 + ((static_cast<size_t>(img) * 2 * p.out_h + 2 * y) * (2 * p.out_w) + 2 * x) * p.out2_ld;
    }
#pragma unroll 1
    for (int c = col_group * 32; c < p.tile_n; c += 64) {
      uint32_t r[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged here
      tmem_ld_32x32b_x32(taddr + c, r);
      tmem_ld_wait();
      const int n = n0 + c;
      if (row_ok && n < p.N) {
        // 32 independent bias+activation chains first (MUFU latency overlaps), then pack + store
        const int valid = min(32, p.tile_n - c);  // 32 or 16 (tile_n % 16 == 0)
        uint32_t pk[16];
        if (act == ACT_SILU) epi_pack32<ACT_SILU>(r, s_bias + c, s_hbias + c, pk);
        else if (act == ACT_NONE) epi_pack32<ACT_NONE>(r, s_bias + c, s_hbias + c, pk);
        else if (act == ACT_RELU) epi_pack32<ACT_RELU>(r, s_bias + c, s_hbias + c, pk);
        else epi_pack32<ACT_SIGMOID>(r, s_bias + c, s_hbias + c, pk);
        if (p.st256 && n + valid <= p.N) {
          // full 32-byte sectors: 2 (or 1) x 256-bit stores for this thread's 64 (32) contiguous bytes
          st_global_256(orow + n, pk);
          if (valid > 16) st_global_256(orow + n + 16, pk + 8);
          if (urow != nullptr) {  // x2 nearest-upsampled copy: the same sectors to the four target pixels
            st_global_256(urow + n, pk);
            st_global_256(urow + up_dx + n, pk);
            st_global_256(urow + up_dy + n, pk);
            st_global_256(urow + up_dy + up_dx + n, pk);
            if (valid > 16) {
              st_global_256(urow + n + 16, pk + 8);
              st_global_256(urow + up_dx + n + 16, pk + 8);
              st_global_256(urow + up_dy + n + 16, pk + 8);
              st_global_256(urow + up_dy + up_dx + n + 16, pk + 8);
            }
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int ng = n + 8 * g;
            if (8 * g >= valid) break;
            if (ng + 8 <= p.N) {
              const uint4 q = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
              *reinterpret_cast<uint4*>(orow + ng) = q;
              if (urow != nullptr) {
                *reinterpret_cast<uint4*>(urow + ng) = q;
                *reinterpret_cast<uint4*>(urow + up_dx + ng) = q;
                *reinterpret_cast<uint4*>(urow + up_dy + ng) = q;
                *reinterpret_cast<uint4*>(urow + up_dy + up_dx + ng) = q;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (ng + j < p.N) {
                  const uint32_t w2 = pk[4 * g + (j >> 1)];
                  const __half hv = __ushort_as_half(static_cast<unsigned short>((j & 1) ? (w2 >> 16) : (w2 & 0xffffu)));
                  orow[ng + j] = hv;
                  if (urow != nullptr) {
                    urow[ng + j] = hv;
                    urow[up_dx + ng + j] = hv;
                    urow[up_dy + ng + j] = hv;
                    urow[up_dy + up_dx + ng + j] = hv;
                  }
                }
              }
            }
          }
        }
      }
    }
    // all of this warp's TMEM reads of accumulator `as` are complete (wait::ld above): hand it back
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.tmem_empty_bar[as]);
  }
}

// ---- K7 epilogues: the reg_pred / cls_pred GEMMs of Head_DepthUni (yolov6/layers/common.py:1325-1336) finish the
// detect path in registers.  One thread owns one accumulator row = one anchor, so the whole eval branch of
// Detect_yaml.forward (yolov6/models/yolo.py:355-396) is per-thread work on the fp32 accumulators — the 68 reg logits
// and 80 class logits are never rounded to fp16 and never stored:
//   kEpi 1 (reg_pred): softmax over the 17 bins of each side, expectation with proj = 0..16 (yolo.py:377-378), anchor
//       point (x + .5, y + .5) (anchor_generator.py:11-25), dist2bbox 'xywh' (general.py:29-40), x stride (yolo.py:389)
//       -> pred[b, a, 0:5] = (cx, cy, w, h, 1) and / or boxes[b, a, 0:4].
//       The packed weight rows are permuted (ops.pack_head_reg) so that a side never straddles a TMEM load:
//       column j < 64: side j / 16, bin j % 16;  64 <= j < 68: side j - 64, bin 16;  68..79: zero padding.
//   kEpi 2 (cls_pred): sigmoid (common.py:1332) -> pred[b, a, 5:] and / or the candidate filter of
//       non_max_suppression (yolov6/utils/nms.py:48-84: score = cls * obj with obj == 1) -> 64-bit keys.
// The two warp sets (epilogue warps 0-3 / 4-7) take alternate accumulator stages, 4 arrivals free a stage.
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.0f, __fadd_rn(1.0f, __expf(-v))); }

__device__ __forceinline__ float dfl_side(const float (&v)[17]) {
  float mx = v[0];
#pragma unroll
  for (int i = 1; i < 17; ++i) mx = fmaxf(mx, v[i]);
  float s = 0.f, e = 0.f;
#pragma unroll
  for (int i = 0; i < 17; ++i) {
    const float ex = __expf(v[i] - mx);
    s += ex;
    e = fmaf(static_cast<float>(i), ex, e);
  }
  return e / s;
}

template <int kEpi>
__device__ __forceinline__ void head_epilogue_loop(const GemmParams& p, const GemmSmem& sm, uint32_t tmem_base,
                                                   uint32_t acc_cols, int mt0, int mt_step, int m_tiles, int warp,
                                                   int lane, int ew) {
  const float* s_bias = sm.bias;
  const int quarter = warp & 3;
  const int set = ew >> 2;
  const int row = quarter * 32 + lane;
  float conf = 2.0f, skip_below = INFINITY;
  int multi_label = 0;
  const uint8_t* filt = nullptr;
  if (kEpi == 2 && p.cfg != nullptr) {
    conf = p.cfg->conf;
    skip_below = p.cfg->skip_below;
    multi_label = p.cfg->multi_label;
    if (p.cfg->has_filter) filt = p.cfg->class_filter;
  }
  int it = 0;
  for (int mt = mt0; mt < m_tiles; mt += mt_step, ++it) {
    const int as = it & 1;
    if (as != set) continue;  // the other warp set drains this accumulator stage
    const int m = mt * kBlockM + row;
    const bool row_ok = m < p.M;
    mbar_wait(&sm.tmem_full_bar[as], (it >> 1) & 1);
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base + as * acc_cols + (static_cast<uint32_t>(quarter * 32) << 16);
    const int b = (row_ok ? m : 0) / p.lvl_L;
    const int al = (row_ok ? m : 0) - b * p.lvl_L;
    const size_t grow = static_cast<size_t>(b) * p.total_anchors + p.anchor_off + al;  // row of pred / boxes
    if (kEpi == 1) {
      float dist[4];
      uint32_t t16[16];
      __syncwarp();
      tmem_ld_32x32b_x16(taddr + 64, t16);  // bin 16 of the four sides (+ padding columns)
      tmem_ld_wait();
      float last[4];
#pragma unroll
      for (int sd = 0; sd < 4; ++sd) last[sd] = __uint_as_float(t16[sd]) + s_bias[64 + sd];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld_32x32b_x32(taddr + 32 * half, r);
        tmem_ld_wait();
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          float v[17];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[16 * s2 + i]) + s_bias[32 * half + 16 * s2 + i];
          v[16] = last[2 * half + s2];
          if (p.raw_reg) {
            // train-form output (Detect_yaml.forward train / val_loss branch, yolo.py:333-346): reg_output flattened to
            // [B, A, 68] = side-major, bin-minor, fp32, no softmax
            if (row_ok) {
              float* o = p.pred + grow * 68 + (2 * half + s2) * 17;
#pragma unroll
              for (int i = 0; i < 17; ++i) o[i] = v[i];
            }
          } else {
            dist[2 * half + s2] = dfl_side(v);
          }
        }
      }
      if (row_ok && !p.raw_reg) {
        const int gy = al / p.lvl_w, gx = al - gy * p.lvl_w;
        const float ax = static_cast<float>(gx) + 0.5f, ay = static_cast<float>(gy) + 0.5f;
        const float x1 = ax - dist[0], y1 = ay - dist[1], x2 = ax + dist[2], y2 = ay + dist[3];
        const float st = p.lvl_stride;
        const float cx = ((x1 + x2) / 2.0f) * st, cy = ((y1 + y2) / 2.0f) * st, bw = (x2 - x1) * st, bh = (y2 - y1) * st;
        if (p.pred != nullptr) {
          float* dst = p.pred + grow * p.no;
          dst[0] = cx;
          dst[1] = cy;
          dst[2] = bw;
          dst[3] = bh;
          dst[4] = 1.0f;
        }
        if (p.boxes != nullptr) *reinterpret_cast<float4*>(p.boxes + grow * 4) = make_float4(cx, cy, bw, bh);
      }
    } else {
      float* dst = (p.pred != nullptr && row_ok) ? p.pred + grow * p.no + p.cls_off : nullptr;
      const bool emit = p.cfg != nullptr && row_ok;
      const int a = p.anchor_off + al;
      float best = -INFINITY;  // single-label: best class probability so far (first maximum)
      int best_c = 0;
#pragma unroll 1
      for (int c = 0; c < p.tile_n; c += 32) {
        uint32_t r[32];
        __syncwarp();
        if (p.tile_n - c >= 32) {
          tmem_ld_32x32b_x32(taddr + c, r);
        } else {
          uint32_t q[16];
          tmem_ld_32x32b_x16(taddr + c, q);
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = q[i], r[16 + i] = 0u;
        }
        tmem_ld_wait();
        const int valid = min(32, p.N - c);
        uint32_t bits = 0;  // multi-label: columns of this chunk that are candidates of this thread's row
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j >= valid) break;
          const float v = __fadd_rn(__uint_as_float(r[j]), s_bias[c + j]);
          if (dst != nullptr) dst[c + j] = sigmoid_fast(v);
          if (emit && v > skip_below) {
            const float s = sigmoid_fast(v);  // score = cls * obj, obj == 1 (nms.py:69)
            if (multi_label) {
              if (s > conf && (filt == nullptr || filt[c + j] != 0)) bits |= 1u << j;
            } else if (s > best) {
              best = s;
              best_c = c + j;
            }
          }
        }
        // Emission, warp-aggregated: one atomicAdd per warp and chunk instead of one per candidate (random-weight S / M
        // heads put 10-50 k candidates on an image; the per-candidate atomics on the image's single counter made this
        // kernel 5x slower than the GEMM itself).  The common case — no candidate in the chunk — is one vote.
        if (__any_sync(0xffffffffu, bits != 0)) {
          const int cnt = __popc(bits);
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
          }
          const int total = __shfl_sync(0xffffffffu, incl, 31);
          long long base;
          // (every lane must execute the shuffles: no short-circuit evaluation around them)
          const int b_lane0 = __shfl_sync(0xffffffffu, b, 0);
          const int lane0_ok = __shfl_sync(0xffffffffu, row_ok ? 1 : 0, 0);
          const bool same_image = __all_sync(0xffffffffu, !row_ok || b == b_lane0) != 0;
          if (same_image && lane0_ok) {
            int b0 = 0;
            if (lane == 0) b0 = atomicAdd(&p.ncand[b], total);  // every candidate row of the warp belongs to image b
            base = static_cast<long long>(__shfl_sync(0xffffffffu, b0, 0)) + (incl - cnt);
          } else {  // the 32 rows straddle two images (or lane 0 is past the end): per-lane reservation
            base = cnt > 0 ? atomicAdd(&p.ncand[b], cnt) : 0;
          }
          if (bits != 0) {
            unsigned long long* kb = p.keys + static_cast<size_t>(b) * p.cap_pow2;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (bits & (1u << j)) {
                const float s = sigmoid_fast(__fadd_rn(__uint_as_float(r[j]), s_bias[c + j]));  // the same value as above
                if (base < p.cap_pow2)
                  kb[base] = (static_cast<unsigned long long>(~__float_as_uint(s)) << 32) |
                             static_cast<unsigned long long>(static_cast<unsigned>(a) * p.nc + (c + j));
                ++base;
              }
            }
          }
        }
      }
      if (emit && !multi_label && best > conf && (filt == nullptr || filt[best_c] != 0)) {
        const long long slot = atomicAdd(&p.ncand[b], 1);
        if (slot < p.cap_pow2)
          p.keys[static_cast<size_t>(b) * p.cap_pow2 + slot] =
              (static_cast<unsigned long long>(~__float_as_uint(best)) << 32) |
              static_cast<unsigned long long>(static_cast<unsigned>(a) * p.nc + best_c);
      }
    }
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.tmem_empty_bar[as]);
  }
}

template <bool kIm2col, int kEpi = 0>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // keep the shared address space (no generic-pointer arithmetic): pad up to the 1024-B swizzle alignment
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_kb = p.total_kb;
  const int b_bytes = p.tile_n * kBlockK * 2;  // one k-block of the W tile
  const bool w_res = p.w_resident != 0;
  const int slot_bytes = w_res ? kABytes : kABytes + b_bytes;
  const int stages = p.stages;

  // layout: [W panel (resident mode)] [ring slots] [barriers] [bias]
  uint8_t* s_wpanel = smem;
  uint8_t* s_ring = smem + (w_res ? total_kb * b_bytes : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ring + stages * slot_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint64_t* w_bar = tmem_empty_bar + 2;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);  // [tile_n]
  const GemmSmem sm{s_wpanel, s_ring, full_bar, empty_bar, tmem_full_bar, tmem_empty_bar, w_bar, s_bias, slot_bytes, b_bytes, w_res};

  const int m_tiles = ceil_div(p.M, kBlockM);
  const int nt = blockIdx.x % p.n_tiles;  // this CTA's column tile
  const int n0 = nt * p.tile_n;
  const int mt0 = blockIdx.x / p.n_tiles;
  const int mt_step = p.m_stride;

  // ---- one-time setup -------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < (kIm2col ? 1 : p.nsrc); ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmW);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kEpi == 0 ? kEpiWarps : kEpiWarps / 2);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.tile_n + 32; i += blockDim.x) {
    const float bv = i < p.tile_n ? p.bias[n0 + i] : 0.0f;
    s_bias[i] = bv;
    s_bias[p.tile_n + 32 + i] = 0.5f * bv;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_cols = p.tmem_cols >> 1;  // columns of one accumulator buffer
  // TMEM is allocated: the next kernel of the stream may now be scheduled behind this one (PDL)
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp == 0) {
    // ---- TMA producer ---------------------------------------------------------------------------
    if (lane == 0 && mt0 < m_tiles) {
      if (w_res) {  // weights are constants: fetched while the previous kernel may still be running
        mbar_arrive_expect_tx(w_bar, total_kb * b_bytes);
        for (int k = 0; k < total_kb; ++k) tma_load_2d(s_wpanel + k * b_bytes, &p.tmW, w_bar, k * kBlockK, n0);
      }
      // every global read of activations (and, through the MMA -> epilogue chain, every store) is ordered
      // after the prerequisite kernels by this wait
      pdl_wait();
      int kb = 0;  // running ring position across tiles
      for (int mt = mt0; mt < m_tiles; mt += mt_step) {
        const int m0 = mt * kBlockM;
        int q0 = 0, p0 = 0, img = 0;
        if (kIm2col) {
          const int hw = p.out_h * p.out_w;
          img = m0 / hw;
          const int rem = m0 - img * hw;
          p0 = rem / p.out_w;
          q0 = rem - p0 * p.out_w;
        }
        int wk = 0;  // column in the packed weights
        const int n_outer = kIm2col ? (p.pair ? 6 : 9) : p.nsrc;
        for (int o = 0; o < n_outer; ++o) {
          const int nblk = kIm2col ? p.kblocks[0] : p.kblocks[o];
          for (int j = 0; j < nblk; ++j, ++kb, wk += kBlockK) {
            const int slot = kb % stages;
            const uint32_t phase = (kb / stages) & 1;
            mbar_wait(&empty_bar[slot], phase ^ 1);
            uint8_t* sa = s_ring + slot * slot_bytes;
            mbar_arrive_expect_tx(&full_bar[slot], slot_bytes);
            if (kIm2col) {
              if (p.pair) {
                // the tensor is viewed as [N][H][W/2][2*ld]: input columns 2x-1, 2x, 2x+1 of output x are the
                // second pixel of pair x-1 and both pixels of pair x -> base pair x-1, offsets {0, 1}, x stride 1
                const int ky = o >> 1, po = o & 1;
                tma_load_im2col_4d(sa, &p.tmA[0], &full_bar[slot], 0, q0 - 1, 2 * p0 - 1, img,
                                   static_cast<uint16_t>(po), static_cast<uint16_t>(ky));
              } else {
                const int ky = o / 3, kx = o - ky * 3;
                tma_load_im2col_4d(sa, &p.tmA[0], &full_bar[slot], j * kBlockK, 2 * q0 - 1, 2 * p0 - 1, img,
                                   static_cast<uint16_t>(kx), static_cast<uint16_t>(ky));
              }
            } else {
              tma_load_2d(sa, &p.tmA[o], &full_bar[slot], j * kBlockK, m0);
            }
            if (!w_res) tma_load_2d(sa + kABytes, &p.tmW, &full_bar[slot], wk, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && mt0 < m_tiles) gemm_mma_loop(p, sm, tmem_base, acc_cols, mt0, mt_step, m_tiles);
  } else if (kEpi == 0) {
    gemm_epilogue_loop(p, sm, tmem_base, acc_cols, mt0, mt_step, m_tiles, n0, warp, lane, warp - 2);
  } else {
    head_epilogue_loop<kEpi>(p, sm, tmem_base, acc_cols, mt0, mt_step, m_tiles, warp, lane, warp - 2);
  }

  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int32_t encode_2d(CUtensorMap* tm, const void* base, cuuint64_t inner, cuuint64_t rows, cuuint64_t row_bytes,
                         cuuint32_t box_inner, cuuint32_t box_rows, CUtensorMapSwizzle swz, CUtensorMapL2promotion l2,
                         const char* what) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(MAF_E_ARCH, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MAF_E_CUDA, "cuTensorMapEncodeTiled(%s: inner=%llu rows=%llu pitch=%llu) failed: %d", what,
                (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_bytes, (int)r);
  return MAF_OK;
}

static int32_t encode_w_map(CUtensorMap* tm, const void* w, int k_packed, int rows, int tile_n) {
  return encode_2d(tm, w, k_packed, rows, static_cast<cuuint64_t>(k_packed) * 2, kBlockK, tile_n,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "W");
}

static int32_t encode_a_map_2d(CUtensorMap* tm, const maf_tensor* t) {
  const cuuint64_t M = static_cast<cuuint64_t>(t->n) * t->h * t->w;
  return encode_2d(tm, t->ptr, t->c, M, static_cast<cuuint64_t>(t->c_stride) * 2, kBlockK, kBlockM,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "A");
}

static int32_t encode_a_map_im2col(CUtensorMap* tm, const maf_tensor* t) {
  EncodeIm2colFn enc = encode_im2col_fn();
  if (!enc) return fail(MAF_E_ARCH, "cuTensorMapEncodeIm2col entry point not available");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(t->c), static_cast<cuuint64_t>(t->w), static_cast<cuuint64_t>(t->h),
                        static_cast<cuuint64_t>(t->n)};
  const cuuint64_t px = static_cast<cuuint64_t>(t->c_stride) * 2;
  cuuint64_t strides[3] = {px, px * t->w, px * t->w * t->h};
  // 3x3, pad 1, dilation 1: lower corner = -pad, upper corner = pad - (k-1) = -1 (W, H order).
  int lower[2] = {-1, -1};
  int upper[2] = {-1, -1};
  cuuint32_t estr[4] = {1, 2, 2, 1};  // traversal stride 2 in W and H
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, t->ptr, dims, strides, lower, upper, kBlockK, kBlockM, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MAF_E_CUDA, "cuTensorMapEncodeIm2col(c=%d w=%d h=%d n=%d) failed: %d", t->c, t->w, t->h, t->n, (int)r);
  return MAF_OK;
}

// Pixel-pair view for the 3x3 s2 conv of a narrow map (2 * c_stride <= 64): dims {2*ld, W/2, H, N}, x traversal
// stride 1 over pairs (kernel extent 2: pair offsets {0,1}, one pair of left padding), y as before.  The im2col TMA
// costs ~0.4 us per 128-pixel box whatever the channel count, so 6 dense boxes per tile replace 9 sparse ones.
static int32_t encode_a_map_im2col_pair(CUtensorMap* tm, const maf_tensor* t) {
  EncodeIm2colFn enc = encode_im2col_fn();
  if (!enc) return fail(MAF_E_ARCH, "cuTensorMapEncodeIm2col entry point not available");
  const cuuint64_t pair_ch = static_cast<cuuint64_t>(2) * t->c_stride;
  cuuint64_t dims[4] = {pair_ch, static_cast<cuuint64_t>(t->w / 2), static_cast<cuuint64_t>(t->h),
                        static_cast<cuuint64_t>(t->n)};
  const cuuint64_t px = static_cast<cuuint64_t>(t->c_stride) * 2;
  cuuint64_t strides[3] = {2 * px, px * t->w, px * t->w * t->h};
  int lower[2] = {-1, -1};
  int upper[2] = {-1, -1};  // W: pad_right 0 - (2 - 1); H: pad 1 - (3 - 1)
  cuuint32_t estr[4] = {1, 1, 2, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, t->ptr, dims, strides, lower, upper,
                   static_cast<cuuint32_t>(pair_ch), kBlockM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MAF_E_CUDA, "cuTensorMapEncodeIm2col(pair view, ld=%d w=%d h=%d n=%d) failed: %d", t->c_stride, t->w,
                t->h, t->n, (int)r);
  return MAF_OK;
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

static int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

enum GemmMode { kModeTma2d = 0, kModeIm2colTma = 1 };

template <int kMode, int kEpi = 0>
static int32_t launch_gemm(GemmParams& p, int n_tiles, int total_kb, cudaStream_t stream) {
  const int b_bytes = p.tile_n * kBlockK * 2;
  const int m_tiles = ceil_div(p.M, kBlockM);
  // Two shapes of persistent CTA (2 per SM): weight panel resident + A-only ring when it fits, else A and W
  // k-blocks stream together.  (A 1-CTA/SM ~216 KB shape for the 54-72 KB panels of the small-Cin 3x3
  // layers was measured slower: those layers are bound by the im2col TMA's per-pixel-row request rate —
  // ~0.35 us per 128-row im2col box regardless of Cin — not by weight re-streaming or ring depth.)
  const int panel = total_kb * b_bytes;
  static const int ctas_per_sm = [] {
    const char* v = getenv("MAFB200_GEMM_CTAS_PER_SM");  // experiment knob: 1 leaves room for a co-resident kernel
    const int n = v ? atoi(v) : 2;
    return n >= 1 && n <= 2 ? n : 2;
  }();
  const int budget = 104 * 1024;
  int stages;
  p.w_resident = (panel + 3 * kABytes <= budget) ? 1 : 0;
  if (p.w_resident) {
    stages = (budget - panel) / kABytes;
    const int want = total_kb * 4 > 3 ? total_kb * 4 : 3;  // ~4 row tiles in flight
    if (stages > want) stages = want;
    if (stages > 12) stages = 12;
  } else {
    stages = budget / (kABytes + b_bytes);
    if (stages > 2 * total_kb) stages = 2 * total_kb;
    if (stages > 6) stages = 6;
  }
  {
    static const int cap = [] {
      const char* v = getenv("MAFB200_GEMM_MAX_STAGES");  // experiment knob (no effect measured for 3..12)
      return v ? atoi(v) : 0;
    }();
    if (cap > 0 && stages > cap) stages = cap;
  }
  if (stages < 1) stages = 1;
  // persistent grid: a multiple of n_tiles (each CTA owns one column tile)
  int per_n = (ctas_per_sm * sm_count()) / n_tiles;
  if (per_n < 1) per_n = 1;
  if (per_n > m_tiles) per_n = m_tiles;
  const int grid = per_n * n_tiles;
  const int slot_bytes = p.w_resident ? kABytes : kABytes + b_bytes;
  p.stages = stages;
  p.total_kb = total_kb;
  p.n_tiles = n_tiles;
  p.m_stride = per_n;
  p.tmem_cols = 2 * pow2_cols(p.tile_n);
  p.idesc = umma_idesc_f16(kBlockM, p.tile_n);
  p.st256 = ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0 && (p.out_ld % 16) == 0 && (p.tile_n % 16) == 0 &&
             (p.out2 == nullptr || ((reinterpret_cast<uintptr_t>(p.out2) & 31) == 0 && (p.out2_ld % 16) == 0)))
                 ? 1
                 : 0;
  const size_t smem = static_cast<size_t>(p.w_resident ? panel : 0) + static_cast<size_t>(stages) * slot_bytes +
                      (2 * stages + 5) * 8 + 16 + static_cast<size_t>(p.tile_n + 32) * 8 + 1024;
  {
    static SmemOptIn opt_in;  // per device (a process-wide flag skipped the opt-in on a second GPU)
    const int32_t rc_attr = smem_opt_in(opt_in, gemm_tc_kernel<kMode == kModeIm2colTma, kEpi>, 227 * 1024, "gemm");
    if (rc_attr) return rc_attr;
  }
  if (smem > 227 * 1024) return fail(MAF_E_ARG, "gemm: %zu B of shared memory needed", smem);
  launch_pdl(gemm_tc_kernel<kMode == kModeIm2colTma, kEpi>, dim3(grid), dim3(kGemmThreads), smem, stream, p);
  return check_launch(kEpi != 0 ? "head_pred kernel launch" : kMode == kModeTma2d ? "conv1x1 kernel launch" : "conv3x3s2 kernel launch");
}

}  // namespace mafb200

using namespace mafb200;

extern "C" int32_t mafb200_conv1x1(const maf_tensor* srcs, int32_t n_src, const void* w_packed, const float* bias,
                                   int32_t act, const maf_tensor* dst, const maf_tensor* dst_up2x, void* stream) {
  if (!srcs || n_src < 1 || n_src > MAF_MAX_SRC) return fail(MAF_E_ARG, "conv1x1: n_src=%d (1..%d)", n_src, MAF_MAX_SRC);
  if (!w_packed || !bias) return fail(MAF_E_ARG, "conv1x1: null weights/bias");
  if (!valid_f16_view(dst)) return fail(MAF_E_ARG, "conv1x1: bad dst tensor");
  if (!aligned_f16_view(dst)) return fail(MAF_E_ALIGN, "conv1x1: dst must be 16-B aligned with c_stride %% 8 == 0");
  if (act < MAF_ACT_NONE || act > MAF_ACT_SIGMOID) return fail(MAF_E_ARG, "conv1x1: bad act %d", act);
  if ((reinterpret_cast<uintptr_t>(w_packed) & 15) != 0) return fail(MAF_E_ALIGN, "conv1x1: w_packed not 16-B aligned");
  for (int s = 0; s < n_src; ++s) {
    const maf_tensor* t = &srcs[s];
    if (!valid_f16_view(t)) return fail(MAF_E_ARG, "conv1x1: bad src[%d]", s);
    if (!aligned_f16_view(t)) return fail(MAF_E_ALIGN, "conv1x1: src[%d] alignment", s);
    if (!same_nhw(t, dst)) return fail(MAF_E_ARG, "conv1x1: src[%d] n/h/w differ from dst", s);
  }
  if (dst_up2x) {
    if (!valid_f16_view(dst_up2x) || !aligned_f16_view(dst_up2x) || dst_up2x->n != dst->n ||
        dst_up2x->h != 2 * dst->h || dst_up2x->w != 2 * dst->w || dst_up2x->c != dst->c)
      return fail(MAF_E_ARG, "conv1x1: dst_up2x must be [n,2h,2w,c] fp16");
  }
  int32_t rc = require_sm100();
  if (rc) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  int chans[MAF_MAX_SRC];
  int total_kb = 0;
  for (int s = 0; s < n_src; ++s) {
    const maf_tensor* t = &srcs[s];
    rc = encode_a_map_2d(&p.tmA[s], t);
    if (rc) return rc;
    chans[s] = t->c;
    p.kblocks[s] = ceil_div(t->c, kBlockK);
    total_kb += p.kblocks[s];
  }
  int n_tiles = 0, tile_n = 0;
  mafb200_gemm_tiling(dst->c, &n_tiles, &tile_n);
  const int k_packed = mafb200_packed_k_1x1(chans, n_src);
  rc = encode_w_map(&p.tmW, w_packed, k_packed, n_tiles * tile_n, tile_n);
  if (rc) return rc;

  p.bias = bias;
  p.out = static_cast<__half*>(dst->ptr);
  p.out_ld = dst->c_stride;
  p.M = dst->n * dst->h * dst->w;
  p.N = dst->c;
  p.tile_n = tile_n;
  p.out_h = dst->h;
  p.out_w = dst->w;
  p.nsrc = n_src;
  p.act = act;
  if (dst_up2x) {
    p.out2 = static_cast<__half*>(dst_up2x->ptr);
    p.out2_ld = dst_up2x->c_stride;
  }
  return launch_gemm<kModeTma2d>(p, n_tiles, total_kb, static_cast<cudaStream_t>(stream));
}

extern "C" int32_t mafb200_conv3x3s2(const maf_tensor* src, const void* w_packed, const float* bias, int32_t act,
                                     const maf_tensor* dst, void* stream) {
  if (!valid_f16_view(src) || !valid_f16_view(dst)) return fail(MAF_E_ARG, "conv3x3s2: bad src/dst tensor");
  if (!aligned_f16_view(src) || !aligned_f16_view(dst)) return fail(MAF_E_ALIGN, "conv3x3s2: alignment");
  if (!w_packed || !bias) return fail(MAF_E_ARG, "conv3x3s2: null weights/bias");
  if ((reinterpret_cast<uintptr_t>(w_packed) & 15) != 0) return fail(MAF_E_ALIGN, "conv3x3s2: w_packed alignment");
  if ((src->h & 1) || (src->w & 1) || dst->h != src->h / 2 || dst->w != src->w / 2 || dst->n != src->n)
    return fail(MAF_E_ARG, "conv3x3s2: need even h,w and dst = [n,h/2,w/2,cout] (src %dx%d dst %dx%d)", src->h, src->w,
                dst->h, dst->w);
  if (act < MAF_ACT_NONE || act > MAF_ACT_SIGMOID) return fail(MAF_E_ARG, "conv3x3s2: bad act %d", act);
  int32_t rc = require_sm100();
  if (rc) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  rc = encode_a_map_im2col(&p.tmA[0], src);
  if (rc) return rc;
  int n_tiles = 0, tile_n = 0;
  mafb200_gemm_tiling(dst->c, &n_tiles, &tile_n);
  const int k_packed = mafb200_packed_k_3x3(src->c);
  rc = encode_w_map(&p.tmW, w_packed, k_packed, n_tiles * tile_n, tile_n);
  if (rc) return rc;
  p.kblocks[0] = ceil_div(src->c, kBlockK);
  p.nsrc = 1;
  p.bias = bias;
  p.out = static_cast<__half*>(dst->ptr);
  p.out_ld = dst->c_stride;
  p.M = dst->n * dst->h * dst->w;
  p.N = dst->c;
  p.tile_n = tile_n;
  p.out_h = dst->h;
  p.out_w = dst->w;
  p.act = act;
  return launch_gemm<kModeIm2colTma>(p, n_tiles, 9 * p.kblocks[0], static_cast<cudaStream_t>(stream));
}

// 3x3 s2 pad 1 over pixel pairs: requires src->c_stride == 32 (so a pair is one 128-byte swizzle row), even w,
// and FINITE values in the padding channels [c, c_stride) of src (they meet zero weights).  Packed weights: fp16
// [rows][6 * 64], block o = (ky, po): po = 0 -> kx = 0 at channel offset c_stride (second pixel of the left pair);
// po = 1 -> kx = 1 at offset 0 and kx = 2 at offset c_stride.
extern "C" int32_t mafb200_conv3x3s2_pair(const maf_tensor* src, const void* w_packed, const float* bias, int32_t act,
                                          const maf_tensor* dst, void* stream) {
  if (!valid_f16_view(src) || !valid_f16_view(dst)) return fail(MAF_E_ARG, "conv3x3s2_pair: bad src/dst tensor");
  if (!aligned_f16_view(src) || !aligned_f16_view(dst)) return fail(MAF_E_ALIGN, "conv3x3s2_pair: alignment");
  if (!w_packed || !bias) return fail(MAF_E_ARG, "conv3x3s2_pair: null weights/bias");
  if ((reinterpret_cast<uintptr_t>(w_packed) & 15) != 0) return fail(MAF_E_ALIGN, "conv3x3s2_pair: w_packed alignment");
  if (2 * src->c_stride != kBlockK)
    return fail(MAF_E_ARG, "conv3x3s2_pair: needs c_stride == 32 (a pixel pair = one 128-byte row), got %d", src->c_stride);
  if ((src->h & 1) || (src->w & 1) || dst->h != src->h / 2 || dst->w != src->w / 2 || dst->n != src->n)
    return fail(MAF_E_ARG, "conv3x3s2_pair: need even h,w and dst = [n,h/2,w/2,cout]");
  if (act < MAF_ACT_NONE || act > MAF_ACT_SIGMOID) return fail(MAF_E_ARG, "conv3x3s2_pair: bad act %d", act);
  int32_t rc = require_sm100();
  if (rc) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  rc = encode_a_map_im2col_pair(&p.tmA[0], src);
  if (rc) return rc;
  int n_tiles = 0, tile_n = 0;
  mafb200_gemm_tiling(dst->c, &n_tiles, &tile_n);
  rc = encode_w_map(&p.tmW, w_packed, 6 * kBlockK, n_tiles * tile_n, tile_n);
  if (rc) return rc;
  p.kblocks[0] = 1;
  p.nsrc = 1;
  p.pair = 1;
  p.bias = bias;
  p.out = static_cast<__half*>(dst->ptr);
  p.out_ld = dst->c_stride;
  p.M = dst->n * dst->h * dst->w;
  p.N = dst->c;
  p.tile_n = tile_n;
  p.out_h = dst->h;
  p.out_w = dst->w;
  p.act = act;
  return launch_gemm<kModeIm2colTma>(p, n_tiles, 6, static_cast<cudaStream_t>(stream));
}

// ---- K7 host side ---------------------------------------------------------------------------------------------
extern "C" int32_t mafb200_detect_cfg_fill(maf_detect_cfg* cfg, double conf_thres, int32_t multi_label, int32_t nc,
                                           const uint8_t* class_filter_host) {
  if (!cfg || nc < 1 || nc > 256) return fail(MAF_E_ARG, "detect_cfg_fill: bad arguments (nc=%d)", nc);
  if (!(conf_thres >= 0.0 && conf_thres <= 1.0))
    return fail(MAF_E_ARG, "detect_cfg_fill: conf_thres must be in [0,1], got %g", conf_thres);
  memset(cfg, 0, sizeof(*cfg));
  cfg->conf = static_cast<float>(conf_thres);
  // raw-value bound under which sigmoid(z) cannot exceed conf (a safety margin below logit(conf))
  if (conf_thres <= 0.0) cfg->skip_below = -INFINITY;
  else if (conf_thres >= 1.0) cfg->skip_below = 30.0f;
  else cfg->skip_below = static_cast<float>(log(conf_thres / (1.0 - conf_thres)) - 0.05);
  cfg->multi_label = (multi_label != 0 && nc > 1) ? 1 : 0;  // nms.py:57
  if (class_filter_host) {
    cfg->has_filter = 1;
    memcpy(cfg->class_filter, class_filter_host, static_cast<size_t>(nc));
  }
  return MAF_OK;
}

extern "C" int32_t mafb200_detect_reset(void* workspace, int32_t batch, void* stream) {
  if (!workspace || batch < 1) return fail(MAF_E_ARG, "detect_reset: bad arguments");
  cudaError_t e = cudaMemsetAsync(workspace, 0, static_cast<size_t>(batch) * 4, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(MAF_E_CUDA, "detect_reset: cudaMemsetAsync: %s", cudaGetErrorString(e));
  return MAF_OK;
}

extern "C" int32_t mafb200_head_pred(const maf_tensor* src, const void* w_packed, const float* bias, int32_t kind,
                                     int32_t anchor_off, int32_t total_anchors, float stride, int32_t nc, float* pred,
                                     float* boxes, const maf_detect_cfg* detect_cfg, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (!valid_f16_view(src) || !aligned_f16_view(src)) return fail(MAF_E_ARG, "head_pred: bad src tensor");
  if (!w_packed || !bias || (reinterpret_cast<uintptr_t>(w_packed) & 15)) return fail(MAF_E_ARG, "head_pred: bad weights");
  const bool train_form = kind == MAF_HEAD_CLS_TRAIN || kind == MAF_HEAD_REG_TRAIN;
  if (train_form) {
    if (!pred || boxes || detect_cfg) return fail(MAF_E_ARG, "head_pred(train form): pred only");
    kind = kind == MAF_HEAD_CLS_TRAIN ? MAF_HEAD_CLS : MAF_HEAD_REG;
  }
  if (kind != MAF_HEAD_CLS && kind != MAF_HEAD_REG) return fail(MAF_E_ARG, "head_pred: kind %d", kind);
  if (nc < 1 || nc > 128) return fail(MAF_E_ARG, "head_pred: nc=%d (1..128)", nc);
  const long long L = static_cast<long long>(src->h) * src->w;
  if (anchor_off < 0 || anchor_off + L > total_anchors) return fail(MAF_E_ARG, "head_pred: anchor range");
  if (static_cast<long long>(total_anchors) * nc > 0x7fffffffll) return fail(MAF_E_ARG, "head_pred: anchors*nc overflows int32");
  if (kind == MAF_HEAD_REG && !pred && !boxes) return fail(MAF_E_ARG, "head_pred(reg): neither pred nor boxes given");
  if (kind == MAF_HEAD_CLS && !pred && !detect_cfg) return fail(MAF_E_ARG, "head_pred(cls): neither pred nor detect_cfg given");
  if (boxes && (reinterpret_cast<uintptr_t>(boxes) & 15)) return fail(MAF_E_ALIGN, "head_pred: boxes must be 16-B aligned");
  int32_t rc = require_sm100();
  if (rc) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  rc = encode_a_map_2d(&p.tmA[0], src);
  if (rc) return rc;
  const int cout = kind == MAF_HEAD_CLS ? nc : 68;
  int n_tiles = 0, tile_n = 0;
  mafb200_gemm_tiling(cout, &n_tiles, &tile_n);  // nc <= 128 and 68 -> one column tile
  const int chans[1] = {src->c};
  const int k_packed = mafb200_packed_k_1x1(chans, 1);
  rc = encode_w_map(&p.tmW, w_packed, k_packed, tile_n, tile_n);
  if (rc) return rc;
  p.kblocks[0] = ceil_div(src->c, kBlockK);
  p.nsrc = 1;
  p.bias = bias;
  p.M = src->n * src->h * src->w;
  p.N = cout;
  p.tile_n = tile_n;
  p.out_h = src->h;
  p.out_w = src->w;
  p.pred = pred;
  p.boxes = kind == MAF_HEAD_REG ? boxes : nullptr;
  p.lvl_L = static_cast<int32_t>(L);
  p.lvl_w = src->w;
  p.anchor_off = anchor_off;
  p.total_anchors = total_anchors;
  p.no = train_form ? nc : 5 + nc;
  p.cls_off = train_form ? 0 : 5;
  p.raw_reg = train_form && kind == MAF_HEAD_REG;
  p.nc = nc;
  p.lvl_stride = stride;
  if (kind == MAF_HEAD_CLS && detect_cfg) {
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return fail(MAF_E_ALIGN, "head_pred: workspace must be 256-B aligned");
    if (workspace_bytes < mafb200_nms_workspace_bytes(src->n, total_anchors, nc))
      return fail(MAF_E_WORKSPACE, "head_pred: workspace %zu < required %zu", workspace_bytes,
                  mafb200_nms_workspace_bytes(src->n, total_anchors, nc));
    const size_t hdr = ((static_cast<size_t>(src->n) * 4 + 255) / 256) * 256;
    p.cfg = detect_cfg;
    p.ncand = static_cast<int32_t*>(workspace);
    p.keys = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) + hdr);
    long long cap = 1;
    while (cap < static_cast<long long>(total_anchors) * nc) cap <<= 1;
    p.cap_pow2 = cap;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return kind == MAF_HEAD_REG ? launch_gemm<kModeTma2d, 1>(p, 1, p.kblocks[0], st)
                              : launch_gemm<kModeTma2d, 2>(p, 1, p.kblocks[0], st);
}
