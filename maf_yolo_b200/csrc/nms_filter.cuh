// Candidate selection of one prediction row, shared by the NMS compaction kernel (nms.cu) and the fused
// decode + compaction kernel (decode.cu).  Semantics of yolov6/utils/nms.py:48,69,75-84 (see nms.cu): the row is a
// candidate if obj > conf and max_c cls > conf; multi_label emits every (anchor, class) with cls * obj > conf, else
// the best class (first maximum).  Keys are (~score_bits << 32) | (anchor * nc + class), emitted unordered.
#pragma once
#include <math.h>
#include <stdint.h>

namespace mafb200 {

// Called by all 32 lanes of a warp with the same arguments; `row` points to 5 + nc floats (shared or global).
__device__ __forceinline__ void nms_filter_row(const float* row, int a, int nc, float conf, int multi_label,
                                               const uint8_t* class_filter, int32_t* ncand_b,
                                               unsigned long long* keys, long long cap, int lane) {
  const float obj = row[4];
  float mx = -INFINITY;  // raw class maximum (nms.py:48)
  for (int c = lane; c < nc; c += 32) mx = fmaxf(mx, row[5 + c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (!(obj > conf) || !(mx > conf)) return;

  if (multi_label) {
    for (int c0 = 0; c0 < nc; c0 += 32) {
      const int c = c0 + lane;
      float s = 0.f;
      bool pass = false;
      if (c < nc) {
        s = __fmul_rn(row[5 + c], obj);
        pass = s > conf && (class_filter == nullptr || class_filter[c] != 0);
      }
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (m == 0) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(ncand_b, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (pass) {
        const long long slot = base + __popc(m & ((1u << lane) - 1));
        if (slot < cap) {
          const unsigned sb = __float_as_uint(s);
          keys[slot] = (static_cast<unsigned long long>(~sb) << 32) |
                       static_cast<unsigned long long>(static_cast<unsigned>(a) * nc + c);
        }
      }
    }
  } else {
    // best class by score, first maximum on ties (torch.max semantics on CPU)
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < nc; c += 32) {
      const float s = __fmul_rn(row[5 + c], obj);
      if (s > best) {
        best = s;
        bi = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    if (lane == 0 && best > conf && bi < nc && (class_filter == nullptr || class_filter[bi] != 0)) {
      const long long slot = atomicAdd(ncand_b, 1);
      if (slot < cap) {
        const unsigned sb = __float_as_uint(best);
        keys[slot] = (static_cast<unsigned long long>(~sb) << 32) |
                     static_cast<unsigned long long>(static_cast<unsigned>(a) * nc + bi);
      }
    }
  }
}

}  // namespace mafb200
