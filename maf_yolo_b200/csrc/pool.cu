// Pooling / resampling / layout kernels (all pure bandwidth; NHWC fp16, 8 channels = one 16-B
// vector per thread, lanes walk consecutive channel groups -> coalesced).
//
//   maxpool2x2      MP (yolov6/layers/common.py:667-673) inside MPRep (common.py:787-792)
//   sppf_pool       the three chained MaxPool2d(5,1,2) of SPPF (common.py:121-129) in ONE pass:
//                   chained 5x5 stride-1 max-pools with -inf padding equal 5x5 / 9x9 / 13x13 windows
//   upsample2x      nn.Upsample(None, 2, 'nearest') (configs/yaml/MAF-YOLO-n.yaml:21,26)
//   nchw<->nhwc     boundary converters for the block-level nn.Module drop-ins
#include "common.cuh"
#include "host.h"

namespace mafb200 {

__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  *reinterpret_cast<__half2*>(&r.x) = __hmax2(*reinterpret_cast<__half2*>(&a.x), *reinterpret_cast<__half2*>(&b.x));
  *reinterpret_cast<__half2*>(&r.y) = __hmax2(*reinterpret_cast<__half2*>(&a.y), *reinterpret_cast<__half2*>(&b.y));
  *reinterpret_cast<__half2*>(&r.z) = __hmax2(*reinterpret_cast<__half2*>(&a.z), *reinterpret_cast<__half2*>(&b.z));
  *reinterpret_cast<__half2*>(&r.w) = __hmax2(*reinterpret_cast<__half2*>(&a.w), *reinterpret_cast<__half2*>(&b.w));
  return r;
}

// ---- 2x2 stride-2 max pool -------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    maxpool2x2_kernel(const __half* __restrict__ in, int in_ld, __half* __restrict__ out, int out_ld, int B, int H,
                      int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int cg = C >> 3;
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * cg;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int g = static_cast<int>(t % cg);
  long long px = t / cg;
  const int ox = static_cast<int>(px % Wo);
  px /= Wo;
  const int oy = static_cast<int>(px % Ho);
  const int b = static_cast<int>(px / Ho);
  const __half* p00 = in + ((static_cast<size_t>(b) * H + 2 * oy) * W + 2 * ox) * in_ld + g * 8;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p00));
  const uint4 c = __ldg(reinterpret_cast<const uint4*>(p00 + in_ld));
  const uint4 d = __ldg(reinterpret_cast<const uint4*>(p00 + static_cast<size_t>(W) * in_ld));
  const uint4 e = __ldg(reinterpret_cast<const uint4*>(p00 + static_cast<size_t>(W) * in_ld + in_ld));
  const uint4 r = hmax8(hmax8(a, c), hmax8(d, e));
  *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * Ho + oy) * Wo + ox) * out_ld + g * 8) = r;
}

// ---- SPPF: 5x5 / 9x9 / 13x13 stride-1 max windows in one pass ----------------------------------
// One CTA = one image x one 8-channel group; the whole map (<= 40x40 here: 20x20 at 640 input)
// lives in smem.  Separable: row-max for radius 2/4/6, then column-max.
__global__ void __launch_bounds__(256)
    sppf_pool_kernel(const __half* __restrict__ in, int in_ld, __half* __restrict__ y1, int ld1,
                     __half* __restrict__ y2, int ld2, __half* __restrict__ y3, int ld3, int H, int W, int C) {
  extern __shared__ uint4 s_map[];  // [4][H*W]: input, row-max r2, r4, r6
  pdl_launch_dependents();
  pdl_wait();
  const int cg = C >> 3;
  const int g = blockIdx.x % cg;
  const int b = blockIdx.x / cg;
  const int hw = H * W;
  uint4* s_in = s_map;
  uint4* s_r2 = s_map + hw;
  uint4* s_r4 = s_map + 2 * hw;
  uint4* s_r6 = s_map + 3 * hw;
  const size_t img = static_cast<size_t>(b) * hw;
  for (int i = threadIdx.x; i < hw; i += blockDim.x)
    s_in[i] = __ldg(reinterpret_cast<const uint4*>(in + (img + i) * in_ld + g * 8));
  __syncthreads();
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    uint4 m = s_in[i];
#pragma unroll
    for (int d = 1; d <= 6; ++d) {
      if (x - d >= 0) m = hmax8(m, s_in[i - d]);
      if (x + d < W) m = hmax8(m, s_in[i + d]);
      if (d == 2) s_r2[i] = m;
      if (d == 4) s_r4[i] = m;
    }
    s_r6[i] = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const int y = i / W;
    uint4 m2 = s_r2[i], m4 = s_r4[i], m6 = s_r6[i];
#pragma unroll
    for (int d = 1; d <= 6; ++d) {
      const bool up = y - d >= 0, dn = y + d < H;
      if (d <= 2) {
        if (up) m2 = hmax8(m2, s_r2[i - d * W]);
        if (dn) m2 = hmax8(m2, s_r2[i + d * W]);
      }
      if (d <= 4) {
        if (up) m4 = hmax8(m4, s_r4[i - d * W]);
        if (dn) m4 = hmax8(m4, s_r4[i + d * W]);
      }
      if (up) m6 = hmax8(m6, s_r6[i - d * W]);
      if (dn) m6 = hmax8(m6, s_r6[i + d * W]);
    }
    *reinterpret_cast<uint4*>(y1 + (img + i) * ld1 + g * 8) = m2;
    *reinterpret_cast<uint4*>(y2 + (img + i) * ld2 + g * 8) = m4;
    *reinterpret_cast<uint4*>(y3 + (img + i) * ld3 + g * 8) = m6;
  }
}

// ---- nearest x2 upsample -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    upsample2x_kernel(const __half* __restrict__ in, int in_ld, __half* __restrict__ out, int out_ld, int B, int H,
                      int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int cg = C >> 3;
  const long long total = static_cast<long long>(B) * H * W * cg;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int g = static_cast<int>(t % cg);
  long long px = t / cg;
  const int x = static_cast<int>(px % W);
  px /= W;
  const int y = static_cast<int>(px % H);
  const int b = static_cast<int>(px / H);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(b) * H + y) * W + x) * in_ld + g * 8));
  __half* o = out + ((static_cast<size_t>(b) * 2 * H + 2 * y) * (2 * W) + 2 * x) * out_ld + g * 8;
  const size_t dy = static_cast<size_t>(2) * W * out_ld;
  *reinterpret_cast<uint4*>(o) = v;
  *reinterpret_cast<uint4*>(o + out_ld) = v;
  *reinterpret_cast<uint4*>(o + dy) = v;
  *reinterpret_cast<uint4*>(o + dy + out_ld) = v;
}

// ---- NCHW (f32/f16) -> NHWC fp16 via a 32x32 smem transpose over (c, hw) ------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    nchw_to_nhwc_kernel(const T* __restrict__ in, __half* __restrict__ out, int out_ld, int C, int HW) {
  __shared__ float tile[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + tx;
    float v = 0.f;
    if (c < C && p < HW) v = static_cast<float>(in[(static_cast<size_t>(b) * C + c) * HW + p]);
    tile[j][tx] = v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + tx;
    if (c < C && p < HW) out[(static_cast<size_t>(b) * HW + p) * out_ld + c] = __float2half_rn(tile[tx][j]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    nhwc_to_nchw_kernel(const __half* __restrict__ in, int in_ld, T* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + tx;
    float v = 0.f;
    if (c < C && p < HW) v = __half2float(in[(static_cast<size_t>(b) * HW + p) * in_ld + c]);
    tile[j][tx] = v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + tx;
    if (c < C && p < HW) out[(static_cast<size_t>(b) * C + c) * HW + p] = static_cast<T>(tile[tx][j]);
  }
}

static bool vec8_ok(const maf_tensor* t) { return aligned_f16_view(t) && (t->c % 8) == 0; }

}  // namespace mafb200

using namespace mafb200;

extern "C" int32_t mafb200_maxpool2x2(const maf_tensor* src, const maf_tensor* dst, void* stream) {
  if (!valid_f16_view(src) || !valid_f16_view(dst)) return fail(MAF_E_ARG, "maxpool2x2: bad src/dst");
  if ((src->h & 1) || (src->w & 1) || dst->n != src->n || dst->h != src->h / 2 || dst->w != src->w / 2 ||
      dst->c != src->c)
    return fail(MAF_E_ARG, "maxpool2x2: need even h,w and dst [n,h/2,w/2,c]");
  if (!vec8_ok(src) || !vec8_ok(dst)) return fail(MAF_E_ALIGN, "maxpool2x2: c %% 8, c_stride %% 8, 16-B pointers");
  int32_t rc = require_sm100();
  if (rc) return rc;
  const long long total = static_cast<long long>(dst->n) * dst->h * dst->w * (dst->c / 8);
  launch_pdl(maxpool2x2_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
             static_cast<cudaStream_t>(stream),
      static_cast<const __half*>(src->ptr), src->c_stride, static_cast<__half*>(dst->ptr), dst->c_stride, src->n,
      src->h, src->w, src->c);
  return check_launch("maxpool2x2 kernel launch");
}

extern "C" int32_t mafb200_sppf_pool(const maf_tensor* src, const maf_tensor* y1, const maf_tensor* y2,
                                     const maf_tensor* y3, void* stream) {
  const maf_tensor* ys[3] = {y1, y2, y3};
  if (!valid_f16_view(src)) return fail(MAF_E_ARG, "sppf_pool: bad src");
  if (!vec8_ok(src)) return fail(MAF_E_ALIGN, "sppf_pool: src alignment");
  for (int i = 0; i < 3; ++i) {
    if (!valid_f16_view(ys[i]) || !same_nhw(ys[i], src) || ys[i]->c != src->c)
      return fail(MAF_E_ARG, "sppf_pool: y%d shape mismatch", i + 1);
    if (!vec8_ok(ys[i])) return fail(MAF_E_ALIGN, "sppf_pool: y%d alignment", i + 1);
  }
  const size_t smem = static_cast<size_t>(4) * src->h * src->w * sizeof(uint4);
  if (smem > 200 * 1024) return fail(MAF_E_ARG, "sppf_pool: map %dx%d too large for the one-CTA-per-map kernel", src->h, src->w);
  int32_t rc = require_sm100();
  if (rc) return rc;
  {
    static SmemOptIn opt_in;  // per device (a process-wide flag skipped the opt-in on a second GPU)
    const int32_t rc_attr = smem_opt_in(opt_in, sppf_pool_kernel, 200 * 1024, "sppf_pool");
    if (rc_attr) return rc_attr;
  }
  const unsigned blocks = static_cast<unsigned>(src->n) * (src->c / 8);
  launch_pdl(sppf_pool_kernel, dim3(blocks), dim3(256), smem, static_cast<cudaStream_t>(stream),
      static_cast<const __half*>(src->ptr), src->c_stride, static_cast<__half*>(y1->ptr), y1->c_stride,
      static_cast<__half*>(y2->ptr), y2->c_stride, static_cast<__half*>(y3->ptr), y3->c_stride, src->h, src->w,
      src->c);
  return check_launch("sppf_pool kernel launch");
}

extern "C" int32_t mafb200_upsample2x(const maf_tensor* src, const maf_tensor* dst, void* stream) {
  if (!valid_f16_view(src) || !valid_f16_view(dst)) return fail(MAF_E_ARG, "upsample2x: bad src/dst");
  if (dst->n != src->n || dst->h != 2 * src->h || dst->w != 2 * src->w || dst->c != src->c)
    return fail(MAF_E_ARG, "upsample2x: dst must be [n,2h,2w,c]");
  if (!vec8_ok(src) || !vec8_ok(dst)) return fail(MAF_E_ALIGN, "upsample2x: alignment");
  int32_t rc = require_sm100();
  if (rc) return rc;
  const long long total = static_cast<long long>(src->n) * src->h * src->w * (src->c / 8);
  launch_pdl(upsample2x_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
             static_cast<cudaStream_t>(stream),
      static_cast<const __half*>(src->ptr), src->c_stride, static_cast<__half*>(dst->ptr), dst->c_stride, src->n,
      src->h, src->w, src->c);
  return check_launch("upsample2x kernel launch");
}

extern "C" int32_t mafb200_nchw_to_nhwc_f16(const void* src, int32_t src_dtype, const maf_tensor* dst, void* stream) {
  if (!src || !valid_f16_view(dst)) return fail(MAF_E_ARG, "nchw_to_nhwc: bad arguments");
  int32_t rc = require_sm100();
  if (rc) return rc;
  const int HW = dst->h * dst->w;
  dim3 grid(ceil_div(HW, 32), ceil_div(dst->c, 32), dst->n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_dtype == MAF_F32)
    launch_pdl(nchw_to_nhwc_kernel<float>, grid, dim3(256), 0, st, static_cast<const float*>(src), static_cast<__half*>(dst->ptr),
                                                     dst->c_stride, dst->c, HW);
  else if (src_dtype == MAF_F16)
    launch_pdl(nchw_to_nhwc_kernel<__half>, grid, dim3(256), 0, st, static_cast<const __half*>(src), static_cast<__half*>(dst->ptr),
                                                      dst->c_stride, dst->c, HW);
  else
    return fail(MAF_E_ARG, "nchw_to_nhwc: unsupported dtype %d", src_dtype);
  return check_launch("nchw_to_nhwc kernel launch");
}

extern "C" int32_t mafb200_nhwc_f16_to_nchw(const maf_tensor* src, void* dst, int32_t dst_dtype, void* stream) {
  if (!dst || !valid_f16_view(src)) return fail(MAF_E_ARG, "nhwc_to_nchw: bad arguments");
  int32_t rc = require_sm100();
  if (rc) return rc;
  const int HW = src->h * src->w;
  dim3 grid(ceil_div(HW, 32), ceil_div(src->c, 32), src->n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dst_dtype == MAF_F32)
    launch_pdl(nhwc_to_nchw_kernel<float>, grid, dim3(256), 0, st, static_cast<const __half*>(src->ptr), src->c_stride,
                                                     static_cast<float*>(dst), src->c, HW);
  else if (dst_dtype == MAF_F16)
    launch_pdl(nhwc_to_nchw_kernel<__half>, grid, dim3(256), 0, st, static_cast<const __half*>(src->ptr), src->c_stride,
                                                      static_cast<__half*>(dst), src->c, HW);
  else
    return fail(MAF_E_ARG, "nhwc_to_nchw: unsupported dtype %d", dst_dtype);
  return check_launch("nhwc_to_nchw kernel launch");
}
