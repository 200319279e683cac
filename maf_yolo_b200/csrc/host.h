// Host-side plumbing shared by the C-ABI translation units: error text, argument checks,
// launch counting and the driver entry points for tensor-map encoding.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mafb200.h"

namespace mafb200 {

// Sets the calling thread's last-error text and returns `code` (so call sites read
// `return fail(MAF_E_ARG, "...")`).
int32_t fail(int32_t code, const char* fmt, ...);
void count_launch();

// Checks the sticky launch error right after a kernel launch (no synchronisation).
int32_t check_launch(const char* what);

// Returns MAF_OK if the current device is sm_100-class (cached per device).
int32_t require_sm100();

// cuTensorMapEncodeTiled / cuTensorMapEncodeIm2col resolved through
// cudaGetDriverEntryPoint, so the library does not link libcuda directly.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();
EncodeIm2colFn encode_im2col_fn();

// Programmatic dependent launch: every kernel of this library executes pdl_wait() (griddepcontrol.wait)
// before it touches memory another kernel may have produced, so consecutive launches in a stream (and
// the kernel->kernel edges of a captured graph) may overlap the next kernel's prologue (barrier init,
// TMEM allocation, weight loads) with the previous kernel's tail.  MAFB200_PDL=0 turns it off.
bool pdl_enabled();

template <bool kPdl = true, typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (kPdl && pdl_enabled()) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // error picked up by check_launch()
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies per device (context): remembered per (kernel
// instantiation, device) — each call site owns a zero-initialised `static SmemOptIn`.  Two threads racing on the first
// call both set the same value, which is harmless.
struct SmemOptIn {
  std::atomic<int> bytes[64];
};
template <typename Kernel>
inline int32_t smem_opt_in(SmemOptIn& st, Kernel kernel, int bytes, const char* what) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) return fail(MAF_E_CUDA, "%s: cudaGetDevice: %s", what, cudaGetErrorString(e));
  if (st.bytes[dev].load(std::memory_order_acquire) >= bytes) return MAF_OK;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return fail(MAF_E_CUDA, "%s: cudaFuncSetAttribute(%d B): %s", what, bytes, cudaGetErrorString(e));
  st.bytes[dev].store(bytes, std::memory_order_release);
  return MAF_OK;
}

inline bool valid_f16_view(const maf_tensor* t) {
  return t && t->ptr && t->dtype == MAF_F16 && t->n > 0 && t->h > 0 && t->w > 0 && t->c > 0 && t->c_stride >= t->c;
}
inline bool aligned_f16_view(const maf_tensor* t) {
  return (reinterpret_cast<uintptr_t>(t->ptr) & 15) == 0 && (t->c_stride % 8) == 0;
}
inline bool same_nhw(const maf_tensor* a, const maf_tensor* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w;
}

}  // namespace mafb200
