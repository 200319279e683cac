// Host-side plumbing: thread-local error text, launch counter, device check, driver entry points.
#include "host.h"

#include <stdlib.h>

#include <atomic>
#include <mutex>

namespace mafb200 {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int32_t fail(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

const char* last_error_text() { return g_err; }

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("MAFB200_PDL");
    return !(v && v[0] == '0');
  }();
  return on;
}

int32_t check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MAF_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  count_launch();
  return MAF_OK;
}

int32_t require_sm100() {
  static std::mutex mu;
  static int cached[64];
  static bool init = false;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(MAF_E_ARCH, "no CUDA device: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(mu);
  if (!init) {
    for (int i = 0; i < 64; ++i) cached[i] = -1;
    init = true;
  }
  if (dev < 0 || dev >= 64) return fail(MAF_E_ARCH, "device index %d out of range", dev);
  if (cached[dev] < 0) {
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return fail(MAF_E_ARCH, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    cached[dev] = major;
  }
  if (cached[dev] != 10)
    return fail(MAF_E_ARCH, "device %d has compute capability %d.x; libmafb200 is sm_100a only (no fallback)", dev,
                cached[dev]);
  return MAF_OK;
}

static void* driver_entry(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
  return fn;
}

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_entry("cuTensorMapEncodeTiled"));
  return fn;
}
EncodeIm2colFn encode_im2col_fn() {
  static EncodeIm2colFn fn = reinterpret_cast<EncodeIm2colFn>(driver_entry("cuTensorMapEncodeIm2col"));
  return fn;
}

}  // namespace mafb200

extern "C" {

int32_t mafb200_version(void) { return MAFB200_VERSION; }

const char* mafb200_last_error(void) { return mafb200::last_error_text(); }
int64_t mafb200_launch_count(void) { return mafb200::launch_count(); }

int32_t mafb200_device_ok(int32_t device) {
  if (device >= 0) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return mafb200::fail(MAF_E_ARCH, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  }
  return mafb200::require_sm100();
}

int32_t mafb200_gemm_tiling(int32_t cout, int32_t* n_tiles, int32_t* tile_n) {
  if (cout <= 0 || !n_tiles || !tile_n) return mafb200::fail(MAF_E_ARG, "gemm_tiling: bad arguments");
  int nt = (cout + 127) / 128;  // <= 128 columns per tile: 4 CTAs (32 epilogue warps) fit one SM's TMEM
  int tn = ((cout + nt - 1) / nt + 15) / 16 * 16;
  *n_tiles = nt;
  *tile_n = tn;
  return MAF_OK;
}

int32_t mafb200_packed_k_1x1(const int32_t* src_channels, int32_t n_src) {
  if (!src_channels || n_src <= 0 || n_src > MAF_MAX_SRC) return mafb200::fail(MAF_E_ARG, "packed_k_1x1: bad n_src");
  int k = 0;
  for (int i = 0; i < n_src; ++i) k += (src_channels[i] + 63) / 64 * 64;
  return k;
}

int32_t mafb200_packed_k_3x3(int32_t cin) {
  if (cin <= 0) return mafb200::fail(MAF_E_ARG, "packed_k_3x3: bad cin");
  return 9 * ((cin + 63) / 64 * 64);
}

}  // extern "C"
