// Can LSU gathers (cp.async 16 B, hand-swizzled into the SWIZZLE_128B K-major layout) feed a 3x3 stride-2 implicit GEMM
// faster than the im2col TMA (~0.35-0.55 us per 128-pixel box per SM)?  P producer warps fill a ring of 16 KB slots
// (one slot = one tap of one 128-pixel tile, C channels), a consumer thread releases them.  Reports ns per slot per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../maf_yolo_b200/csrc/common.cuh"
using namespace mafb200;
template <int PW>
__global__ void __launch_bounds__(32 * (PW + 1)) k(const __half* in, int n, int H, int W, int C, int ld, int stages, int m_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * 16384);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], PW); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2, nch = C / 8;  // 16-byte chunks per pixel
  if (warp < PW) {
    // producer warp `warp` owns rows [warp*128/PW, (warp+1)*128/PW) of every slot
    const int rows_per = 128 / PW;
    constexpr int LAG = 2;
    int kb = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      for (int tap = 0; tap < 9; ++tap, ++kb) {
        const int slot = kb % stages; const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&empty[slot], ph ^ 1);
        const int ky = tap / 3, kx = tap - ky * 3;
        const uint32_t sbase = smem_u32(smem + slot * 16384);
        for (int idx = lane; idx < rows_per * nch; idx += 32) {
          const int rl = idx / nch, ch = idx - rl * nch;
          const int r = warp * rows_per + rl;
          const long long m = (long long)mt * 128 + r;
          const int ox = (int)(m % Wo), oy = (int)((m / Wo) % Ho), img = (int)(m / ((long long)Wo * Ho));
          const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
          const bool ok = img < n && iy >= 0 && iy < H && ix >= 0 && ix < W;
          const __half* src = ok ? in + (((size_t)img * H + iy) * W + ix) * ld + ch * 8 : in;
          const uint32_t dst = sbase + r * 128 + ((ch ^ (r & 7)) << 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
        // completion, writer side: commit this slot's copies as a group; signal the slot issued LAG slots ago once its
        // group has landed (the copies of the newer slots stay in flight), with the proxy fence the tensor core needs
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (kb >= LAG) {
          asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[(kb - LAG) % stages]);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
      for (int j = (kb >= LAG ? kb - LAG : 0); j < kb; ++j) mbar_arrive(&full[j % stages]);
  } else if (lane == 0) {
    int kb = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x)
      for (int tap = 0; tap < 9; ++tap, ++kb) {
        const int slot = kb % stages; const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&full[slot], ph);
        mbar_arrive(&empty[slot]);
      }
  }
}
template <int PW> void run(const __half* d, int n, int H, int W, int C, int ld, int stages, int cps, int sms) {
  const int m_tiles = n * (H / 2) * (W / 2) / 128;
  const size_t smem = (size_t)stages * 16384 + 2 * stages * 8 + 1024;
  cudaFuncSetAttribute(k<PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PW><<<sms * cps, 32 * (PW + 1), smem>>>(d, n, H, W, C, ld, stages, m_tiles); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<PW><<<sms * cps, 32 * (PW + 1), smem>>>(d, n, H, W, C, ld, stages, m_tiles); cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize(); float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double slots = (double)m_tiles * 9;
  printf("C=%3d %dx%d n=%d  producer warps=%d stages=%d ctas/SM=%d : %7.1f us  %6.1f ns/slot/SM  (%s)\n", C, H, W, n, PW, stages, cps,
         ms * 1e3, ms * 1e6 / (slots / sms), cudaGetErrorString(err));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  __half* d; cudaMalloc(&d, 1ll << 30); cudaMemset(d, 0, 1ll << 30);
  // (C, H=W of the input, ld): L3/L18 48@160, L5 96@80 (first 64-ch block), L23 128@80 (one block), L1 24@320
  run<1>(d, 32, 160, 160, 48, 48, 6, 2, sms); run<2>(d, 32, 160, 160, 48, 48, 6, 2, sms); run<4>(d, 32, 160, 160, 48, 48, 6, 2, sms);
  run<4>(d, 32, 160, 160, 48, 48, 3, 2, sms); run<4>(d, 32, 160, 160, 48, 48, 6, 1, sms);
  run<2>(d, 32, 80, 80, 64, 128, 6, 2, sms); run<4>(d, 32, 80, 80, 64, 128, 6, 2, sms);
  run<4>(d, 32, 320, 320, 24, 32, 6, 2, sms); run<2>(d, 32, 320, 320, 24, 32, 6, 2, sms);
  return 0;
}
