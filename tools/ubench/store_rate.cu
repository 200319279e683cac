// Pure-store bandwidth on B200 for the two epilogue store patterns (M x N fp16 output, row pitch = N * 2 B):
//  (a) row-per-thread: lane i of a warp writes 32 B chunks of row (base + i)  [what a TMEM-lane epilogue does]
//  (b) coalesced: a warp writes 1 KB contiguous per instruction               [what a smem-staged / TMA store does]
//  (c) row-per-thread but 16-B stores
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void st256(void* p, uint32_t v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(256) k(uint8_t* out, long long M, int row_bytes) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  if (MODE == 0 || MODE == 2) {
    for (long long r0 = warp * 32; r0 < M; r0 += nwarps * 32) {
      uint8_t* row = out + (r0 + lane) * row_bytes;
      if (MODE == 0) for (int c = 0; c < row_bytes; c += 32) st256(row + c, lane);
      else for (int c = 0; c < row_bytes; c += 16) *reinterpret_cast<uint4*>(row + c) = make_uint4(lane, lane, lane, lane);
    }
  } else {
    const long long total = M * row_bytes;
    for (long long o = warp * 1024; o < total; o += nwarps * 1024) st256(out + o + lane * 32, lane);
  }
}
template <int MODE> void run(const char* name, uint8_t* d, long long M, int row_bytes, int sms, int cps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * cps, 256>>>(d, M, row_bytes); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k<MODE><<<sms * cps, 256>>>(d, M, row_bytes); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("%-34s M=%lld row=%4d B ctas/SM=%d : %7.1f us  %7.1f GB/s  (%s)\n", name, M, row_bytes, cps, ms * 1e3, (double)M * row_bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* d; cudaMalloc(&d, 1ll << 30);
  for (int rb : {384, 256, 160, 128}) {
    for (long long M : {204800ll, 819200ll}) {
      for (int cps : {2, 8}) {
        run<0>("row-per-thread 32B stores", d, M, rb, sms, cps);
        run<2>("row-per-thread 16B stores", d, M, rb, sms, cps);
        run<1>("coalesced 1KB/warp", d, M, rb, sms, cps);
      }
    }
  }
  return 0;
}
