// Row-per-thread output through cp.async.bulk (shared -> global, 1-D, one request per output row) versus
// st.global.v8 (tools/ubench/store_rate.cu).  Each CTA (256 threads) stages 256 rows of `row_bytes` in padded smem
// and every thread bulk-copies its own row; rows are contiguous in global memory (row pitch = row_bytes).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
template <int MODE>
__global__ void __launch_bounds__(256) k(uint8_t* out, long long M, int row_bytes) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int pitch = row_bytes + 16;
  const long long nblk = (M + 255) / 256;
  uint8_t* my = sm + threadIdx.x * pitch;
  for (long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const long long row = b * 256 + threadIdx.x;
    // "epilogue": write the row into smem (16-byte stores, conflict-free thanks to the padding)
    for (int c = 0; c < row_bytes; c += 16) *reinterpret_cast<uint4*>(my + c) = make_uint4(threadIdx.x, c, b, 7);
    if (MODE == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (row < M)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + row * row_bytes), "r"(smem_u32(my)), "r"(row_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging row reusable
    } else {
      if (row < M)
        for (int c = 0; c < row_bytes; c += 32) {
          const uint4 a = *reinterpret_cast<uint4*>(my + c), d = *reinterpret_cast<uint4*>(my + c + 16);
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + row * row_bytes + c), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w) : "memory");
        }
    }
  }
}
template <int MODE> void run(const char* name, uint8_t* d, long long M, int rb, int sms, int cps) {
  const size_t smem = 256 * (rb + 16);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * cps, 256, smem>>>(d, M, rb); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < 5; ++i) k<MODE><<<sms * cps, 256, smem>>>(d, M, rb); cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize(); float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("%-28s M=%lld row=%4d B ctas/SM=%d : %7.1f us  %7.1f GB/s (%s)\n", name, M, rb, cps, ms * 1e3, (double)M * rb / ms / 1e6, cudaGetErrorString(err));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* d; cudaMalloc(&d, 1ll << 30);
  for (int rb : {384, 256, 160, 128, 64})
    for (long long M : {204800ll, 819200ll})
      for (int cps : {1, 2}) {
        if ((size_t)256 * (rb + 16) * cps > 220 * 1024) continue;
        run<0>("bulk S2G per row", d, M, rb, sms, cps);
        run<1>("smem -> st.global.v8 per row", d, M, rb, sms, cps);
      }
  return 0;
}
