// How fast can one SM pull [128 rows x inner bytes] boxes with cp.async.bulk.tensor.2d?  (B200, sm_100a)
// Grid = SMs x ctas_per_sm persistent CTAs, each with a ring of `stages` 16 KB slots; thread 0 produces,
// thread 32 consumes (waits full, arrives empty).  Reports GB/s of useful bytes and ns per box per SM.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../maf_yolo_b200/csrc/common.cuh"
using namespace mafb200;
__global__ void __launch_bounds__(64) k(const __grid_constant__ CUtensorMap tm, int tiles, int stages, int boxes_per_tile,
                                        int box_bytes, int swz_rows) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * 16384);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int kb = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x)
      for (int j = 0; j < boxes_per_tile; ++j, ++kb) {
        const int slot = kb % stages; const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&empty[slot], ph ^ 1);
        mbar_arrive_expect_tx(&full[slot], box_bytes);
        tma_load_2d(smem + slot * 16384, &tm, &full[slot], j * 64, t * swz_rows);
      }
  } else if (threadIdx.x == 32) {
    int kb = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x)
      for (int j = 0; j < boxes_per_tile; ++j, ++kb) {
        const int slot = kb % stages; const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&full[slot], ph);
        mbar_arrive(&empty[slot]);
      }
  }
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  Enc enc; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const long long M = 819200;
  __half* buf; cudaMalloc(&buf, M * 256 * 2); cudaMemset(buf, 0, M * 256 * 2);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  // cases: channels c (tensor inner dim), c_stride, box rows, stages, ctas/SM
  struct C { int c, ld, rows, stages, cps; CUtensorMapL2promotion l2; };
  C cases[] = {{64, 64, 128, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {64, 64, 128, 3, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
               {64, 64, 128, 6, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {64, 64, 128, 12, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
               {64, 80, 128, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {64, 80, 128, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_NONE},
               {64, 80, 128, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
               {24, 80, 128, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {24, 80, 128, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_NONE},
               {8, 80, 128, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {72, 80, 128, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
               {128, 128, 128, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {192, 192, 128, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
               {64, 64, 256, 3, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {64, 64, 64, 6, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}};
  for (auto& c : cases) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)c.c, (cuuint64_t)M}; cuuint64_t strides[1] = {(cuuint64_t)c.ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)c.rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, c.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode failed %d\n", (int)r); continue; }
    const int bpt = (c.c + 63) / 64, tiles = M / c.rows, box_bytes = c.rows * 128;
    const size_t smem = (size_t)c.stages * (c.rows > 128 ? 32768 : 16384) + 2 * c.stages * 8 + 1024;
    if (c.rows > 128) continue;  // slot is 16 KB
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<sms * c.cps, 64, smem>>>(tm, tiles, c.stages, bpt, box_bytes, c.rows); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<sms * c.cps, 64, smem>>>(tm, tiles, c.stages, bpt, box_bytes, c.rows); cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize(); float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double useful = (double)M * c.c * 2, boxes = (double)tiles * bpt;
    printf("c=%3d ld=%3d rows=%3d stages=%2d ctas/SM=%d l2promo=%d : %7.1f us  %7.1f GB/s useful  %6.1f ns/box/SM  %5.2f ns/row/SM  (%s)\n", c.c, c.ld, c.rows, c.stages, c.cps,
           (int)c.l2, ms * 1e3, useful / ms / 1e6, ms * 1e6 / (boxes / sms), ms * 1e6 / (boxes * c.rows / sms), cudaGetErrorString(err));
  }
  return 0;
}
