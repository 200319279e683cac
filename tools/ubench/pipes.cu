// Micro-benchmarks of per-SM instruction rates on B200 (used to size the epilogue / depth-wise kernels).
// Each kernel runs ITER iterations of 8 independent chains per thread; 148*4 CTAs x 256 threads.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <mma.h>
constexpr int ITER = 4096;
#define CHAINS 8
template <int OP>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
  float v[CHAINS];
  uint32_t u[CHAINS];
  unsigned long long q[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { v[i] = seed + i * 0.01f + threadIdx.x * 1e-4f; u[i] = __float_as_uint(v[i]); q[i] = (unsigned long long)u[i] << 32 | u[i]; }
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(v[i]) : "f"(seed));
      if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(q[i]) : "l"(q[(i + 1) % CHAINS]));
      if (OP == 5) asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
      if (OP == 6) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (OP == 7) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(v[i]), "f"(__uint_as_float(u[(i + 1) % CHAINS])));
      if (OP == 8) asm volatile("{.reg .f16 l,h; mov.b32 {l,h}, %1; cvt.f32.f16 %0, l;}" : "=f"(v[i]) : "r"(__float_as_uint(v[i])));
      if (OP == 9) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(q[(i + 1) % CHAINS]));
      if (OP == 10) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += v[i] + __uint_as_float(u[i]) + __uint_as_float((uint32_t)q[i]);
  if (s == 123.456f) out[0] = s;
}
// mma.sync m16n8k16 f16 -> f32 rate
__global__ void __launch_bounds__(256) kmma(float* out) {
  uint32_t a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, 6};
  float c[4][4] = {};
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
}
template <int OP> void run(const char* name, float* d, int sms, double per_thread_ops) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<sms * 4, 256>>>(d, 0.3f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<OP><<<sms * 4, 256>>>(d, 0.3f); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)sms * 4 * 256 * ITER * CHAINS * per_thread_ops;
  printf("%-28s %8.3f ms  %8.2f Gop/s/SM (thread-level results)  = %6.2f /clk/SM @1.965GHz\n", name, ms, ops / ms / 1e6 / sms, ops / ms / 1e6 / sms / 1.965);
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* d; cudaMalloc(&d, 4);
  run<0>("tanh.approx.f32", d, sms, 1); run<1>("ex2.approx.f32", d, sms, 1); run<2>("rcp.approx.f32", d, sms, 1);
  run<3>("fma.f32", d, sms, 1); run<4>("fma.f32x2 (2 fma)", d, sms, 2); run<5>("fma.f16x2 (2 fma)", d, sms, 2);
  run<6>("tanh.f16x2 (2 res)", d, sms, 2); run<7>("cvt.f16x2.f32 (pack)", d, sms, 1); run<8>("cvt.f32.f16", d, sms, 1);
  run<9>("add.f32x2 (2 add)", d, sms, 2); run<10>("ex2.f16x2 (2 res)", d, sms, 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kmma<<<sms * 4, 256>>>(d); cudaDeviceSynchronize();
  cudaEventRecord(e0); kmma<<<sms * 4, 256>>>(d); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = (double)sms * 4 * 8 * ITER * 4 * 2.0 * 16 * 8 * 16;
  printf("mma.sync m16n8k16 f16/f32     %8.3f ms  %8.1f TFLOP/s dense\n", ms, fl / ms / 1e9);
  return 0;
}
