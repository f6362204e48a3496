// Micro-benchmark: SFU exponential throughput on sm_100a (B200), fp32 vs packed f16x2, vs an FMA-pipe polynomial.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bench_sfu tools/bench_sfu.cu ; run: tools/bench_sfu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// 2^x for x <= 0 on the FMA/ALU pipes: round-to-nearest split + degree-4 polynomial on [-0.5, 0.5] + exponent add
__device__ __forceinline__ float ex2poly(float x) {
  x = fmaxf(x, -120.f);
  const float t = x + 12582912.f;              // 1.5 * 2^23: integer part lands in the low mantissa bits
  const float n = t - 12582912.f;
  const float f = x - n;
  float p = fmaf(f, 0.0096181291f, 0.0555041087f);
  p = fmaf(p, f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB800B800u + threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.0f;
      else if (MODE == 1) h[i] = ex2h2(h[i]) ^ 0x80008000u;
      else a[i] = ex2poly(a[i]) - 1.0f;
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, double vals_per_op) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* d; cudaMalloc(&d, 4);
  const int iters = 4096, threads = 512, blocks = sms * 4;
  k<MODE><<<blocks, threads>>>(d, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * threads * iters * 8;
  printf("%-28s %8.3f ms  %7.2f G values/s  %6.2f values/clk/SM (at %d MHz nominal)\n", name, ms, ops * vals_per_op / ms / 1e6,
         ops * vals_per_op / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(d);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("fma-pipe polynomial exp2", 1);
  return 0;
}
