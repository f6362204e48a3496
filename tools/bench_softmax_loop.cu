// Micro-benchmark of the attention forward's per-row softmax inner loop on sm_100a (no tensor core, no TMEM): what the SM's
// issue slots + XU (MUFU) pipe can sustain for  p = exp2(s * c - m),  row sum,  16-bit pack,  swizzled smem store  with
//   V0  the r02 kernel's instruction mix (FFMA, MUFU.EX2, FADD, F2FP per pair, STS.128 per 8)
//   V1  packed f32x2 arithmetic (FFMA2 for the scale, FADD2 for the row sum)
//   V2  V1 without the row sum (it can come from the tensor core: a ones column appended to V)
//   V3  V2 + every 4th pair through an FMA-pipe polynomial instead of MUFU (packed f32x2 Cody-Waite + degree-3 Horner)
//   V4  V2 + every 2nd pair through the polynomial
//   V5  V2 + 3 of 8 pairs through the polynomial
// at 8 and 16 "softmax warps" per SM.  Output: elements / clk / SM (XU bound = 16).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bench_softmax_loop tools/bench_softmax_loop.cu
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void un2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint32_t packh2(float a, float b) { const __half2 t = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&t); }

// 2^x for a pair, x <= 0, on the FMA / ALU pipes
__device__ __forceinline__ void ex2poly2(uint64_t x, float& o0, float& o1) {
  float x0, x1; un2(x, x0, x1);
  x0 = fmaxf(x0, -125.f); x1 = fmaxf(x1, -125.f);
  x = pk2(x0, x1);
  const uint64_t magic = pk2(12582912.f, 12582912.f), nmagic = pk2(-12582912.f, -12582912.f);
  const uint64_t t = fadd2(x, magic);
  const uint64_t n = fadd2(t, nmagic);
  const uint64_t f = ffma2(n, pk2(-1.f, -1.f), x);
  uint64_t p = ffma2(f, pk2(0.0555041087f, 0.0555041087f), pk2(0.2402265070f, 0.2402265070f));
  p = ffma2(p, f, pk2(0.6931471806f, 0.6931471806f));
  p = ffma2(p, f, pk2(1.f, 1.f));
  float p0, p1, t0, t1; un2(p, p0, p1); un2(t, t0, t1);
  o0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  o1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

template <int V>
__global__ void __launch_bounds__(128) k(const float* __restrict__ in, float* out, int tiles, float c, float m) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* prow = smem + (threadIdx.x / 8) * 1024 + (threadIdx.x % 8) * 128;
  const int r = threadIdx.x;
  float sv[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) sv[i] = in[(threadIdx.x * 64 + i) & 4095];
  float lrun = 0.f;
  const uint64_t c2 = pk2(c, c);
  for (int t = 0; t < tiles; ++t) {
    const float mneg = m + 1e-9f * t;
    const uint64_t mn2 = pk2(-mneg, -mneg);
    float e[64];
    if (V == 0) {
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; ++i) e[i] = ex2f(fmaf(sv[i], c, -mneg));
#pragma unroll
      for (int i = 0; i < 64; ++i) ls[i & 3] += e[i];
      lrun += (ls[0] + ls[1]) + (ls[2] + ls[3]);
    } else {
      uint64_t acc0 = pk2(0.f, 0.f), acc1 = acc0;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const uint64_t x = ffma2(pk2(sv[i], sv[i + 1]), c2, mn2);
        const int pi = i / 2;
        const bool poly = (V == 3 && (pi & 3) == 3) || (V == 4 && (pi & 1) == 1) || (V == 5 && ((pi & 7) == 1 || (pi & 7) == 4 || (pi & 7) == 6));
        if (poly) ex2poly2(x, e[i], e[i + 1]);
        else { float x0, x1; un2(x, x0, x1); e[i] = ex2f(x0); e[i + 1] = ex2f(x1); }
        if (V == 1) { if (pi & 1) acc1 = fadd2(acc1, pk2(e[i], e[i + 1])); else acc0 = fadd2(acc0, pk2(e[i], e[i + 1])); }
      }
      if (V == 1) { float a, b; un2(fadd2(acc0, acc1), a, b); lrun += a + b; }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4 u;
      u.x = packh2(e[q * 8], e[q * 8 + 1]); u.y = packh2(e[q * 8 + 2], e[q * 8 + 3]);
      u.z = packh2(e[q * 8 + 4], e[q * 8 + 5]); u.w = packh2(e[q * 8 + 6], e[q * 8 + 7]);
      *reinterpret_cast<uint4*>(prow + ((q ^ (r % 8)) * 16)) = u;
    }
    // next tile's "scores": cheap data-dependent refresh so nothing is loop invariant (2 ALU ops per 8 elements)
#pragma unroll
    for (int i = 0; i < 64; i += 8) sv[i] = __int_as_float(__float_as_int(sv[i]) ^ (t & 1));
  }
  if (lrun == 123.456f || smem[threadIdx.x] == 77) out[0] = lrun;
}

template <int V>
void run(const char* name, int ctas_per_sm) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float *d, *in; cudaMalloc(&d, 4); cudaMalloc(&in, 4096 * 4);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = -0.01f * (i % 997);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int smem = 200 * 1024 / ctas_per_sm;                 // occupancy limiter: exactly ctas_per_sm CTAs of 4 warps per SM
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles = 4096, blocks = sms * ctas_per_sm;
  k<V><<<blocks, 128, smem>>>(in, d, 16, 0.18f, 1.0f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<V><<<blocks, 128, smem>>>(in, d, tiles, 0.18f, 1.0f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double el = (double)blocks * 128 * tiles * 64;
  printf("%-44s %2d warps/SM %8.3f ms  %6.2f elements/clk/SM (%d MHz nominal)  %s\n", name, ctas_per_sm * 4, ms,
         el / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d); cudaFree(in);
}

int main() {
  for (int c = 2; c <= 4; c += 2) {
    run<0>("V0 FFMA + MUFU + FADD + F2FP + STS", c);
    run<1>("V1 FFMA2 + MUFU + FADD2 + F2FP + STS", c);
    run<2>("V2 V1 without the row sum", c);
    run<3>("V3 V2, 1/4 of the pairs via polynomial", c);
    run<5>("V5 V2, 3/8 of the pairs via polynomial", c);
    run<4>("V4 V2, 1/2 of the pairs via polynomial", c);
  }
  return 0;
}
