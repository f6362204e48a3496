// Discriminator head of the fidelity GAN, fused: per-pixel Linear(4, 1) on the D UNet's noise prediction + BCEWithLogits against an
// all-ones (generator side) or [zeros | ones] (discriminator side) target, mean over n * H * W logits - forward, and backward for the
// prediction, the head's weight and bias.  Replaces permute(0,2,3,1) -> nn.Linear(4,1) -> nn.BCEWithLogitsLoss of
// training_utils/gan_sdxl.py:31-34, :84-89, :118-132 (three aten launches forward and ~six backward per GAN pass).
//   eps (n, C, H, W) fp32 NCHW (C <= 8), w (C) fp32, b (1) fp32.  Samples b < n_zero have target 0, the others target 1.
#include "common.cuh"

namespace comat {

constexpr int GH_MAXC = 8;

__device__ __forceinline__ float gh_logit(const float* __restrict__ eps, const float* w, float bias, int C, long long hw, long long base) {
  float x = bias;
#pragma unroll
  for (int c = 0; c < GH_MAXC; ++c)
    if (c < C) x = fmaf(w[c], eps[base + (long long)c * hw], x);
  return x;
}

// loss_sum[0] += sum_i  max(x,0) - x*t + log1p(exp(-|x|))     (numerically stable BCE-with-logits, as aten computes it)
__global__ void __launch_bounds__(256) gan_head_fwd_kernel(const float* __restrict__ eps, const float* __restrict__ w, const float* __restrict__ b,
                                                           float* __restrict__ loss_sum, int n, int C, long long hw, int n_zero) {
  pdl_grid_dependency_sync();
  __shared__ float sm[8];
  float wr[GH_MAXC];
#pragma unroll
  for (int c = 0; c < GH_MAXC; ++c) wr[c] = c < C ? w[c] : 0.f;
  const float bias = b[0];
  const long long total = (long long)n * hw;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / hw, px = i - s * hw;
    const float x = gh_logit(eps, wr, bias, C, hw, s * C * hw + px);
    const float t = s < n_zero ? 0.f : 1.f;
    acc += fmaxf(x, 0.f) - x * t + log1pf(__expf(-fabsf(x)));
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = sm[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(loss_sum, v);
  }
}

// d_eps[s, c, px] = g * w[c] ;  dw[c] += sum g * eps[s, c, px] ;  db += sum g,   g = gout * (sigmoid(x) - t) / (n * hw)
__global__ void __launch_bounds__(256) gan_head_bwd_kernel(const float* __restrict__ eps, const float* __restrict__ w, const float* __restrict__ b,
                                                           const float* __restrict__ gout, float* __restrict__ d_eps, float* __restrict__ dw,
                                                           float* __restrict__ db, int n, int C, long long hw, int n_zero) {
  pdl_grid_dependency_sync();
  __shared__ float sm[8][GH_MAXC + 1];
  float wr[GH_MAXC], aw[GH_MAXC];
#pragma unroll
  for (int c = 0; c < GH_MAXC; ++c) { wr[c] = c < C ? w[c] : 0.f; aw[c] = 0.f; }
  const float bias = b[0];
  const long long total = (long long)n * hw;
  const float k = gout[0] / (float)total;
  float ab = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / hw, px = i - s * hw, base = s * C * hw + px;
    const float x = gh_logit(eps, wr, bias, C, hw, base);
    const float t = s < n_zero ? 0.f : 1.f;
    const float g = k * (1.f / (1.f + __expf(-x)) - t);
    ab += g;
#pragma unroll
    for (int c = 0; c < GH_MAXC; ++c) {
      if (c < C) {
        aw[c] = fmaf(g, eps[base + (long long)c * hw], aw[c]);
        if (d_eps != nullptr) d_eps[base + (long long)c * hw] = g * wr[c];
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < GH_MAXC; ++c) aw[c] = warp_sum(aw[c]);
  ab = warp_sum(ab);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < GH_MAXC; ++c) sm[warp][c] = aw[c];
    sm[warp][GH_MAXC] = ab;
  }
  __syncthreads();
  if (threadIdx.x <= GH_MAXC) {
    float v = 0.f;
    for (int wv = 0; wv < 8; ++wv) v += sm[wv][threadIdx.x];
    if (threadIdx.x < C && dw != nullptr) atomicAdd(&dw[threadIdx.x], v);
    if (threadIdx.x == GH_MAXC && db != nullptr) atomicAdd(db, v);
  }
}

}  // namespace comat
using namespace comat;

// loss[0] = mean BCEWithLogits(w . eps + b, target).  `loss` must be zeroed by the caller (the kernel accumulates the sum; the mean
// is taken by comat_gan_head_bce_bwd's 1/N and by the caller's scale for the forward value: loss_sum / (n*H*W)).
extern "C" int comat_gan_head_bce_fwd(const float* eps, const float* w, const float* b, float* loss_sum, int n, int C, int HW, int n_zero,
                                      void* stream) {
  if (!eps || !w || !b || !loss_sum || n <= 0 || C <= 0 || C > GH_MAXC || HW <= 0 || n_zero < 0 || n_zero > n) return COMAT_ERR_INVALID;
  const long long total = (long long)n * HW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
  launch_k(gan_head_fwd_kernel, blocks, 256, 0, (cudaStream_t)stream, eps, w, b, loss_sum, n, C, (long long)HW, n_zero);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

// d_eps (n, C, H, W) may be null (discriminator side: the latents need no gradient); dw (C) / db (1) are ACCUMULATED into (fp32
// atomics) and may be null (generator side: the head is frozen).  gout = upstream gradient of the scalar loss (device scalar).
extern "C" int comat_gan_head_bce_bwd(const float* eps, const float* w, const float* b, const float* gout, float* d_eps, float* dw, float* db,
                                      int n, int C, int HW, int n_zero, void* stream) {
  if (!eps || !w || !b || !gout || n <= 0 || C <= 0 || C > GH_MAXC || HW <= 0 || n_zero < 0 || n_zero > n) return COMAT_ERR_INVALID;
  const long long total = (long long)n * HW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
  launch_k(gan_head_bwd_kernel, blocks, 256, 0, (cudaStream_t)stream, eps, w, b, gout, d_eps, dw, db, n, C, (long long)HW, n_zero);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
