// Fused gradient-norm clip + AdamW over one flat fp32 parameter buffer (the LoRA A/B matrices).
// Replaces accelerator.clip_grad_norm_ + torch.optim.AdamW.step (training_script.py:661-664, :692-694): the
// reference launches hundreds of foreach kernels and host-syncs on the norm; here the norm stays on the device and the
// 1/world_size of the data-parallel all-reduce and the clip coefficient are folded into the update.
#include "common.cuh"

namespace comat {

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  pdl_grid_dependency_sync();
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += g[i] * g[i];
  __shared__ float sm[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = sm[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
  }
}
__global__ void sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  pdl_grid_dependency_sync();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void __launch_bounds__(256) adamw_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                         float wd, float bc1, float bc2_sqrt, float max_norm, float grad_scale,
                                                         const float* __restrict__ sumsq) {
  pdl_grid_dependency_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float coef = grad_scale;
  if (max_norm > 0.f) {
    const float total = sqrtf(sumsq[0]) * grad_scale;
    coef *= fminf(1.f, max_norm / (total + 1e-6f));                 // torch.nn.utils.clip_grad_norm_
  }
  const float gi = g[i] * coef;
  float pi = p[i] * (1.f - lr * wd);                                // decoupled weight decay
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}

}  // namespace comat
using namespace comat;

extern "C" int comat_grad_sumsq(const float* g, long long n, float* partial /* >= 1024 floats */, float* out, void* stream) {
  if (!g || !partial || !out || n <= 0) return COMAT_ERR_INVALID;
  int blocks = num_sms() * 4;
  if (blocks > 1024) blocks = 1024;
  launch_k(sumsq_partial_kernel, blocks, 256, 0, (cudaStream_t)stream, g, n, partial);
  launch_k(sumsq_final_kernel, 1, 32, 0, (cudaStream_t)stream, partial, blocks, out);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

extern "C" int comat_adamw_clip(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                                float eps, float weight_decay, int step, float max_norm, float grad_scale, const float* sumsq,
                                void* stream) {
  if (!p || !g || !m || !v || n <= 0 || step < 1 || (max_norm > 0.f && !sumsq)) return COMAT_ERR_INVALID;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  launch_k(adamw_clip_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                                  bc1, bc2s, max_norm, grad_scale, sumsq);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Fused classifier-free guidance + DDPM ancestral step on the fp32 latent chain (TrainableSDPipeline.py:155-167):
//   x_prev = c_x * x + c_eps * (e_u + s * (e_c - e_u)) + sigma * z        (epsilon prediction folded into c_x, c_eps)
// replaces ~10 aten launches per sampler step; backward is the matching linear map.
namespace comat {
__global__ void __launch_bounds__(256) cfg_ddpm_fwd_kernel(const float* __restrict__ eps2, const float* __restrict__ x,
                                                           const float* __restrict__ z, float* __restrict__ out, long long n,
                                                           float s, float c_eps, float c_x, float sigma, int cfg) {
  pdl_grid_dependency_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e = cfg ? (eps2[i] + s * (eps2[n + i] - eps2[i])) : eps2[i];
  float r = c_x * x[i] + c_eps * e;
  if (z != nullptr) r += sigma * z[i];
  out[i] = r;
}
__global__ void __launch_bounds__(256) cfg_ddpm_bwd_kernel(const float* __restrict__ g, float* __restrict__ d_eps2,
                                                           float* __restrict__ dx, long long n, float s, float c_eps, float c_x, int cfg) {
  pdl_grid_dependency_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  if (dx != nullptr) dx[i] = c_x * gi;
  if (d_eps2 != nullptr) {
    if (cfg) { d_eps2[i] = c_eps * (1.f - s) * gi; d_eps2[n + i] = c_eps * s * gi; }
    else d_eps2[i] = c_eps * gi;
  }
}
}  // namespace comat

extern "C" int comat_cfg_ddpm_step_fwd(const float* eps, const float* x, const float* noise, float* out, long long n, float guidance,
                                       float c_eps, float c_x, float sigma, int cfg, void* stream) {
  if (!eps || !x || !out || n <= 0) return COMAT_ERR_INVALID;
  launch_k(comat::cfg_ddpm_fwd_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, eps, x, noise, out, n, guidance, c_eps, c_x, sigma, cfg);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
extern "C" int comat_cfg_ddpm_step_bwd(const float* grad_out, float* d_eps, float* dx, long long n, float guidance, float c_eps,
                                       float c_x, int cfg, void* stream) {
  if (!grad_out || n <= 0) return COMAT_ERR_INVALID;
  launch_k(comat::cfg_ddpm_bwd_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, grad_out, d_eps, dx, n, guidance, c_eps, c_x, cfg);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
