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

// Overflow guard (the reference relies on accelerate's GradScaler, which skips the optimiser step when the unscaled gradients
// hold an inf / NaN): one thread turns the gradient's sum of squares into the step's constants.  A non-finite sum leaves the
// device step counter alone and raises the skip flag, and the update kernel then returns without touching p, m, v.
//   state[0] = skip (0 / 1)   state[1] = clip coefficient * grad_scale   state[2] = 1 - beta1^step   state[3] = sqrt(1 - beta2^step)
//   counters[0] = optimiser steps taken   counters[1] = steps skipped   counters[2] = 1 if the LAST step was skipped
__global__ void adamw_prepare_kernel(const float* __restrict__ sumsq, float* __restrict__ state, int* __restrict__ counters,
                                     double b1, double b2, double lr, float max_norm, float grad_scale) {
  pdl_grid_dependency_sync();
  if (threadIdx.x != 0) return;
  const float ss = sumsq[0];
  if (!isfinite(ss)) {
    state[0] = 1.f; state[1] = 0.f; state[2] = 0.f; state[3] = 1.f;
    counters[1] += 1; counters[2] = 1;
    return;
  }
  const int step = counters[0] + 1;
  counters[0] = step; counters[2] = 0;
  float coef = grad_scale;
  if (max_norm > 0.f) coef *= fminf(1.f, max_norm / (sqrtf(ss) * grad_scale + 1e-6f));      // torch.nn.utils.clip_grad_norm_
  // torch.optim.AdamW evaluates the bias corrections in Python doubles and hands the kernels their fp32 roundings
  const double bc1 = 1.0 - pow(b1, (double)step);
  const double bc2 = 1.0 - pow(b2, (double)step);
  state[0] = 0.f; state[1] = coef;
  state[2] = (float)(lr / bc1);                      // step_size
  state[3] = (float)sqrt(bc2);                       // bias_correction2_sqrt
}

// torch's single-tensor AdamW arithmetic, operation by operation (fp32):  p *= 1 - lr*wd ;  m = lerp(m, g, 1 - b1) ;
// v = v*b2 + (1 - b2)*g*g ;  p -= step_size * m / (sqrt(v) / bc2_sqrt + eps)
__global__ void __launch_bounds__(256) adamw_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long n, float decay, float b2, float omb1,
                                                         float omb2, float eps, const float* __restrict__ state) {
  pdl_grid_dependency_sync();
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n || state[0] != 0.f) return;
  const float coef = state[1], step_size = state[2], bc2_sqrt = state[3];
  auto upd = [&](float& pp, float& mm, float& vv, float gg) {
    gg *= coef;
    mm = omb1 < 0.5f ? fmaf(omb1, gg - mm, mm) : gg - (gg - mm) * (1.f - omb1);      // at::lerp
    vv = fmaf(omb2 * gg, gg, vv * b2);
    pp = pp * decay - step_size * (mm / (sqrtf(vv) / bc2_sqrt + eps));
  };
  if (i4 + 4 <= n) {
    const float4 g4 = *reinterpret_cast<const float4*>(g + i4);
    float4 p4 = *reinterpret_cast<float4*>(p + i4), m4 = *reinterpret_cast<float4*>(m + i4), v4 = *reinterpret_cast<float4*>(v + i4);
    upd(p4.x, m4.x, v4.x, g4.x); upd(p4.y, m4.y, v4.y, g4.y); upd(p4.z, m4.z, v4.z, g4.z); upd(p4.w, m4.w, v4.w, g4.w);
    *reinterpret_cast<float4*>(p + i4) = p4; *reinterpret_cast<float4*>(m + i4) = m4; *reinterpret_cast<float4*>(v + i4) = v4;
  } else {
    for (long long i = i4; i < n; ++i) upd(p[i], m[i], v[i], g[i]);
  }
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

extern "C" int comat_adamw_clip(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2,
                                double eps, double weight_decay, float max_norm, float grad_scale, const float* sumsq,
                                float* state /* 4 floats */, int* counters /* 3 ints: steps, skipped, last skipped */, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || !sumsq || !state || !counters) return COMAT_ERR_INVALID;
  if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) != 0) return COMAT_ERR_INVALID;
  launch_k(adamw_prepare_kernel, 1, 32, 0, (cudaStream_t)stream, sumsq, state, counters, beta1, beta2, lr, max_norm, grad_scale);
  // hyper-parameters arrive as doubles (Python floats) and are rounded to fp32 exactly where torch rounds them
  launch_k(adamw_clip_kernel, (unsigned)((n + 1023) / 1024), 256, 0, (cudaStream_t)stream, p, g, m, v, n, (float)(1.0 - lr * weight_decay),
           (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (const float*)state);
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
