// Cross-entropy with label smoothing over the BLIP vocabulary (30 524), mean over non-ignored rows — the loss of
// HF BlipTextLMHeadModel.forward (modeling_blip_text.py:764-775; reached from concept_mat_utils/caption_blip.py:57) — and its
// backward, which emits the 16-bit, zero-padded dlogits the LM-head dgrad GEMM consumes.  Logits are never copied or
// re-materialised: 3 streaming passes over each fp32 row forward, one pass backward.
#include "common.cuh"

namespace comat {

__device__ __forceinline__ float block_reduce(float v, float* sm, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sm[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, sm[i]) : (r + sm[i]);
  return r;
}

// one CTA per row: stats[r] = {lse, loss_row}
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                      float* __restrict__ stats, int V, long long ld, float eps, long long ignore) {
  pdl_grid_dependency_sync();
  __shared__ float sm[8];
  const int r = blockIdx.x;
  const float* x = logits + (size_t)r * ld;
  const long long y = labels[r];
  float m = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) m = fmaxf(m, x[v]);
  m = block_reduce(m, sm, true);
  float se = 0.f, sx = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) { const float t = x[v]; se += __expf(t - m); sx += t; }
  se = block_reduce(se, sm, false);
  sx = block_reduce(sx, sm, false);
  if (threadIdx.x == 0) {
    const float lse = m + logf(se);
    float loss = 0.f;
    if (y != ignore) loss = (1.f - eps) * (lse - x[y]) + eps * (lse - sx / (float)V);
    stats[r * 2] = lse;
    stats[r * 2 + 1] = loss;
  }
}
// out[0] = mean loss over non-ignored rows, out[1] = count (fixed-order sum)
__global__ void ce_mean_kernel(const float* __restrict__ stats, const long long* __restrict__ labels, float* __restrict__ out, int R,
                               long long ignore) {
  pdl_grid_dependency_sync();
  if (threadIdx.x != 0) return;
  float s = 0.f, c = 0.f;
  for (int r = 0; r < R; ++r)
    if (labels[r] != ignore) { s += stats[r * 2 + 1]; c += 1.f; }
  out[0] = c > 0.f ? s / c : 0.f;
  out[1] = c;
}
template <typename T>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                     const float* __restrict__ stats, const float* __restrict__ out2,
                                                     const float* __restrict__ gout, T* __restrict__ dlogits, int V, int Vpad, long long ld,
                                                     float eps, long long ignore) {
  pdl_grid_dependency_sync();
  const int r = blockIdx.x;
  const long long y = labels[r];
  T* d = dlogits + (size_t)r * Vpad;
  if (y == ignore) {
    for (int v = threadIdx.x; v < Vpad; v += blockDim.x) d[v] = from_f32<T>(0.f);
    return;
  }
  const float* x = logits + (size_t)r * ld;
  const float lse = stats[r * 2];
  const float g = gout[0] / out2[1];
  const float sm = eps / (float)V;
  for (int v = threadIdx.x; v < Vpad; v += blockDim.x) {
    float val = 0.f;
    if (v < V) val = g * (__expf(x[v] - lse) - ((v == y) ? (1.f - eps) : 0.f) - sm);
    d[v] = from_f32<T>(val);
  }
}
}  // namespace comat
using namespace comat;

extern "C" int comat_ce_label_smooth_fwd(const float* logits, const long long* labels, float* row_stats, float* out2, int R, int V,
                                         long long ld, float eps, long long ignore_index, void* stream) {
  if (!logits || !labels || !row_stats || !out2 || R <= 0 || V <= 0) return COMAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(ce_rows_kernel, R, 256, 0, st, logits, labels, row_stats, V, ld, eps, ignore_index);
  launch_k(ce_mean_kernel, 1, 32, 0, st, row_stats, labels, out2, R, ignore_index);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
extern "C" int comat_ce_label_smooth_bwd(const float* logits, const long long* labels, const float* row_stats, const float* out2,
                                         const float* grad_out, void* dlogits16, int R, int V, int Vpad, long long ld, float eps,
                                         long long ignore_index, int dtype, void* stream) {
  if (!logits || !labels || !row_stats || !out2 || !grad_out || !dlogits16 || Vpad < V) return COMAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == COMAT_F16) launch_k(ce_bwd_kernel<__half>, R, 256, 0, st, logits, labels, row_stats, out2, grad_out, (__half*)dlogits16, V, Vpad, ld, eps, ignore_index);
  else if (dtype == COMAT_BF16) launch_k(ce_bwd_kernel<__nv_bfloat16>, R, 256, 0, st, logits, labels, row_stats, out2, grad_out, (__nv_bfloat16*)dlogits16, V, Vpad, ld, eps, ignore_index);
  else return COMAT_ERR_UNSUPPORTED;
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
