// HBM-bound normalisation / activation / rearrangement kernels of the UNet, VAE and BLIP paths (fwd + bwd).
// All activations are 16-bit (fp16 or bf16) row-major [rows, C] (images are NHWC), statistics and math in fp32.
// Base weights are frozen on this path (training_utils/pipeline.py:66-71), so no gamma/beta gradients are produced.
#include "common.cuh"

namespace comat {

template <typename T>
struct Vec8 {
  uint4 u;
  __device__ __forceinline__ void load(const T* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(T* p) const { *reinterpret_cast<uint4*>(p) = u; }
  __device__ __forceinline__ float get(int i) const { return to_f32<T>(reinterpret_cast<const T*>(&u)[i]); }
  __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<T*>(&u)[i] = from_f32<T>(v); }
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_grad(float x) {
  const float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// ============================================================================================== GroupNorm
// x: (n, HW, C), G groups, cpg = C/G.  Pass 1: per-(n, chunk) partial sums of (a, b) per group, where
//   stats  pass: a = x,          b = x^2
//   bwd    pass: a = dy*gamma*act'(.), b = a * xhat
// Pass 2 (apply) consumes the chunk-reduced sums.  Two passes => x is read twice (algorithmic minimum for a
// normalisation whose statistics span the whole image) and written once.
// Both kernels use the same thread layout: a block covers rows_par rows x (C/8) 16-byte column vectors, each thread
// keeps its 8 channels' constants in registers and walks rows with four independent 16-byte loads in flight.

template <typename T, int MODE, int GN_UNROLL>   // MODE 0: stats of x ; MODE 1: backward sums.  GN_UNROLL = 16-byte loads in flight per thread
__global__ void __launch_bounds__(512) gn_partial_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const float* __restrict__ mean_rstd, float* __restrict__ part,
                                                                int HW, int C, int G, int chunks, int silu) {
  pdl_grid_dependency_sync();
  extern __shared__ float sm[];   // [2*C]
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int vcols = C / 8;
  const int rows_par = blockDim.x / vcols;
  const int v = threadIdx.x % vcols, pr = threadIdx.x / vcols;
  const int cpg = C / G;
  const int p_begin = (int)((long long)HW * chunk / chunks), p_end = (int)((long long)HW * (chunk + 1) / chunks);
  float sa[8], sb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sa[i] = sb[i] = 0.f;
  if (pr < rows_par) {
    float g8[8], b8[8], mu[8], rs[8];
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = v * 8 + i, gidx = c / cpg;
        g8[i] = gamma[c]; b8[i] = beta[c];
        rs[i] = mean_rstd[((size_t)n * G + gidx) * 2 + 1]; mu[i] = -mean_rstd[((size_t)n * G + gidx) * 2] * rs[i];   // xh = x*rs + mu
      }
    }
    const T* xb = x + (size_t)n * HW * C + v * 8;
    const T* db = (MODE == 1) ? dy + (size_t)n * HW * C + v * 8 : nullptr;
    for (int p = p_begin + pr; p < p_end; p += GN_UNROLL * rows_par) {
      Vec8<T> xv[GN_UNROLL], dv[GN_UNROLL];
#pragma unroll
      for (int k = 0; k < GN_UNROLL; ++k) {
        const int pk = p + k * rows_par;
        if (pk < p_end) {
          xv[k].load(xb + (size_t)pk * C);
          if (MODE == 1) dv[k].load(db + (size_t)pk * C);
        }
      }
#pragma unroll
      for (int k = 0; k < GN_UNROLL; ++k) {
        if (p + k * rows_par < p_end) {
          if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float f = xv[k].get(i); sa[i] += f; sb[i] = fmaf(f, f, sb[i]); }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float xh = fmaf(xv[k].get(i), rs[i], mu[i]);
              float d = dv[k].get(i);
              if (silu) d *= silu_grad(fmaf(xh, g8[i], b8[i]));
              const float a = d * g8[i];
              sa[i] += a; sb[i] = fmaf(a, xh, sb[i]);
            }
          }
        }
      }
    }
  }
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  if (pr < rows_par) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { atomicAdd(&sm[v * 8 + i], sa[i]); atomicAdd(&sm[C + v * 8 + i], sb[i]); }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += sm[c]; b += sm[C + c]; }
    float* o = part + (((size_t)n * chunks + chunk) * G + g) * 2;
    o[0] = a; o[1] = b;
  }
}

// Pass 2.  Every block first reduces the chunk partials of its image (fixed order, so the result does not depend on the
// grid) into per-group (mean, rstd) [MODE 0; block 0 of each image also saves them for the backward] or the backward's
// (mean(a), mean(a*xhat)) [MODE 1]; then y = x*A + B with A = rstd*gamma, B = beta - mean*A  (or dx) streams through.
template <typename T, int MODE, int GN_UNROLL>   // MODE 0: y = [silu](xhat*gamma+beta) ; MODE 1: dx
__global__ void __launch_bounds__(512) gn_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ out,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ part, float* __restrict__ mean_rstd,
                                                       int HW, int C, int G, int chunks, int achunks, float inv_cnt, float eps, int silu) {
  pdl_grid_dependency_sync();
  extern __shared__ float sm[];   // [2*G]
  const int n = blockIdx.y, chunk = blockIdx.x;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int c = 0; c < chunks; ++c) {
      const float* pp = part + (((size_t)n * chunks + c) * G + g) * 2;
      a += pp[0]; b += pp[1];
    }
    if (MODE == 0) {
      const float mean = a * inv_cnt;
      const float var = fmaxf(b * inv_cnt - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
      sm[g] = mean; sm[G + g] = rstd;
      if (chunk == 0) { mean_rstd[((size_t)n * G + g) * 2] = mean; mean_rstd[((size_t)n * G + g) * 2 + 1] = rstd; }
    } else {
      sm[g] = a * inv_cnt; sm[G + g] = b * inv_cnt;
    }
  }
  __syncthreads();
  const int vcols = C / 8;
  const int rows_par = blockDim.x / vcols;
  const int v = threadIdx.x % vcols, pr = threadIdx.x / vcols;
  if (pr >= rows_par) return;
  const int cpg = C / G;
  const int p_begin = (int)((long long)HW * chunk / achunks), p_end = (int)((long long)HW * (chunk + 1) / achunks);
  float A[8], B[8], g8[8], b8[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = v * 8 + i, g = c / cpg;
    if (MODE == 0) {
      A[i] = sm[G + g] * gamma[c];
      B[i] = fmaf(-sm[g], A[i], beta[c]);
    } else {
      const float rs = mean_rstd[((size_t)n * G + g) * 2 + 1];
      A[i] = rs; B[i] = -mean_rstd[((size_t)n * G + g) * 2] * rs;      // xh = x*A + B
      g8[i] = gamma[c]; b8[i] = beta[c]; m1[i] = sm[g]; m2[i] = sm[G + g];
    }
  }
  const T* xb = x + (size_t)n * HW * C + v * 8;
  const T* db = (MODE == 1) ? dy + (size_t)n * HW * C + v * 8 : nullptr;
  T* ob = out + (size_t)n * HW * C + v * 8;
  for (int p = p_begin + pr; p < p_end; p += GN_UNROLL * rows_par) {
    Vec8<T> xv[GN_UNROLL], dv[GN_UNROLL];
#pragma unroll
    for (int k = 0; k < GN_UNROLL; ++k) {
      const int pk = p + k * rows_par;
      if (pk < p_end) {
        xv[k].load(xb + (size_t)pk * C);
        if (MODE == 1) dv[k].load(db + (size_t)pk * C);
      }
    }
#pragma unroll
    for (int k = 0; k < GN_UNROLL; ++k) {
      const int pk = p + k * rows_par;
      if (pk < p_end) {
        Vec8<T> o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (MODE == 0) {
            float y = fmaf(xv[k].get(i), A[i], B[i]);
            if (silu) y = silu_f(y);
            o.set(i, y);
          } else {
            const float xh = fmaf(xv[k].get(i), A[i], B[i]);
            float d = dv[k].get(i);
            if (silu) d *= silu_grad(fmaf(xh, g8[i], b8[i]));
            const float a = d * g8[i];
            o.set(i, A[i] * (a - m1[i] - xh * m2[i]));
          }
        }
        o.store(ob + (size_t)pk * C);
      }
    }
  }
}

// ============================================================================================== LayerNorm
// A warp normalises R rows at a time: all R x NV 16-byte loads of its rows are issued before the first reduction, so an SM
// keeps >= 40 KB in flight (the one-row-per-warp version had ~20 KB per SM outstanding and ran the UNet's 64x64-level
// LayerNorms - 32768 rows x 320 channels - at ~45 % of the HBM rate, profiles/r01_step_kernels_v5.md).  Two-pass statistics
// on the register copy (mean, then centred sum of squares).
template <typename T, int MODE, int NV, int R>   // MODE 0 fwd (saves mean,rstd) ; 1 bwd.  NV = 16-byte vectors per lane, R = rows per warp
__global__ void __launch_bounds__(256) ln_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ out,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 float* __restrict__ mean_rstd, long long rows, int C, float eps) {
  pdl_grid_dependency_sync();
  const long long row0 = ((long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * R;
  if (row0 >= rows) return;
  const int lane = threadIdx.x & 31;
  const int vcols = C / 8;
  const float inv_c = 1.f / (float)C;
  Vec8<T> xv[R][NV];
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + k * 32;
      if (v < vcols && row0 + r < rows) xv[r][k].load(x + (size_t)(row0 + r) * C + v * 8);
    }
  }
  if (MODE == 0) {
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (lane + k * 32 < vcols && row0 + r < rows) {
#pragma unroll
          for (int i = 0; i < 8; ++i) s += xv[r][k].get(i);
        }
      }
      mean[r] = s;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] = warp_sum(mean[r]) * inv_c;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (lane + k * 32 < vcols && row0 + r < rows) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float d = xv[r][k].get(i) - mean[r]; q = fmaf(d, d, q); }
        }
      }
      rstd[r] = q;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(warp_sum(rstd[r]) * inv_c + eps);
    if (lane < R && row0 + lane < rows) {
      float m = mean[0], rs = rstd[0];
#pragma unroll
      for (int r = 1; r < R; ++r) if (lane == r) { m = mean[r]; rs = rstd[r]; }
      mean_rstd[(row0 + lane) * 2] = m; mean_rstd[(row0 + lane) * 2 + 1] = rs;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + k * 32;
      if (v < vcols) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8), b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
        const float g8[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b8[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (row0 + r < rows) {
            Vec8<T> o;
#pragma unroll
            for (int i = 0; i < 8; ++i) o.set(i, (xv[r][k].get(i) - mean[r]) * rstd[r] * g8[i] + b8[i]);
            o.store(out + (size_t)(row0 + r) * C + v * 8);
          }
        }
      }
    }
  } else {
    Vec8<T> dv[R][NV];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int v = lane + k * 32;
        if (v < vcols && row0 + r < rows) dv[r][k].load(dy + (size_t)(row0 + r) * C + v * 8);
      }
    }
    float mean[R], rstd[R], m1[R], m2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long rr = row0 + r < rows ? row0 + r : rows - 1;
      mean[r] = mean_rstd[rr * 2]; rstd[r] = mean_rstd[rr * 2 + 1];
      m1[r] = 0.f; m2[r] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + k * 32;
      if (v < vcols) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
        const float g8[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (row0 + r < rows) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float a = dv[r][k].get(i) * g8[i];
              m1[r] += a; m2[r] = fmaf(a, (xv[r][k].get(i) - mean[r]) * rstd[r], m2[r]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) { m1[r] = warp_sum(m1[r]) * inv_c; m2[r] = warp_sum(m2[r]) * inv_c; }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + k * 32;
      if (v < vcols) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
        const float g8[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (row0 + r < rows) {
            Vec8<T> o;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float a = dv[r][k].get(i) * g8[i];
              const float xh = (xv[r][k].get(i) - mean[r]) * rstd[r];
              o.set(i, rstd[r] * (a - m1[r] - xh * m2[r]));
            }
            o.store(out + (size_t)(row0 + r) * C + v * 8);
          }
        }
      }
    }
  }
}

// ============================================================================================== GEGLU  (hidden * gelu(gate))
template <typename T, int MODE>
__global__ void __launch_bounds__(256) geglu_kernel(const T* __restrict__ hg, const T* __restrict__ dy, T* __restrict__ out,
                                                    long long rows, int Ch) {
  pdl_grid_dependency_sync();   // hg: [rows, 2*Ch] ; out fwd [rows,Ch], bwd [rows,2*Ch]
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int vcols = Ch / 8;
  if (idx >= rows * vcols) return;
  const long long row = idx / vcols;
  const int v = (int)(idx % vcols);
  Vec8<T> h, g;
  h.load(hg + (size_t)row * 2 * Ch + v * 8);
  g.load(hg + (size_t)row * 2 * Ch + Ch + v * 8);
  if (MODE == 0) {
    Vec8<T> o;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.set(i, h.get(i) * gelu_f(g.get(i)));
    o.store(out + (size_t)row * Ch + v * 8);
  } else {
    Vec8<T> d, oh, og;
    d.load(dy + (size_t)row * Ch + v * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gi = g.get(i), di = d.get(i);
      oh.set(i, di * gelu_f(gi));
      og.set(i, di * h.get(i) * gelu_grad(gi));
    }
    oh.store(out + (size_t)row * 2 * Ch + v * 8);
    og.store(out + (size_t)row * 2 * Ch + Ch + v * 8);
  }
}

// ============================================================================================== unary / binary elementwise
// op: 0 silu  1 silu_bwd(x, dy)  2 gelu  3 gelu_bwd(x, dy)  4 add(x, y)  5 scale(x)*alpha  6 axpby: alpha*x + beta*y
template <typename T>
__global__ void __launch_bounds__(256) ew_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ out, long long nvec,
                                                 int op, float alpha, float beta) {
  pdl_grid_dependency_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nvec) return;
  Vec8<T> a, b, o;
  a.load(x + idx * 8);
  if (y != nullptr) b.load(y + idx * 8);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float xa = a.get(i), yb = (y != nullptr) ? b.get(i) : 0.f;
    float r;
    switch (op) {
      case 0: r = silu_f(xa); break;
      case 1: r = yb * silu_grad(xa); break;
      case 2: r = gelu_f(xa); break;
      case 3: r = yb * gelu_grad(xa); break;
      case 4: r = xa + yb; break;
      case 5: r = xa * alpha; break;
      default: r = alpha * xa + beta * yb; break;
    }
    o.set(i, r);
  }
  o.store(out + idx * 8);
}

// ============================================================================================== spatial rearrangements (NHWC)
// mode 0: nearest x2 upsample fwd (n,H,W,C)->(n,2H,2W,C)     mode 1: its backward (sum of the 2x2 block)
// mode 2: space-to-depth (n,H,W,C)->(n,H/2,W/2,4C), channel block order (dy,dx)   mode 3: depth-to-space (inverse)
template <typename T>
__global__ void __launch_bounds__(256) spatial_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int H, int W, int C, int mode) {
  pdl_grid_dependency_sync();
  // H, W are the dims of the *smaller* tensor for modes 0/1 and of the *larger* tensor for modes 2/3
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int vcols = C / 8;
  if (mode == 0 || mode == 1) {
    if (mode == 0) {
      const long long total = (long long)n * 2 * H * 2 * W * vcols;
      if (idx >= total) return;
      const int v = (int)(idx % vcols);
      long long r = idx / vcols;
      const int x = (int)(r % (2 * W)); r /= 2 * W;
      const int y = (int)(r % (2 * H));
      const int b = (int)(r / (2 * H));
      Vec8<T> t; t.load(in + (((size_t)b * H + y / 2) * W + x / 2) * C + v * 8);
      t.store(out + (((size_t)b * 2 * H + y) * 2 * W + x) * C + v * 8);
    } else {
      const long long total = (long long)n * H * W * vcols;
      if (idx >= total) return;
      const int v = (int)(idx % vcols);
      long long r = idx / vcols;
      const int x = (int)(r % W); r /= W;
      const int y = (int)(r % H);
      const int b = (int)(r / H);
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          Vec8<T> t; t.load(in + (((size_t)b * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dx) * C + v * 8);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += t.get(i);
        }
      Vec8<T> o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.set(i, acc[i]);
      o.store(out + (((size_t)b * H + y) * W + x) * C + v * 8);
    }
  } else {
    const long long total = (long long)n * H * W * vcols;
    if (idx >= total) return;
    const int v = (int)(idx % vcols);
    long long r = idx / vcols;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    const size_t big = (((size_t)b * H + y) * W + x) * C + v * 8;
    const size_t small = (((size_t)b * (H / 2) + y / 2) * (W / 2) + x / 2) * (4 * C) + ((y & 1) * 2 + (x & 1)) * C + v * 8;
    Vec8<T> t;
    if (mode == 2) { t.load(in + big); t.store(out + small); }
    else           { t.load(in + small); t.store(out + big); }
  }
}

// 16-bit matrix transpose [R, Cc] -> [Cc, R] (LoRA wgrad operands), 32x32 tiles through padded smem
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int R, int Cc, int ld_out) {
  pdl_grid_dependency_sync();
  __shared__ T tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = by + j, c = bx + threadIdx.x;
    if (r < R && c < Cc) tile[j][threadIdx.x] = in[(size_t)r * Cc + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = bx + j, r = by + threadIdx.x;
    if (r < R && c < Cc) out[(size_t)c * ld_out + r] = tile[threadIdx.x][j];
  }
}

// fp32 <-> 16-bit casts with layout change for the 4-channel latents: NCHW fp32 -> NHWC 16-bit padded to Cpad channels
template <typename T>
__global__ void latent_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int n, int Cin, int HW, int Cpad, float scale) {
  pdl_grid_dependency_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * HW * Cpad) return;
  const int c = (int)(idx % Cpad);
  const long long r = idx / Cpad;
  const int p = (int)(r % HW), b = (int)(r / HW);
  out[idx] = from_f32<T>(c < Cin ? in[((size_t)b * Cin + c) * HW + p] * scale : 0.f);
}
// NHWC 16-bit (first Cout of ld channels) -> NCHW fp32
template <typename T>
__global__ void nhwc_to_nchw_f32_kernel(const T* __restrict__ in, float* __restrict__ out, int n, int Cout, int HW, int ld, float scale) {
  pdl_grid_dependency_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * Cout * HW) return;
  const int p = (int)(idx % HW);
  const long long r = idx / HW;
  const int c = (int)(r % Cout), b = (int)(r / Cout);
  out[idx] = to_f32<T>(in[((size_t)b * HW + p) * ld + c]) * scale;
}

// strided 2-D copy of 16-bit rows: dst[r, 0:cols] = src[r, 0:cols]  (torch.cat([hidden, skip], dim=1) in NHWC)
__global__ void __launch_bounds__(256) copy2d_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long rows, int vcols,
                                                     long long ld_src_v, long long ld_dst_v) {
  pdl_grid_dependency_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * vcols) return;
  const long long r = idx / vcols;
  const int v = (int)(idx % vcols);
  dst[r * ld_dst_v + v] = src[r * ld_src_v + v];
}

}  // namespace comat

using namespace comat;

#define DISPATCH_T(dtype, ...)                                       \
  if ((dtype) == COMAT_F16) { using T = __half; __VA_ARGS__; }       \
  else if ((dtype) == COMAT_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
  else return COMAT_ERR_UNSUPPORTED;

static inline int gn_threads(int C) { int v = C / 8; return v <= 256 ? 256 : ((v + 31) / 32) * 32; }
static inline int gn_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && atoi(e) > 0) ? atoi(e) : dflt;
}
// tuning knobs (measured defaults, profiles/r02_groupnorm_bench.md): loads in flight per thread and CTAs per SM of the two passes
static inline int gn_unroll() { static int u = gn_env("COMAT_GN_UNROLL", 4); return u; }
static inline int gn_chunks(int n, int HW) {
  static const int per_sm = gn_env("COMAT_GN_PARTIAL_CTAS", 2);
  int c = (per_sm * num_sms() + n - 1) / n;
  if (c > HW / 16) c = HW / 16;
  return c < 1 ? 1 : c;
}
// row blocks of the apply pass: ~8 blocks per SM so enough 16-byte loads are in flight to cover HBM latency
static inline int gn_apply_chunks(int n, int HW, int rows_par) {
  static const int per_sm = gn_env("COMAT_GN_APPLY_CTAS", 8);
  int c = (per_sm * num_sms() + n - 1) / n;
  const int cap = HW / (rows_par * gn_unroll()) > 0 ? HW / (rows_par * gn_unroll()) : 1;
  if (c > cap) c = cap;
  return c < 1 ? 1 : c;
}

namespace comat {
int gn_fused_launch(int mode, const void* x, const void* dy, void* out, const float* gamma, const float* beta, float* mean_rstd,
                    int n, int HW, int C, int G, float eps, int silu, int dtype, cudaStream_t st);     // groupnorm_fused.cu
}
// COMAT_GN=fused selects the single-launch cluster kernel (groupnorm_fused.cu).  Default: the two-launch kernels below - measured
// on B200 (profiles/r02_groupnorm_bench.md) the cluster kernel wins 10-30 % on slabs that fit in shared memory and loses up to 2x on
// the VAE-sized ones, and the whole train step is the same to 0.1 % (505.5 vs 505.0 ms), so the simpler pair stays the product path.
static inline bool gn_use_fused() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("COMAT_GN"); on = (e && !strcmp(e, "fused")) ? 1 : 0; }
  return on == 1;
}

extern "C" size_t comat_groupnorm_workspace_floats(int n, int HW, int G) { return (size_t)n * gn_chunks(n, HW) * G * 2 + (size_t)n * G * 2; }

extern "C" int comat_groupnorm_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mean_rstd, float* ws,
                                   int n, int HW, int C, int G, float eps, int silu, int dtype, void* stream) {
  if (!x || !y || !gamma || !beta || !mean_rstd || !ws || C % 8 || C % G || C / 8 > 512) return COMAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (gn_use_fused()) {
    const int rc = gn_fused_launch(0, x, nullptr, y, gamma, beta, mean_rstd, n, HW, C, G, eps, silu, dtype, st);
    if (rc != COMAT_ERR_UNSUPPORTED) return rc;
  }
  const int chunks = gn_chunks(n, HW);
  const int thr = gn_threads(C), achunks = gn_apply_chunks(n, HW, thr / (C / 8));
  const float inv_cnt = 1.f / ((float)HW * (C / G));
  DISPATCH_T(dtype, {
    if (gn_unroll() >= 8) {
      launch_k(gn_partial_kernel<T, 0, 8>, dim3(chunks, n), thr, 2 * C * sizeof(float), st, (const T*)x, nullptr, gamma, beta, nullptr, ws, HW, C, G, chunks, 0);
      launch_k(gn_apply_kernel<T, 0, 8>, dim3(achunks, n), thr, 2 * G * sizeof(float), st, (const T*)x, nullptr, (T*)y, gamma, beta, ws, mean_rstd, HW, C, G, chunks, achunks, inv_cnt, eps, silu);
    } else {
      launch_k(gn_partial_kernel<T, 0, 4>, dim3(chunks, n), thr, 2 * C * sizeof(float), st, (const T*)x, nullptr, gamma, beta, nullptr, ws, HW, C, G, chunks, 0);
      launch_k(gn_apply_kernel<T, 0, 4>, dim3(achunks, n), thr, 2 * G * sizeof(float), st, (const T*)x, nullptr, (T*)y, gamma, beta, ws, mean_rstd, HW, C, G, chunks, achunks, inv_cnt, eps, silu);
    }
  });
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

// Statistics already accumulated by the producing GEMM's epilogue (comat_gemm_params.gn_sums: (sum, sum of squares) per
// (image, group) = the chunk-partial layout with ONE chunk): only the apply pass runs - x is read once and y written once.
extern "C" int comat_groupnorm_fwd_from_sums(const void* x, void* y, const float* gamma, const float* beta, float* mean_rstd,
                                             const float* sums, int n, int HW, int C, int G, float eps, int silu, int dtype,
                                             void* stream) {
  if (!x || !y || !gamma || !beta || !mean_rstd || !sums || C % 8 || C % G || C / 8 > 512) return COMAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int thr = gn_threads(C), achunks = gn_apply_chunks(n, HW, thr / (C / 8));
  const float inv_cnt = 1.f / ((float)HW * (C / G));
  DISPATCH_T(dtype, {
    if (gn_unroll() >= 8)
      launch_k(gn_apply_kernel<T, 0, 8>, dim3(achunks, n), thr, 2 * G * sizeof(float), st, (const T*)x, nullptr, (T*)y, gamma, beta, sums, mean_rstd, HW, C, G, 1, achunks, inv_cnt, eps, silu);
    else
      launch_k(gn_apply_kernel<T, 0, 4>, dim3(achunks, n), thr, 2 * G * sizeof(float), st, (const T*)x, nullptr, (T*)y, gamma, beta, sums, mean_rstd, HW, C, G, 1, achunks, inv_cnt, eps, silu);
  });
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

extern "C" int comat_groupnorm_bwd(const void* x, const void* dy, void* dx, const float* gamma, const float* beta,
                                   const float* mean_rstd, float* ws, int n, int HW, int C, int G, int silu, int dtype, void* stream) {
  if (!x || !dy || !dx || !gamma || !beta || !mean_rstd || !ws || C % 8 || C % G || C / 8 > 512) return COMAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (gn_use_fused()) {
    const int rc = gn_fused_launch(1, x, dy, dx, gamma, beta, const_cast<float*>(mean_rstd), n, HW, C, G, 0.f, silu, dtype, st);
    if (rc != COMAT_ERR_UNSUPPORTED) return rc;
  }
  const int chunks = gn_chunks(n, HW);
  const int thr = gn_threads(C), achunks = gn_apply_chunks(n, HW, thr / (C / 8));
  const float inv_cnt = 1.f / ((float)HW * (C / G));
  DISPATCH_T(dtype, {
    launch_k(gn_partial_kernel<T, 1, 4>, dim3(chunks, n), thr, 2 * C * sizeof(float), st, (const T*)x, (const T*)dy, gamma, beta, mean_rstd, ws, HW, C, G, chunks, silu);
    launch_k(gn_apply_kernel<T, 1, 4>, dim3(achunks, n), thr, 2 * G * sizeof(float), st, (const T*)x, (const T*)dy, (T*)dx, gamma, beta, ws, const_cast<float*>(mean_rstd), HW, C, G, chunks, achunks, inv_cnt, 0.f, silu);
  });
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

// vectors per lane / rows per warp by channel count: C <= 512 -> 2 x 4 rows, <= 768 -> 3 x 2, <= 1280 -> 5 x 2, <= 2048 -> 8 x 1
#define LN_CASE(MODE, NV, R, ...)                                                                                            \
  DISPATCH_T(dtype, (launch_k(ln_kernel<T, MODE, NV, R>, (unsigned)((rows + 8 * R - 1) / (8 * R)), 256, 0, (cudaStream_t)stream, __VA_ARGS__)))
#define LN_LAUNCH(MODE, ...)                                    \
  do {                                                          \
    if (C <= 512) { LN_CASE(MODE, 2, 4, __VA_ARGS__); }         \
    else if (C <= 768) { LN_CASE(MODE, 3, 2, __VA_ARGS__); }    \
    else if (C <= 1280) { LN_CASE(MODE, 5, 2, __VA_ARGS__); }   \
    else { LN_CASE(MODE, 8, 1, __VA_ARGS__); }                  \
  } while (0)

extern "C" int comat_layernorm_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mean_rstd, long long rows,
                                   int C, float eps, int dtype, void* stream) {
  if (!x || !y || !gamma || !beta || !mean_rstd || C % 8 || C > 2048) return COMAT_ERR_INVALID;
  LN_LAUNCH(0, (const T*)x, nullptr, (T*)y, gamma, beta, mean_rstd, rows, C, eps);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
extern "C" int comat_layernorm_bwd(const void* x, const void* dy, void* dx, const float* gamma, const float* mean_rstd, long long rows,
                                   int C, int dtype, void* stream) {
  if (!x || !dy || !dx || !gamma || !mean_rstd || C % 8 || C > 2048) return COMAT_ERR_INVALID;
  LN_LAUNCH(1, (const T*)x, (const T*)dy, (T*)dx, gamma, nullptr, const_cast<float*>(mean_rstd), rows, C, 0.f);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

extern "C" int comat_geglu_fwd(const void* hg, void* out, long long rows, int Ch, int dtype, void* stream) {
  if (!hg || !out || Ch % 8) return COMAT_ERR_INVALID;
  const long long nv = rows * (Ch / 8);
  DISPATCH_T(dtype, (launch_k(geglu_kernel<T, 0>, (unsigned)((nv + 255) / 256), 256, 0, (cudaStream_t)stream, (const T*)hg, nullptr, (T*)out, rows, Ch)));
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
extern "C" int comat_geglu_bwd(const void* hg, const void* dy, void* dhg, long long rows, int Ch, int dtype, void* stream) {
  if (!hg || !dy || !dhg || Ch % 8) return COMAT_ERR_INVALID;
  const long long nv = rows * (Ch / 8);
  DISPATCH_T(dtype, (launch_k(geglu_kernel<T, 1>, (unsigned)((nv + 255) / 256), 256, 0, (cudaStream_t)stream, (const T*)hg, (const T*)dy, (T*)dhg, rows, Ch)));
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

extern "C" int comat_elementwise(const void* x, const void* y, void* out, long long numel, int op, float alpha, float beta, int dtype,
                                 void* stream) {
  if (!x || !out || numel % 8) return COMAT_ERR_INVALID;
  const long long nv = numel / 8;
  DISPATCH_T(dtype, (launch_k(ew_kernel<T>, (unsigned)((nv + 255) / 256), 256, 0, (cudaStream_t)stream, (const T*)x, (const T*)y, (T*)out, nv, op, alpha, beta)));
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

extern "C" int comat_spatial(const void* in, void* out, int n, int H, int W, int C, int mode, int dtype, void* stream) {
  if (!in || !out || C % 8 || mode < 0 || mode > 3) return COMAT_ERR_INVALID;
  if ((mode >= 2) && ((H | W) & 1)) return COMAT_ERR_INVALID;
  const long long total = (long long)n * H * W * (C / 8) * (mode == 0 ? 4 : 1);
  DISPATCH_T(dtype, (launch_k(spatial_kernel<T>, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, (const T*)in, (T*)out, n, H, W, C, mode)));
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

extern "C" int comat_transpose16(const void* in, void* out, int R, int Cc, int ld_out, void* stream) {
  if (!in || !out || ld_out < R) return COMAT_ERR_INVALID;
  launch_k(transpose_kernel<__half>, dim3((Cc + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream, (const __half*)in, (__half*)out, R, Cc, ld_out);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

extern "C" int comat_latent_to_nhwc(const float* in, void* out, int n, int Cin, int HW, int Cpad, float scale, int dtype, void* stream) {
  if (!in || !out) return COMAT_ERR_INVALID;
  const long long total = (long long)n * HW * Cpad;
  DISPATCH_T(dtype, (launch_k(latent_to_nhwc_kernel<T>, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, in, (T*)out, n, Cin, HW, Cpad, scale)));
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
extern "C" int comat_nhwc_to_nchw_f32(const void* in, float* out, int n, int Cout, int HW, int ld, float scale, int dtype, void* stream) {
  if (!in || !out) return COMAT_ERR_INVALID;
  const long long total = (long long)n * Cout * HW;
  DISPATCH_T(dtype, (launch_k(nhwc_to_nchw_f32_kernel<T>, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, (const T*)in, out, n, Cout, HW, ld, scale)));
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

extern "C" int comat_copy2d16(const void* src, void* dst, long long rows, int cols, long long ld_src, long long ld_dst, void* stream) {
  if (!src || !dst || cols % 8 || ld_src % 8 || ld_dst % 8 || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return COMAT_ERR_INVALID;
  const long long nv = rows * (cols / 8);
  launch_k(copy2d_kernel, (unsigned)((nv + 255) / 256), 256, 0, (cudaStream_t)stream, (const uint4*)src, (uint4*)dst, rows, cols / 8, ld_src / 8, ld_dst / 8);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

// ============================================================================================== row softmax (unfused attention)
// Used for attention layers whose head dim exceeds the fused kernel's TMEM budget (the VAE decoder's single-head d=512
// mid-block attention over 4096 tokens): S = Q K^T and O = P V run on the tcgen05 GEMM, this kernel is the softmax between.
namespace comat {
template <typename T, int MODE>   // 0: p = softmax(x) ; 1: ds = p * (dp - sum(p*dp))
__global__ void __launch_bounds__(256) softmax_rows_kernel(const T* __restrict__ x, const T* __restrict__ dp, T* __restrict__ out, int Ccols) {
  pdl_grid_dependency_sync();
  __shared__ float sm[8];
  const size_t row = blockIdx.x;
  const T* xr = x + row * Ccols;
  const int nv = Ccols / 8;
  constexpr int MAXV = 4;                       // up to 8192 columns
  Vec8<T> xv[MAXV], dv[MAXV];
  float m = -INFINITY, s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int v = threadIdx.x + k * 256;
    if (v < nv) {
      xv[k].load(xr + v * 8);
      if (MODE == 1) dv[k].load(dp + row * Ccols + v * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) m = fmaxf(m, xv[k].get(i));
        else s += xv[k].get(i) * dv[k].get(i);
      }
    }
  }
  auto block_sum = [&](float v, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sm[0];
    for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, sm[i]) : r + sm[i];
    return r;
  };
  if (MODE == 0) {
    m = block_sum(m, true);
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int v = threadIdx.x + k * 256;
      if (v < nv) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s += __expf(xv[k].get(i) - m);
      }
    }
    s = block_sum(s, false);
    const float inv = 1.f / s;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int v = threadIdx.x + k * 256;
      if (v < nv) {
        Vec8<T> o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.set(i, __expf(xv[k].get(i) - m) * inv);
        o.store(out + row * Ccols + v * 8);
      }
    }
  } else {
    s = block_sum(s, false);
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int v = threadIdx.x + k * 256;
      if (v < nv) {
        Vec8<T> o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.set(i, xv[k].get(i) * (dv[k].get(i) - s));
        o.store(out + row * Ccols + v * 8);
      }
    }
  }
}
}  // namespace comat

extern "C" int comat_softmax_rows(const void* x, const void* dp, void* out, long long rows, int cols, int mode, int dtype, void* stream) {
  if (!x || !out || cols % 8 || cols > 8192 || (mode == 1 && !dp)) return COMAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == COMAT_F16) {
    if (mode == 0) launch_k(comat::softmax_rows_kernel<__half, 0>, (unsigned)rows, 256, 0, st, (const __half*)x, nullptr, (__half*)out, cols);
    else launch_k(comat::softmax_rows_kernel<__half, 1>, (unsigned)rows, 256, 0, st, (const __half*)x, (const __half*)dp, (__half*)out, cols);
  } else if (dtype == COMAT_BF16) {
    if (mode == 0) launch_k(comat::softmax_rows_kernel<__nv_bfloat16, 0>, (unsigned)rows, 256, 0, st, (const __nv_bfloat16*)x, nullptr, (__nv_bfloat16*)out, cols);
    else launch_k(comat::softmax_rows_kernel<__nv_bfloat16, 1>, (unsigned)rows, 256, 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dp, (__nv_bfloat16*)out, cols);
  } else return COMAT_ERR_UNSUPPORTED;
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

// ============================================================================================== fp32 -> bf16 hi/lo split
namespace comat {
__global__ void __launch_bounds__(256) split_f32_bf16x2_kernel(const float4* __restrict__ src, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                                               long long nv, float alpha) {
  pdl_grid_dependency_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    const float f[4] = {v.x * alpha, v.y * alpha, v.z * alpha, v.w * alpha};
    uint16_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat16 hb = __float2bfloat16_rn(f[k]);
      const __nv_bfloat16 lb = __float2bfloat16_rn(f[k] - __bfloat162float(hb));
      h[k] = *reinterpret_cast<const uint16_t*>(&hb);
      l[k] = *reinterpret_cast<const uint16_t*>(&lb);
    }
    hi[i] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    lo[i] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
  }
}
}  // namespace comat

extern "C" int comat_split_f32_bf16x2(const float* src, void* hi, void* lo, long long n, float alpha, void* stream) {
  if (!src || !hi || !lo || n <= 0 || (n % 4) != 0) return COMAT_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(hi) & 7) || (reinterpret_cast<uintptr_t>(lo) & 7)) return COMAT_ERR_INVALID;
  const long long nv = n / 4;
  long long blocks = (nv + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(comat::split_f32_bf16x2_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, (const float4*)src, (uint2*)hi, (uint2*)lo, nv, alpha);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
