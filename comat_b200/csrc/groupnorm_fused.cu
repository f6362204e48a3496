// Single-launch GroupNorm (+SiLU) forward / backward: the activation is read from HBM ONCE.
//
// Replaces aten group_norm + silu (diffusers ResnetBlock2D.norm1/norm2, Transformer2DModel.norm, conv_norm_out, the VAE
// decoder's norms) and their autograd.  The two-pass kernels in elementwise.cu read x twice (statistics, then apply) in two
// launches: 58 ms of a 527 ms step, 28 % of the HBM rate (profiles/r01_step_launches_v17.md).  Here a thread-block CLUSTER owns
// one (sample, channel-set) slab: every CTA of the cluster streams its rows into shared memory with 16-byte cp.async copies,
// reduces its per-group partial sums, the CTAs exchange the partials through distributed shared memory (fixed rank order), and
// each CTA then normalises its rows straight out of shared memory.  Slabs that do not fit keep what fits and re-read the rest
// (L2-resident in practice), so one kernel covers every shape.
//
//   channel set = lcm(channels per group, 8) channels: whole groups AND whole 16-byte vectors, e.g. C = 320, G = 32 -> 40
//   channels = 4 groups = 5 vectors per row.  grid = (cluster size, sets, samples), cluster = (CL, 1, 1).
#include "common.cuh"

namespace comat {

struct GnFP {
  const void* x;
  const void* dy;
  void* out;
  const float* gamma;
  const float* beta;
  float* mean_rstd;        // (n, G, 2): written by the forward, read by the backward
  int HW, C, G, cpg, SW, VW, gps;
  int keep_rows;           // rows of a CTA's slice that stay in shared memory
  float inv_cnt, eps;
  int silu;
};

__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* p, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <typename T>
struct V8 {
  uint4 u;
  __device__ __forceinline__ float get(int i) const { return to_f32<T>(reinterpret_cast<const T*>(&u)[i]); }
  __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<T*>(&u)[i] = from_f32<T>(v); }
};

__device__ __forceinline__ float gnf_silu(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float gnf_silu_grad(float x) {
  const float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}

template <typename T, int MODE>   // MODE 0: y = [silu](xhat * gamma + beta), saves (mean, rstd) ; MODE 1: dx
__global__ void __launch_bounds__(256) gn_fused_kernel(const GnFP p) {
  pdl_grid_dependency_sync();
  extern __shared__ __align__(16) unsigned char gsm[];
  const int CL = gridDim.x, rank = blockIdx.x, set = blockIdx.y, n = blockIdx.z;
  const int VW = p.VW, SW = p.SW, gps = p.gps, cpg = p.cpg, C = p.C, HW = p.HW;
  // warp-aligned mapping: a warp covers RPW = 32 / VW consecutive rows per step, lane = (row_sub, vector); the lanes that share a
  // vector column are VW apart, so per-channel sums reduce with a few shuffles and ONE plain store per (warp, channel) - no
  // shared-memory atomics (the first version added 16 partials per thread onto 2 * SW addresses: ~100-way CAS contention made
  // the kernel 2x slower than the two-pass pair, profiles/r02_groupnorm_bench.md)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  const int RPW = 32 / VW;
  const int v = lane % VW, rs = lane / VW;
  const bool active = rs < RPW;
  const int pr = warp * RPW + rs, RP = NW * RPW;       // first row of this thread, row step
  const int p0 = (int)((long long)HW * rank / CL), p1 = (int)((long long)HW * (rank + 1) / CL);
  const int R = p1 - p0;
  const int keep = R < p.keep_rows ? R : p.keep_rows;

  float* s_ch = reinterpret_cast<float*>(gsm);                 // [NW][2 * SW] per-warp per-channel sums
  float* s_part = s_ch + (size_t)NW * 2 * SW;                  // [2 * gps] per-group partials (read by the peers)
  float* s_stat = s_part + 2 * gps;                            // [2 * gps] cluster-wide statistics
  const int stat_bytes = ((NW * 2 * SW + 4 * gps) * 4 + 15) & ~15;
  uint4* s_x = reinterpret_cast<uint4*>(gsm + stat_bytes);     // [keep_rows][VW]
  uint4* s_d = s_x + (size_t)p.keep_rows * VW;                 // [keep_rows][VW]  (MODE 1)

  const size_t base = ((size_t)n * HW + p0) * C + (size_t)set * SW + v * 8;
  const T* xg = reinterpret_cast<const T*>(p.x) + base;
  const T* dg = (MODE == 1) ? reinterpret_cast<const T*>(p.dy) + base : nullptr;
  T* og = reinterpret_cast<T*>(p.out) + base;

  // ---- phase 1a: this CTA's rows -> shared memory, all copies in flight at once
  if (active) {
    for (int r = pr; r < keep; r += RP) {
      cp_async16(&s_x[(size_t)r * VW + v], xg + (size_t)r * C);
      if (MODE == 1) cp_async16(&s_d[(size_t)r * VW + v], dg + (size_t)r * C);
    }
  }

  float g8[8], b8[8], rs8[8], mu8[8];
  if (MODE == 1 && active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = set * SW + v * 8 + i, gi = c / cpg;
      g8[i] = p.gamma[c]; b8[i] = p.beta[c];
      rs8[i] = p.mean_rstd[((size_t)n * p.G + gi) * 2 + 1];
      mu8[i] = -p.mean_rstd[((size_t)n * p.G + gi) * 2] * rs8[i];          // xhat = x * rs + mu
    }
  }
  float sa[8], sb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sa[i] = sb[i] = 0.f;

  auto accum = [&](const V8<T>& xv, const V8<T>& dv) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float f = xv.get(i); sa[i] += f; sb[i] = fmaf(f, f, sb[i]); }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = fmaf(xv.get(i), rs8[i], mu8[i]);
        float d = dv.get(i);
        if (p.silu) d *= gnf_silu_grad(fmaf(xh, g8[i], b8[i]));
        const float a = d * g8[i];
        sa[i] += a; sb[i] = fmaf(a, xh, sb[i]);
      }
    }
  };
  // ---- phase 1b: rows that do not fit in shared memory are summed straight from global memory
  if (active) {
    for (int r = keep + pr; r < R; r += RP) {
      V8<T> xv, dv;
      xv.u = *reinterpret_cast<const uint4*>(xg + (size_t)r * C);
      if (MODE == 1) dv.u = *reinterpret_cast<const uint4*>(dg + (size_t)r * C);
      accum(xv, dv);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  if (active) {
    for (int r = pr; r < keep; r += RP) {               // each thread reads back exactly the vectors it copied
      V8<T> xv, dv;
      xv.u = s_x[(size_t)r * VW + v];
      if (MODE == 1) dv.u = s_d[(size_t)r * VW + v];
      accum(xv, dv);
    }
  }
  // lanes v, v + VW, v + 2 VW ... hold the same channels: fold them onto lane v (inactive lanes carry zeros)
  const bool pow2 = (32 % VW) == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a = sa[i], b = sb[i];
    if (pow2) {
      for (int off = 16; off >= VW; off >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); }
    } else {
      for (int off = VW; off < 32; off += VW) {
        const float a2 = __shfl_down_sync(0xffffffffu, sa[i], off), b2 = __shfl_down_sync(0xffffffffu, sb[i], off);
        if (lane + off < RPW * VW) { a += a2; b += b2; }
      }
    }
    if (lane < VW) { s_ch[(size_t)warp * 2 * SW + v * 8 + i] = a; s_ch[(size_t)warp * 2 * SW + SW + v * 8 + i] = b; }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * SW; t += blockDim.x) {        // over the warps, fixed order; a thread touches its own column only
    float acc = 0.f;
    for (int w = 0; w < NW; ++w) acc += s_ch[(size_t)w * 2 * SW + t];
    s_ch[t] = acc;
  }
  __syncthreads();
  for (int g = warp; g < gps; g += NW) {                          // channels of a group: one warp, shuffle tree (deterministic)
    float a = 0.f, b = 0.f;
    for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) { a += s_ch[c]; b += s_ch[SW + c]; }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { s_part[2 * g] = a; s_part[2 * g + 1] = b; }
  }
  cl_sync();                                             // every CTA's partials are published
  if (threadIdx.x < gps) {
    float a = 0.f, b = 0.f;
    for (int rk = 0; rk < CL; ++rk) {                    // fixed order: identical statistics in every CTA of the cluster
      a += ld_dsmem_f32(&s_part[2 * threadIdx.x], rk);
      b += ld_dsmem_f32(&s_part[2 * threadIdx.x + 1], rk);
    }
    const int gi = set * gps + threadIdx.x;
    if (MODE == 0) {
      const float mean = a * p.inv_cnt;
      const float var = fmaxf(b * p.inv_cnt - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      s_stat[2 * threadIdx.x] = mean; s_stat[2 * threadIdx.x + 1] = rstd;
      if (rank == 0) { p.mean_rstd[((size_t)n * p.G + gi) * 2] = mean; p.mean_rstd[((size_t)n * p.G + gi) * 2 + 1] = rstd; }
    } else {
      s_stat[2 * threadIdx.x] = a * p.inv_cnt; s_stat[2 * threadIdx.x + 1] = b * p.inv_cnt;
    }
  }
  cl_sync();                                             // peers are done reading this CTA's partials; s_stat is visible
  if (!active) return;

  // ---- phase 2: normalise (or dx), from shared memory where the rows were kept
  float A[8], B[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int cl = v * 8 + i, c = set * SW + cl, gl = cl / cpg;
    if (MODE == 0) {
      A[i] = s_stat[2 * gl + 1] * p.gamma[c];
      B[i] = fmaf(-s_stat[2 * gl], A[i], p.beta[c]);
    } else {
      m1[i] = s_stat[2 * gl]; m2[i] = s_stat[2 * gl + 1];
    }
  }
  auto apply = [&](const V8<T>& xv, const V8<T>& dv, T* dst) {
    V8<T> o;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        float y = fmaf(xv.get(i), A[i], B[i]);
        if (p.silu) y = gnf_silu(y);
        o.set(i, y);
      } else {
        const float xh = fmaf(xv.get(i), rs8[i], mu8[i]);
        float d = dv.get(i);
        if (p.silu) d *= gnf_silu_grad(fmaf(xh, g8[i], b8[i]));
        const float a = d * g8[i];
        o.set(i, rs8[i] * (a - m1[i] - xh * m2[i]));
      }
    }
    *reinterpret_cast<uint4*>(dst) = o.u;
  };
  for (int r = pr; r < keep; r += RP) {
    V8<T> xv, dv;
    xv.u = s_x[(size_t)r * VW + v];
    if (MODE == 1) dv.u = s_d[(size_t)r * VW + v];
    apply(xv, dv, og + (size_t)r * C);
  }
  for (int r = keep + pr; r < R; r += RP) {
    V8<T> xv, dv;
    xv.u = *reinterpret_cast<const uint4*>(xg + (size_t)r * C);
    if (MODE == 1) dv.u = *reinterpret_cast<const uint4*>(dg + (size_t)r * C);
    apply(xv, dv, og + (size_t)r * C);
  }
}

static int gcd_i(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

template <typename T, int MODE>
static int gn_fused_launch_t(const GnFP& p0, int n, cudaStream_t st) {
  GnFP p = p0;
  static bool configured = false;
  if (!configured) {
    COMAT_CUDA(cudaFuncSetAttribute(gn_fused_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  const int nsets = p.C / p.SW;
  const int row_bytes = p.VW * 16 * (MODE == 1 ? 2 : 1);
  int CL = 1;
  while (CL < 8 && p.HW / (2 * CL) >= 16 &&
         (((long long)(p.HW + CL - 1) / CL) * row_bytes > 48 * 1024 || (long long)n * nsets * CL < 2LL * num_sms()))
    CL *= 2;
  const int R = (p.HW + CL - 1) / CL;
  const int RPW = 32 / p.VW;
  int NW = (R + RPW - 1) / RPW;                          // one row per thread is enough for the smallest slabs
  if (NW > 8) NW = 8;                                    // 256 threads: up to 4 CTAs per SM, their load / reduce / store phases overlap
  if (NW < 2) NW = 2;
  const int thr = NW * 32;
  const int stat_bytes = ((NW * 2 * p.SW + 4 * p.gps) * 4 + 15) & ~15;
  const long long slab = (long long)R * row_bytes;
  const long long cap = slab + stat_bytes <= 200 * 1024 ? slab : 96 * 1024;      // fits: keep everything (1 CTA / SM if large)
  p.keep_rows = (int)(cap / row_bytes);
  if (p.keep_rows > R) p.keep_rows = R;
  const size_t smem = (size_t)stat_bytes + (size_t)p.keep_rows * row_bytes;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(CL, nsets, n); cfg.blockDim = dim3(thr); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = comat_pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, gn_fused_kernel<T, MODE>, p);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

// returns COMAT_ERR_UNSUPPORTED when the geometry has no whole-vector channel set (caller falls back to the two-pass kernels)
int gn_fused_launch(int mode, const void* x, const void* dy, void* out, const float* gamma, const float* beta, float* mean_rstd,
                    int n, int HW, int C, int G, float eps, int silu, int dtype, cudaStream_t st) {
  if (C % 8 || C % G || n > 65535) return COMAT_ERR_UNSUPPORTED;
  const int cpg = C / G;
  const int SW = cpg / gcd_i(cpg, 8) * 8;               // lcm(cpg, 8)
  if (SW > 512 || C % SW) return COMAT_ERR_UNSUPPORTED;
  if (C / SW > 65535) return COMAT_ERR_UNSUPPORTED;
  GnFP p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.dy = dy; p.out = out; p.gamma = gamma; p.beta = beta; p.mean_rstd = mean_rstd;
  p.HW = HW; p.C = C; p.G = G; p.cpg = cpg; p.SW = SW; p.VW = SW / 8; p.gps = SW / cpg;
  p.inv_cnt = 1.f / ((float)HW * cpg); p.eps = eps; p.silu = silu;
  if (dtype == COMAT_F16) return mode == 0 ? gn_fused_launch_t<__half, 0>(p, n, st) : gn_fused_launch_t<__half, 1>(p, n, st);
  if (dtype == COMAT_BF16) return mode == 0 ? gn_fused_launch_t<__nv_bfloat16, 0>(p, n, st) : gn_fused_launch_t<__nv_bfloat16, 1>(p, n, st);
  return COMAT_ERR_UNSUPPORTED;
}

}  // namespace comat
