// Tensor-core GEMM / implicit-GEMM convolution for the UNet / VAE / BLIP hot path  (tcgen05 + TMEM + TMA).
//
//   out[m, n] = act(alpha * sum_k A[m, k] * B[n, k] + bias[n] + rowvec[m / rows_per_group, n]) + residual[m, n]
//
// replaces the cuBLAS / cuDNN calls behind diffusers' Linear / Conv2d / LoRACompatibleLinear modules
// (ResBlock conv3x3, proj_in/out, attention projections + LoRA, GEGLU feed-forward, BLIP linears).
//
// * A is K-major 16-bit.  "plain" mode: a [M, K] matrix.  "conv" mode: an NHWC activation tensor read through a 4-D
//   TMA tensor map — one (tap, 64-channel) slab per k-block at spatial offset (dh, dw); out-of-bounds rows are
//   zero-filled by the TMA unit, which *is* the convolution's zero padding, so no im2col buffer ever exists.
// * Up to two K-segments accumulate into the same TMEM accumulator: [x | x*down^T] against [W | up] fuses the LoRA
//   branch (training_utils/pipeline.py:94-115) into the projection, and [hidden | skip] fuses the UNet's
//   torch.cat([hidden, skip], dim=1) into the ResBlock's first convolution.
// * B is the weight, [N, K_total] K-major (for conv: k = tap * c_total + channel).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
// warps 2-5 = epilogue (tcgen05.ld -> registers -> fused bias / time-embedding / activation / residual -> global).
// One 128 x BN output tile per CTA, 3-stage smem ring; two CTAs fit per SM so one CTA's epilogue overlaps the
// other's main loop.
#include "tc_common.cuh"

namespace comat {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 x 16-bit = 128 B = one swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int PERSIST_EPI_WARPS = 8;                       // persistent kernel: 2 epilogue warps per TMEM lane quarter
constexpr int PERSIST_THREADS = 64 + 32 * PERSIST_EPI_WARPS;

struct GemmKP {
  int M, N;
  int n_seg, seg_kblocks[2], seg_bkoff[2];
  int conv, H, W, n_img, TW, TH, TN, tiles_w, tiles_h, n_taps, c_total;
  int dh[9], dw[9];
  float alpha;
  const float* bias;
  const float* rowvec;
  long long rowvec_ld;
  int rows_per_group, act;
  const void* residual;
  long long res_ld;
  void* out16;
  long long out_ld;
  float* out32;
  long long out32_ld;
  uint32_t idesc;
  int is_bf16;
  int a_mn, b_mn;              // operand stored MN-major ([K, M] / [K, N] row-major): TMA panels + MN-major descriptors
  int tiles_m;                 // number of 128-row (or 128-pixel) output tiles
  int split_k, kb_per_split;   // split-K: blockIdx.z handles k-blocks [z*kb_per_split, ...) and writes raw fp32 partials
  float* splitk_ws;            // [split_k][M][N] fp32
  int tma_res;       // 1 (persistent / pair kernels, lean flavours): the residual tile arrives in the staging tile through TMA
  int tma_store;     // 1: epilogue stages 16-bit tiles in (free) pipeline smem and writes them with TMA bulk tensor stores
  int atomic_acc;    // 1: out32 += alpha * (this CTA's partial sum) with vector fp32 atomics - gradient accumulation across
                     //    K-splits AND across calls in one kernel (no partial workspace, no reduction pass)
  // GroupNorm statistics of the OUTPUT tensor, accumulated by the epilogue (the consumer's GroupNorm then needs no statistics
  // pass over HBM): gn_sums[(image * gn_G + group) * 2 + {0, 1}] += {sum, sum of squares} of this tile's fp32 results
  float* gn_sums;
  int gn_G, gn_cpg, gn_rows_per_img, gn_n_img;
};

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == 1) return x / (1.f + __expf(-x));                       // SiLU
  if (act == 2) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));  // exact GELU (F.gelu default)
  return x;
}

__device__ __forceinline__ uint32_t pack16(float a, float b, int is_bf16) {
  if (is_bf16) { const __nv_bfloat162 t = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&t); }
  const __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}

// GroupNorm statistics of one 32-row x 32-column piece of the output (one epilogue warp, one chunk).  f[] = this lane's row.
// Transposing butterfly: 31 shuffles turn "lane = row, register = column" into "lane = column" with the 32 rows summed, then a
// segmented suffix sum over the lanes of one channel group (groups are runs of gn_cpg consecutive columns) leaves the group
// totals in each run's first lane, which adds them to the (image, group) accumulators with one fp32 atomic each.
// Must be called by all 32 lanes (TMA-store epilogues only: every thread walks every chunk).
__device__ __forceinline__ float gn_transpose_sum(float (&s)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int k = 0; k < w; ++k) {
      const float send = up ? s[k] : s[k + w];
      const float keep = up ? s[k + w] : s[k];
      s[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return s[0];
}
// image that the 32 rows of TMEM lane quarter q of a tile belong to (the host only enables the statistics when a quarter cannot
// straddle two images: conv tiles with TW*TH >= 32 pixels per image, plain GEMMs with rows_per_image % 32 == 0)
__device__ __forceinline__ int gn_warp_image(const GemmKP& p, int m0, int img0, int q) {
  if (p.gn_sums == nullptr) return 0;
  return p.conv ? img0 + (q * 32) / (p.TW * p.TH) : (m0 + q * 32) / p.gn_rows_per_img;
}
__device__ __forceinline__ void gn_stats_chunk(const GemmKP& p, const float (&f)[32], int col0, int ncol, bool row_ok, int img) {
  const int lane = threadIdx.x & 31;
  float s[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j] = (row_ok && j < ncol) ? f[j] : 0.f;
  float S = gn_transpose_sum(s, lane);
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j] = (row_ok && j < ncol) ? f[j] * f[j] : 0.f;
  float Q = gn_transpose_sum(s, lane);
  const int c = col0 + lane;                               // this lane's output channel
  const int g = c / p.gn_cpg;
  const int last = min(31, (g + 1) * p.gn_cpg - 1 - col0);  // last lane of this lane's group inside the chunk
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float ts = __shfl_down_sync(0xffffffffu, S, d), tq = __shfl_down_sync(0xffffffffu, Q, d);
    if (lane + d <= last) { S += ts; Q += tq; }
  }
  const bool head = lane == 0 || (c % p.gn_cpg) == 0;
  if (head && lane < ncol && img < p.gn_n_img) {
    float* o = p.gn_sums + ((size_t)img * p.gn_G + g) * 2;
    atomicAdd(o, S);
    atomicAdd(o + 1, Q);
  }
}

// Residual prefetch: the 64 bytes of the residual row that a thread adds to one 32-column chunk, requested BEFORE the thread waits
// for the accumulator / while it works on the previous chunk.  Without it every chunk exposed one global-memory round trip, and
// the epilogue - not the MMA - paced the K <= 640 linears that carry a residual (out-projections, feed-forward outputs, every
// gradient accumulation `dx += ...`: ~3 us per 128 x 160 tile against 0.8 us of MMA, profiles/r02_gemm_shapes_v7.md).
__device__ __forceinline__ bool res_prefetch(const GemmKP& p, uint4 (&r)[4], long long m, bool row_ok, int n0, int c0) {
  if (p.residual == nullptr || !row_ok || p.split_k > 1 || p.atomic_acc || p.act == 3 || n0 + c0 + 32 > p.N) return false;
  const uint16_t* rp = reinterpret_cast<const uint16_t*>(p.residual) + m * p.res_ld + n0 + c0;
  if ((reinterpret_cast<uintptr_t>(rp) & 15) != 0) return false;
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) r[j4] = *reinterpret_cast<const uint4*>(rp + j4 * 8);
  return true;
}

// Epilogue flavours, chosen on the host and compiled into separate kernels.  One kernel used to carry all of them behind runtime
// branches: 6 200-6 700 instructions (~100 KB) per kernel, of which a call executes ~1 500.  Most launches of the step run 10-25 us
// and start with a cold instruction cache; ncu's top stall on the short-K linears was `no_instruction` (3.8 per issue,
// gpurun_out/r02_lin_32768_320_k64.ncu-rep) and a stripped epilogue was 10-25 % faster on them.  Compile-time flavours keep
// every kernel near its executed footprint.
enum : int {
  EPI_STD = 0,      // bias + row vector + SiLU / GELU + residual -> 16-bit (TMA-store staging or direct) and / or fp32
  EPI_GN = 1,       // EPI_LEAN + GroupNorm statistics of the output
  EPI_SPLITK = 2,   // raw fp32 partials of one K slice
  EPI_ATOMIC = 3,   // out32 += alpha * partial (vector fp32 atomics)
  EPI_GEGLU = 4,    // bias + hidden * gelu(gate) on interleaved columns -> half-width 16-bit
  EPI_LEAN = 5      // the common case of EPI_STD: no activation, 16-bit output through the TMA-store staging only
};

// One 32-column chunk of the epilogue for one accumulator row: v[] holds the raw fp32 accumulators of columns
// [n0+c0, n0+c0+32) of tile row r (global row m).  zsplit = split-K slice (raw partial store), stage = smem staging tile
// for the TMA-store path.  Every runtime option is tested once per chunk (uniform branches), the element loops are
// straight FFMA / pack code; bias comes from smem as float4 (the first version spent ~19 instructions per element here
// and was instruction-issue bound: profiles/r01_gemm_k320_ncu.md).
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmKP& p, const uint32_t (&v)[32], int c0, int n0, long long m, bool row_ok,
                                           const float* rv, const float* s_bias, int r, unsigned char* stage, int zsplit, int gn_img = 0,
                                           const uint4* rpre = nullptr, bool res_staged = false) {
  const int ncol = min(32, p.N - (n0 + c0));
  if constexpr (EPI == EPI_ATOMIC) {
    if (row_ok && ncol > 0) {
      float* op = p.out32 + m * p.out32_ld + n0 + c0;
      if (ncol == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          atomicAdd(reinterpret_cast<float4*>(op + j), make_float4(__uint_as_float(v[j]) * p.alpha, __uint_as_float(v[j + 1]) * p.alpha,
                                                                     __uint_as_float(v[j + 2]) * p.alpha, __uint_as_float(v[j + 3]) * p.alpha));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncol) atomicAdd(op + j, __uint_as_float(v[j]) * p.alpha);
      }
    }
  } else if constexpr (EPI == EPI_SPLITK) {
    if (row_ok && ncol > 0) {
      float* wp = p.splitk_ws + ((size_t)zsplit * p.M + m) * p.N + n0 + c0;
      if (ncol == 32 && ((reinterpret_cast<uintptr_t>(wp) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(wp + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncol) wp[j] = __uint_as_float(v[j]);
      }
    }
  } else if (EPI == EPI_LEAN || EPI == EPI_GN || p.tma_store || (row_ok && ncol > 0)) {
    float f[32];
    const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bb = b4[j / 4];
      f[j] = fmaf(__uint_as_float(v[j]), p.alpha, bb.x);
      f[j + 1] = fmaf(__uint_as_float(v[j + 1]), p.alpha, bb.y);
      f[j + 2] = fmaf(__uint_as_float(v[j + 2]), p.alpha, bb.z);
      f[j + 3] = fmaf(__uint_as_float(v[j + 3]), p.alpha, bb.w);
    }
    if constexpr (EPI == EPI_GEGLU) {
      // GEGLU (diffusers GEGLU.forward: hidden * gelu(gate)) on a projection whose weight rows were interleaved at pack time
      // (column 2j = hidden_j, 2j+1 = gate_j): 32 accumulator columns -> 16 outputs of the half-width tensor.  The (M, 2*inner)
      // pre-activation never reaches HBM and the separate GEGLU pass disappears (no-grad passes; taped passes keep hg).
      if (row_ok && ncol > 0) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float g0 = f[4 * j + 1], g1 = f[4 * j + 3];
          const float o0 = f[4 * j] * (0.5f * g0 * (1.f + erff(g0 * 0.70710678118654752f)));
          const float o1 = f[4 * j + 2] * (0.5f * g1 * (1.f + erff(g1 * 0.70710678118654752f)));
          pk[j] = pack16(o0, o1, p.is_bf16);
        }
        uint16_t* op = reinterpret_cast<uint16_t*>(p.out16) + m * p.out_ld + ((n0 + c0) >> 1);
        if (ncol == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
          *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(op + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (2 * j < ncol) op[j] = (uint16_t)((j & 1) ? (pk[j / 2] >> 16) : (pk[j / 2] & 0xFFFFu));
        }
      }
      return;
    } else {
    if (rv != nullptr) {
      const float* rvc = rv + n0 + c0;
      if (ncol == 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] += rvc[j];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncol) f[j] += rvc[j];
      }
    }
    if constexpr (EPI == EPI_STD) {
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = f[j] / (1.f + __expf(-f[j]));
      } else if (p.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = 0.5f * f[j] * (1.f + erff(f[j] * 0.70710678118654752f));
      }
    }
    if (p.residual != nullptr && (res_staged || row_ok) && ncol > 0) {
      const uint16_t* rp = reinterpret_cast<const uint16_t*>(p.residual) + m * p.res_ld + n0 + c0;
      if (res_staged || rpre != nullptr || (ncol == 32 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0))) {
        uint4 u[4];
        if (res_staged) {
          // the residual tile was bulk-loaded into the staging tile (same 64B-swizzled panels the result is written to): this
          // thread reads the 64 bytes it is about to overwrite
          const unsigned char* prow = stage + (c0 / 32) * 8192 + r * 64;
          const int sw = (r >> 1) & 3;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) u[j4] = *reinterpret_cast<const uint4*>(prow + ((j4 ^ sw) * 16));
        } else if (rpre != nullptr) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) u[j4] = rpre[j4];          // loaded while the accumulator was still being produced
        } else {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) u[j4] = *reinterpret_cast<const uint4*>(rp + j4 * 8);
        }
        if (p.is_bf16) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint32_t w[4] = {u[j4].x, u[j4].y, u[j4].z, u[j4].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              f[j4 * 8 + e * 2 + 0] += __uint_as_float(w[e] << 16);
              f[j4 * 8 + e * 2 + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
            }
          }
        } else {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint32_t w[4] = {u[j4].x, u[j4].y, u[j4].z, u[j4].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
              f[j4 * 8 + e * 2 + 0] += t.x;
              f[j4 * 8 + e * 2 + 1] += t.y;
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {          // static indexing keeps f[] in registers
          if (j < ncol) {
            const uint16_t bv = rp[j];
            f[j] += p.is_bf16 ? __uint_as_float((uint32_t)bv << 16) : __half2float(*reinterpret_cast<const __half*>(&bv));
          }
        }
      }
    }
    if constexpr (EPI == EPI_GN) {
      if (ncol > 0) gn_stats_chunk(p, f, n0 + c0, ncol, row_ok, gn_img);   // warp-uniform: the host only picks EPI_GN with tma_store
    }
    uint32_t pk[16];
    if (EPI != EPI_STD || p.out16 != nullptr) {
      if (p.is_bf16) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { const __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]); pk[j] = *reinterpret_cast<const uint32_t*>(&t); }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { const __half2 t = __floats2half2_rn(f[2 * j], f[2 * j + 1]); pk[j] = *reinterpret_cast<const uint32_t*>(&t); }
      }
    }
    if (EPI != EPI_STD || p.tma_store) {
      // panel (c0/32): [128 rows][64 B], 64B-swizzled (16B chunk q of row r lives at q ^ ((r>>1)&3))
      unsigned char* prow = stage + (c0 / 32) * 8192 + r * 64;
      const int sw = (r >> 1) & 3;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
        *reinterpret_cast<uint4*>(prow + ((q4 ^ sw) * 16)) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
    } else {
      if (p.out16 != nullptr) {
        uint16_t* op = reinterpret_cast<uint16_t*>(p.out16) + m * p.out_ld + n0 + c0;
        if (ncol == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<uint4*>(op + q4 * 8) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncol) op[j] = (uint16_t)((j & 1) ? (pk[j / 2] >> 16) : (pk[j / 2] & 0xFFFFu));
        }
      }
      if (p.out32 != nullptr && row_ok) {
        float* op = p.out32 + m * p.out32_ld + n0 + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncol) op[j] = f[j];
      }
    }
    }
  }
}

template <int BN>
constexpr int tmem_cols() { return BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512; }

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TILES = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = TILES;                       // full[STAGES], empty[STAGES], tmem_full
  static constexpr int TMEMPTR_OFF = BAR_OFF + (2 * STAGES + 1) * 8;
  static constexpr int BIAS_OFF = (TMEMPTR_OFF + 8 + 15) & ~15;     // float4-aligned
  static constexpr int TOTAL = BIAS_OFF + BN * 4 + 1024;      // +1024: manual alignment slack
};

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmO, const GemmKP p) {
  pdl_trigger();
  using S = GemmSmem<BN, STAGES>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::TMEMPTR_OFF);
  float* s_bias = reinterpret_cast<float*>(smem + S::BIAS_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;
  const int n0 = tile_n * BN;
  const int kb_per_tap = p.seg_kblocks[0] + (p.n_seg > 1 ? p.seg_kblocks[1] : 0);
  const int num_kb_total = p.n_taps * kb_per_tap;
  const int kb_begin = p.split_k > 1 ? blockIdx.z * p.kb_per_split : 0;
  const int kb_end = p.split_k > 1 ? min(num_kb_total, kb_begin + p.kb_per_split) : num_kb_total;
  const int num_kb = kb_end - kb_begin;

  // tile origin
  int m0 = tile_m * BM, img0 = 0, h0 = 0, w0 = 0;
  if (p.conv) {
    const int tw = tile_m % p.tiles_w, th = (tile_m / p.tiles_w) % p.tiles_h, tn = tile_m / (p.tiles_w * p.tiles_h);
    w0 = tw * p.TW; h0 = th * p.TH; img0 = tn * p.TN;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (p.n_seg > 1) { tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, tmem_cols<BN>());
    tmem_relinquish();
  }
  pdl_wait();                                    // everything above touched no global memory
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += GEMM_THREADS - 64)
      s_bias[i] = (p.bias != nullptr && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int kbi = kb_begin; kbi < kb_end; ++kbi) {
        const int tap = kbi / kb_per_tap;
        int rem = kbi - tap * kb_per_tap;
        const int sg = (rem >= p.seg_kblocks[0]) ? 1 : 0;
        const int cb = rem - (sg ? p.seg_kblocks[0] : 0);
        const CUtensorMap* mA = sg == 0 ? &tmA0 : &tmA1;
        const CUtensorMap* mB = sg == 0 ? &tmB0 : &tmB1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        unsigned char* sa = smem + stage * S::STAGE_BYTES;
        unsigned char* sb = sa + S::A_BYTES;
        mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
        if (p.conv) tma_load_4d(sa, mA, &full_bar[stage], cb * BK, w0 + p.dw[tap], h0 + p.dh[tap], img0);
        else if (p.a_mn) { tma_load_2d(sa, mA, &full_bar[stage], m0, cb * BK); tma_load_2d(sa + 8192, mA, &full_bar[stage], m0 + 64, cb * BK); }
        else        tma_load_2d(sa, mA, &full_bar[stage], cb * BK, m0);
        if (p.b_mn) {
#pragma unroll
          for (int pn = 0; pn < BN / 64; ++pn) tma_load_2d(sb + pn * 8192, mB, &full_bar[stage], n0 + pn * 64, cb * BK);
        } else tma_load_2d(sb, mB, &full_bar[stage], tap * p.c_total + p.seg_bkoff[sg] + cb * BK, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
        const uint32_t sb = sa + S::A_BYTES;
        const uint64_t da = p.a_mn ? make_mnmajor_sw128_desc(sa, 8192) : make_kmajor_sw128_desc(sa);
        const uint64_t db = p.b_mn ? make_mnmajor_sw128_desc(sb, 8192) : make_kmajor_sw128_desc(sb);
        const uint64_t ka = p.a_mn ? 128 : 2, kbs = p.b_mn ? 128 : 2;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 elements along K: K-major = 32 B inside the 128-B swizzle row (+2 in the addr>>4 field),
          // MN-major = 16 rows of 128 B (+128)
          umma_f16(tmem_base, da + ka * k, db + kbs * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);          // frees this smem stage when the MMAs above have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);                    // accumulator complete
    }
    __syncwarp();
  } else {
    // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // row inside the tile
    long long m;                                 // global output row
    bool row_ok;
    if (p.conv) {
      const int tw = r % p.TW, th = (r / p.TW) % p.TH, tn = r / (p.TW * p.TH);
      const int n_i = img0 + tn, hh = h0 + th, ww = w0 + tw;
      row_ok = (n_i < p.n_img) && (hh < p.H) && (ww < p.W);
      m = ((long long)n_i * p.H + hh) * p.W + ww;
    } else {
      m = (long long)m0 + r;
      row_ok = m < p.M;
    }
    const float* rv = (p.rowvec != nullptr && row_ok) ? p.rowvec + (m / p.rows_per_group) * p.rowvec_ld : nullptr;
    uint4 rcur[4], rnxt[4];
    bool have_nxt = (EPI <= EPI_GN || EPI == EPI_LEAN) && res_prefetch(p, rnxt, m, row_ok, n0, 0);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(trow + (uint32_t)c0, v);
      const bool have = have_nxt;
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) rcur[j4] = rnxt[j4];
      have_nxt = (EPI <= EPI_GN || EPI == EPI_LEAN) && (c0 + 32 < BN) && res_prefetch(p, rnxt, m, row_ok, n0, c0 + 32);
      tmem_ld_wait();
      epilogue_chunk<EPI>(p, v, c0, n0, m, row_ok, rv, s_bias, r, smem, blockIdx.z, EPI == EPI_GN ? gn_warp_image(p, m0, img0, q) : 0, have ? rcur : nullptr);
    }
    tc_fence_before();
    if (p.tma_store) {
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");       // the 4 epilogue warps only
      if (warp == 2 && lane == 0) {
#pragma unroll 1
        for (int pn = 0; pn < BN / 32; ++pn) {
          if (n0 + pn * 32 >= p.N) break;
          if (p.conv) tma_store_4d(&tmO, smem + pn * 8192, n0 + pn * 32, w0, h0, img0);
          else        tma_store_2d(&tmO, smem + pn * 8192, n0 + pn * 32, m0);
        }
        bulk_commit();
        bulk_wait_read<0>();                                 // smem must stay valid until the stores have read it
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols<BN>());
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant (default): one CTA per SM walks the (m-tile, n-tile, k-split) work list, n fastest so the CTAs
// running at the same time share A rows through L2.  The TMA producer runs ahead across tile boundaries (deeper smem
// ring than the one-tile kernel), the accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the
// main loop of tile i+1, and the 16-bit output goes through a dedicated staging tile + TMA bulk tensor stores.
// Measured motivation (profiles/r01_gemm_shapes_v3.md): K <= 640 shapes ran at 400-500 TFLOP/s in the one-tile kernel
// because prologue + first-load latency + epilogue were serialised per tile, and single-wave long-K shapes were
// latency-bound by a 3-stage ring.
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct PersistSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TILES = STAGES * STAGE_BYTES;
  static constexpr int STAGING_OFF = TILES;                         // [(BN+31)/32 panels][128 rows][64 B]
  static constexpr int STAGING_BYTES = ((BN + 31) / 32) * 8192;
  static constexpr int BAR_OFF = STAGING_OFF + STAGING_BYTES;       // full[STAGES], empty[STAGES], tfull[2], tempty[2]
  static constexpr int TMEMPTR_OFF = BAR_OFF + (2 * STAGES + 5) * 8;   // ... + res_full
  static constexpr int BIAS_OFF = (TMEMPTR_OFF + 8 + 15) & ~15;     // [2][BN] floats, float4-aligned
  static constexpr int TOTAL = BIAS_OFF + 2 * BN * 4 + 1024;        // +1024: manual alignment slack
};

struct TileCoord {
  int tile_n, z, m0, img0, h0, w0;
};
__device__ __forceinline__ TileCoord make_coord(const GemmKP& p, int tile_m, int tile_n, int z) {
  TileCoord c;
  c.z = z; c.tile_n = tile_n;
  c.m0 = tile_m * BM; c.img0 = 0; c.h0 = 0; c.w0 = 0;
  if (p.conv) {
    const int tw = tile_m % p.tiles_w, th = (tile_m / p.tiles_w) % p.tiles_h, tn = tile_m / (p.tiles_w * p.tiles_h);
    c.w0 = tw * p.TW; c.h0 = th * p.TH; c.img0 = tn * p.TN;
  }
  return c;
}
__device__ __forceinline__ TileCoord decode_tile(const GemmKP& p, int t, int tiles_n) {
  const int per_z = p.tiles_m * tiles_n;
  const int z = t / per_z;
  const int r = t - z * per_z;
  const int tile_m = r / tiles_n;
  return make_coord(p, tile_m, r - tile_m * tiles_n, z);
}
// CTA-pair kernel: pair-tile t covers m-tiles 2*pm and 2*pm+1 (this CTA takes 2*pm + rank; an odd tail tile is out of bounds:
// TMA zero-fills its loads and clips its stores), n fastest
__device__ __forceinline__ TileCoord decode_pair_tile(const GemmKP& p, int t, int tiles_n, int pairs_m, int rank) {
  const int per_z = pairs_m * tiles_n;
  const int z = t / per_z;
  const int r = t - z * per_z;
  const int pm = r / tiles_n;
  return make_coord(p, 2 * pm + rank, r - pm * tiles_n, z);
}

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                       const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const GemmKP p) {
  pdl_trigger();
  using S = PersistSmem<BN, STAGES>;
  constexpr int ACC = tmem_cols<BN>();           // TMEM columns per accumulator buffer
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull = empty_bar + STAGES;          // [2] accumulator ready for the epilogue
  uint64_t* tempty = tfull + 2;                  // [2] accumulator drained, MMA may overwrite
  uint64_t* res_full = tempty + 2;               // residual tile of the current output tile has landed in the staging tile
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::TMEMPTR_OFF);
  float* s_bias = reinterpret_cast<float*>(smem + S::BIAS_OFF);
  unsigned char* staging = smem + S::STAGING_OFF;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int total_tiles = p.tiles_m * tiles_n * p.split_k;
  const int kb_per_tap = p.seg_kblocks[0] + (p.n_seg > 1 ? p.seg_kblocks[1] : 0);
  const int num_kb_total = p.n_taps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (p.n_seg > 1) { tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); }
    if (p.tma_store) tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], PERSIST_EPI_WARPS); mbar_init(&tempty[1], PERSIST_EPI_WARPS);      // one arrival per epilogue warp
    mbar_init(res_full, 1);
    if (p.tma_res) tma_prefetch_desc(&tmR);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 2 * ACC);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                    // everything above touched no global memory

  if (warp == 0) {
    // ===================== TMA producer: streams k-blocks of tile after tile =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t, tiles_n);
        const int n0 = c.tile_n * BN;
        const int kb_begin = c.z * p.kb_per_split;
        const int kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
        int tap = kb_begin / kb_per_tap;
        int rem = kb_begin - tap * kb_per_tap;
        for (int kbi = kb_begin; kbi < kb_end; ++kbi) {
          const int sg = (rem >= p.seg_kblocks[0]) ? 1 : 0;
          const int cb = rem - (sg ? p.seg_kblocks[0] : 0);
          const CUtensorMap* mA = sg == 0 ? &tmA0 : &tmA1;
          const CUtensorMap* mB = sg == 0 ? &tmB0 : &tmB1;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * S::STAGE_BYTES;
          unsigned char* sb = sa + S::A_BYTES;
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          if (p.conv) tma_load_4d(sa, mA, &full_bar[stage], cb * BK, c.w0 + p.dw[tap], c.h0 + p.dh[tap], c.img0);
          else if (p.a_mn) { tma_load_2d(sa, mA, &full_bar[stage], c.m0, cb * BK); tma_load_2d(sa + 8192, mA, &full_bar[stage], c.m0 + 64, cb * BK); }
          else        tma_load_2d(sa, mA, &full_bar[stage], cb * BK, c.m0);
          if (p.b_mn) {
#pragma unroll
            for (int pn = 0; pn < BN / 64; ++pn) tma_load_2d(sb + pn * 8192, mB, &full_bar[stage], n0 + pn * 64, cb * BK);
          } else tma_load_2d(sb, mB, &full_bar[stage], tap * p.c_total + p.seg_bkoff[sg] + cb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++rem == kb_per_tap) { rem = 0; ++tap; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread), accumulator buffer it & 1 =====================
    if (lane == 0) {
      int stage = 0, phase = 0, it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int z = t / (p.tiles_m * tiles_n);
        const int kb_begin = z * p.kb_per_split;
        const int num_kb = min(num_kb_total, kb_begin + p.kb_per_split) - kb_begin;
        const int acc = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);        // epilogue has drained this buffer (passes at once on first use)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sb = sa + S::A_BYTES;
          const uint64_t da = p.a_mn ? make_mnmajor_sw128_desc(sa, 8192) : make_kmajor_sw128_desc(sa);
          const uint64_t db = p.b_mn ? make_mnmajor_sw128_desc(sb, 8192) : make_kmajor_sw128_desc(sb);
          const uint64_t ka = p.a_mn ? 128 : 2, kbs = p.b_mn ? 128 : 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16(d_tmem, da + ka * k, db + kbs * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (8 warps: a TMEM lane quarter is shared by two warps that take alternate 32-column
    // chunks).  With four warps the ~1000 dependent instructions per 128 x 160 tile ran on ONE warp per scheduler and the
    // epilogue, not the MMA, set the pace of K <= 640 GEMMs (ncu: issue slots 11 % busy, tensor pipe 21 %,
    // profiles/r01_gemm_k320_ncu_v5.md); two warps per scheduler hide each other's TMEM-load and conversion latency.
    const int q = warp & 3;                      // hardware rule: a warp reads TMEM lanes 32 * (warp % 4) ...
    const int half = (warp - 2) >> 2;            // 0: chunks 0, 2, 4 ... ; 1: chunks 1, 3, 5 ...
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;             // 0 .. 255
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const TileCoord c = decode_tile(p, t, tiles_n);
      const int n0 = c.tile_n * BN;
      const int acc = it & 1, aph = (it >> 1) & 1;
      long long m;
      bool row_ok;
      if (p.conv) {
        const int tw = r % p.TW, th = (r / p.TW) % p.TH, tn = r / (p.TW * p.TH);
        const int n_i = c.img0 + tn, hh = c.h0 + th, ww = c.w0 + tw;
        row_ok = (n_i < p.n_img) && (hh < p.H) && (ww < p.W);
        m = ((long long)n_i * p.H + hh) * p.W + ww;
      } else {
        m = (long long)c.m0 + r;
        row_ok = m < p.M;
      }
      const float* rv = (p.rowvec != nullptr && row_ok) ? p.rowvec + (m / p.rows_per_group) * p.rowvec_ld : nullptr;
      float* bias_buf = s_bias + (it & 1) * BN;
      for (int i = et; i < BN; i += 32 * PERSIST_EPI_WARPS) bias_buf[i] = (p.bias != nullptr && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
      constexpr bool RES_TMA_OK = (EPI == EPI_LEAN || EPI == EPI_GN);
      const bool res_staged = RES_TMA_OK && p.tma_res != 0;
      if (p.tma_store && warp == 2 && lane == 0) {
        bulk_wait_read<0>();                                              // previous tile's stores have read the staging tile
        if (res_staged) {
          // residual tile -> staging tile (the panels the result will overwrite): asynchronous, coalesced, no registers; it lands
          // while this tile's MMAs are still running.  (Per-thread 64-byte row loads cost +10 us on (32768, 320, 320).)
          int npan = 0;
          for (int pn = 0; pn < (BN + 31) / 32; ++pn) npan += (n0 + pn * 32 < p.N) ? 1 : 0;
          mbar_expect_tx(res_full, (uint32_t)npan * 8192u);
          for (int pn = 0; pn < npan; ++pn) {
            if (p.conv) tma_load_4d(staging + pn * 8192, &tmR, res_full, n0 + pn * 32, c.w0, c.h0, c.img0);
            else        tma_load_2d(staging + pn * 8192, &tmR, res_full, n0 + pn * 32, c.m0);
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * PERSIST_EPI_WARPS) : "memory");
      uint4 rcur[4], rnxt[4];
      bool have_nxt = !res_staged && (EPI <= EPI_GN || EPI == EPI_LEAN) && res_prefetch(p, rnxt, m, row_ok, n0, half * 32);
      mbar_wait(&tfull[acc], aph);
      if (res_staged) mbar_wait(res_full, it & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)(acc * ACC) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += 64) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(trow + (uint32_t)c0, v);
        const bool have = have_nxt;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) rcur[j4] = rnxt[j4];
        have_nxt = !res_staged && (EPI <= EPI_GN || EPI == EPI_LEAN) && (c0 + 64 < BN) && res_prefetch(p, rnxt, m, row_ok, n0, c0 + 64);
        tmem_ld_wait();
        epilogue_chunk<EPI>(p, v, c0, n0, m, row_ok, rv, bias_buf, r, staging, c.z, EPI == EPI_GN ? gn_warp_image(p, c.m0, c.img0, q) : 0, have ? rcur : nullptr,
                            res_staged);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);  // accumulator buffer free for tile it + 2
      if (p.tma_store) {
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * PERSIST_EPI_WARPS) : "memory");
        if (warp == 2 && lane == 0) {
#pragma unroll 1
          for (int pn = 0; pn < (BN + 31) / 32; ++pn) {
            if (n0 + pn * 32 >= p.N) break;
            if (p.conv) tma_store_4d(&tmO, staging + pn * 8192, n0 + pn * 32, c.w0, c.h0, c.img0);
            else        tma_store_2d(&tmO, staging + pn * 8192, n0 + pn * 32, c.m0);
          }
          bulk_commit();
        }
      }
    }
    if (p.tma_store && warp == 2 && lane == 0) bulk_wait_all<0>();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * ACC);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs on the two SMs of a TPC computes a 256 x BN tile with ONE
// tcgen05.mma per k-step.  Why: in the one-CTA kernels every operand byte crosses an SM's shared memory twice (written by
// TMA, read by the tensor core) - 128 x 256 tiles move 2 x 48 KB per 512 MMA cycles = 187 B/clk against ~128 B/clk of
// shared-memory bandwidth, which is exactly the 67-69 % tensor-pipe ceiling ncu shows for them (57-60 % at BN = 160:
// profiles/r01_gemm_ncu_v8.md).  In a pair each SM stores its own 128 rows of A and only HALF of B; the B halves are read by
// both tensor cores, so the same tile costs 2 x 32 KB per SM.
// Roles as in the persistent kernel (warp 0 TMA producer, warp 1 TMEM alloc (+ MMA issue on the leader), 8 epilogue
// warps); both producers complete their bytes on the LEADER's full barrier, the leader's tcgen05.commit multicasts the
// "stage free" / "accumulator ready" arrivals to both CTAs, and the peer's epilogue warps release the accumulator with
// a remote arrive on the leader's barrier.
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct PairSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;                 // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TILES = STAGES * STAGE_BYTES;
  static constexpr int STAGING_OFF = TILES;
  static constexpr int STAGING_BYTES = ((BN + 31) / 32) * 8192;
  static constexpr int BAR_OFF = STAGING_OFF + STAGING_BYTES;       // full[STAGES], empty[STAGES], tfull[2], tempty[2]
  static constexpr int TMEMPTR_OFF = BAR_OFF + (2 * STAGES + 5) * 8;   // ... + res_full
  static constexpr int BIAS_OFF = (TMEMPTR_OFF + 8 + 15) & ~15;
  static constexpr int TOTAL = BIAS_OFF + 2 * BN * 4 + 1024;
  static_assert(STAGE_BYTES % 1024 == 0, "stage tiles must stay 1024-byte aligned");
};

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const GemmKP p) {
  pdl_trigger();
  using S = PairSmem<BN, STAGES>;
  constexpr int ACC = tmem_cols<BN>();
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn;               // both CTAs of a pair must use identical offsets: no per-CTA realignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull = empty_bar + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* res_full = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S::TMEMPTR_OFF);
  float* s_bias = reinterpret_cast<float*>(smem + S::BIAS_OFF);
  unsigned char* staging = smem + S::STAGING_OFF;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int pairs_m = (p.tiles_m + 1) / 2;
  const int total_pairs = pairs_m * tiles_n * p.split_k;
  const int kb_per_tap = p.seg_kblocks[0] + (p.n_seg > 1 ? p.seg_kblocks[1] : 0);
  const int num_kb_total = p.n_taps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (p.n_seg > 1) { tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); }
    if (p.tma_store) tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 2 * PERSIST_EPI_WARPS); mbar_init(&tempty[1], 2 * PERSIST_EPI_WARPS);   // epilogue warps of BOTH CTAs
    mbar_init(res_full, 1);
    if (p.tma_res) tma_prefetch_desc(&tmR);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_ptr, 2 * ACC);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): own A rows + own half of B, bytes counted on the leader's barrier
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int t = cluster_id; t < total_pairs; t += n_clusters) {
        const TileCoord c = decode_pair_tile(p, t, tiles_n, pairs_m, (int)rank);
        const int nb0 = c.tile_n * BN + (int)rank * (BN / 2);
        const int kb_begin = c.z * p.kb_per_split;
        const int kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
        int tap = kb_begin / kb_per_tap;
        int rem = kb_begin - tap * kb_per_tap;
        for (int kbi = kb_begin; kbi < kb_end; ++kbi) {
          const int sg = (rem >= p.seg_kblocks[0]) ? 1 : 0;
          const int cb = rem - (sg ? p.seg_kblocks[0] : 0);
          const CUtensorMap* mA = sg == 0 ? &tmA0 : &tmA1;
          const CUtensorMap* mB = sg == 0 ? &tmB0 : &tmB1;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * S::STAGE_BYTES;
          unsigned char* sb = sa + S::A_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
          if (p.conv) tma_load_4d_2sm(sa, mA, &full_bar[stage], cb * BK, c.w0 + p.dw[tap], c.h0 + p.dh[tap], c.img0);
          else        tma_load_2d_2sm(sa, mA, &full_bar[stage], cb * BK, c.m0);
          tma_load_2d_2sm(sb, mB, &full_bar[stage], tap * p.c_total + p.seg_bkoff[sg] + cb * BK, nb0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++rem == kb_per_tap) { rem = 0; ++tap; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer: one thread of the LEADER drives both tensor cores
    if (lane == 0 && rank == 0) {
      int stage = 0, phase = 0, it = 0;
      for (int t = cluster_id; t < total_pairs; t += n_clusters, ++it) {
        const int z = t / (pairs_m * tiles_n);
        const int kb_begin = z * p.kb_per_split;
        const int num_kb = min(num_kb_total, kb_begin + p.kb_per_split) - kb_begin;
        const int acc = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint64_t da = make_kmajor_sw128_desc(sa);
          const uint64_t db = make_kmajor_sw128_desc(sa + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_2sm(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage], 3);          // both producers may refill this stage
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull[acc], 3);                  // both epilogues may drain this accumulator
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (8 warps per CTA; this CTA's 128 accumulator rows) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    int it = 0;
    for (int t = cluster_id; t < total_pairs; t += n_clusters, ++it) {
      const TileCoord c = decode_pair_tile(p, t, tiles_n, pairs_m, (int)rank);
      const int n0 = c.tile_n * BN;
      const int acc = it & 1, aph = (it >> 1) & 1;
      long long m;
      bool row_ok;
      if (p.conv) {
        const int tw = r % p.TW, th = (r / p.TW) % p.TH, tn = r / (p.TW * p.TH);
        const int n_i = c.img0 + tn, hh = c.h0 + th, ww = c.w0 + tw;
        row_ok = (n_i < p.n_img) && (hh < p.H) && (ww < p.W);
        m = ((long long)n_i * p.H + hh) * p.W + ww;
      } else {
        m = (long long)c.m0 + r;
        row_ok = m < p.M;
      }
      const float* rv = (p.rowvec != nullptr && row_ok) ? p.rowvec + (m / p.rows_per_group) * p.rowvec_ld : nullptr;
      float* bias_buf = s_bias + (it & 1) * BN;
      for (int i = et; i < BN; i += 32 * PERSIST_EPI_WARPS) bias_buf[i] = (p.bias != nullptr && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
      constexpr bool RES_TMA_OK = (EPI == EPI_LEAN || EPI == EPI_GN);
      const bool res_staged = RES_TMA_OK && p.tma_res != 0;
      if (p.tma_store && warp == 2 && lane == 0) {
        bulk_wait_read<0>();                                              // previous tile's stores have read the staging tile
        if (res_staged) {
          // residual tile -> staging tile (the panels the result will overwrite): asynchronous, coalesced, no registers; it lands
          // while this tile's MMAs are still running.  (Per-thread 64-byte row loads cost +10 us on (32768, 320, 320).)
          int npan = 0;
          for (int pn = 0; pn < (BN + 31) / 32; ++pn) npan += (n0 + pn * 32 < p.N) ? 1 : 0;
          mbar_expect_tx(res_full, (uint32_t)npan * 8192u);
          for (int pn = 0; pn < npan; ++pn) {
            if (p.conv) tma_load_4d(staging + pn * 8192, &tmR, res_full, n0 + pn * 32, c.w0, c.h0, c.img0);
            else        tma_load_2d(staging + pn * 8192, &tmR, res_full, n0 + pn * 32, c.m0);
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * PERSIST_EPI_WARPS) : "memory");
      uint4 rcur[4], rnxt[4];
      bool have_nxt = !res_staged && (EPI <= EPI_GN || EPI == EPI_LEAN) && res_prefetch(p, rnxt, m, row_ok, n0, half * 32);
      mbar_wait(&tfull[acc], aph);
      if (res_staged) mbar_wait(res_full, it & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)(acc * ACC) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += 64) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(trow + (uint32_t)c0, v);
        const bool have = have_nxt;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) rcur[j4] = rnxt[j4];
        have_nxt = !res_staged && (EPI <= EPI_GN || EPI == EPI_LEAN) && (c0 + 64 < BN) && res_prefetch(p, rnxt, m, row_ok, n0, c0 + 64);
        tmem_ld_wait();
        epilogue_chunk<EPI>(p, v, c0, n0, m, row_ok, rv, bias_buf, r, staging, c.z, EPI == EPI_GN ? gn_warp_image(p, c.m0, c.img0, q) : 0, have ? rcur : nullptr,
                            res_staged);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                           // accumulator buffer free for pair-tile it + 2: tell the leader's MMA thread
        if (rank == 0) mbar_arrive(&tempty[acc]);
        else           mbar_arrive_cluster(&tempty[acc], 0);
      }
      if (p.tma_store) {
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * PERSIST_EPI_WARPS) : "memory");
        if (warp == 2 && lane == 0) {
#pragma unroll 1
          for (int pn = 0; pn < (BN + 31) / 32; ++pn) {
            if (n0 + pn * 32 >= p.N) break;
            if (p.conv) tma_store_4d(&tmO, staging + pn * 8192, n0 + pn * 32, c.w0, c.h0, c.img0);
            else        tma_store_2d(&tmO, staging + pn * 8192, n0 + pn * 32, c.m0);
          }
          bulk_commit();
        }
      }
    }
    if (p.tma_store && warp == 2 && lane == 0) bulk_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // the peer's shared memory / barriers stay alive until both CTAs are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * ACC);
  }
}

// split-K second pass, four columns per thread (N % 4 == 0, 16-byte aligned rows): the partials are read as float4 - the scalar
// version below ran at ~24 us per call on the SDXL step's small-M GEMMs (54 ms of a 481 ms step, profiles/r02_cfg4_sdxl_step_kernels_*)
__global__ void __launch_bounds__(256) splitk_reduce4_kernel(const GemmKP p, int accumulate) {
  pdl_grid_dependency_sync();
  const long long idx4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)p.M * p.N;
  if (idx4 * 4 >= total) return;
  const long long m = (idx4 * 4) / p.N;
  const int n = (int)((idx4 * 4) - m * p.N);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* ws4 = reinterpret_cast<const float4*>(p.splitk_ws);
  const long long stride4 = total / 4;
#pragma unroll 4
  for (int s = 0; s < p.split_k; ++s) {
    const float4 v = ws4[(size_t)s * stride4 + idx4];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float x[4] = {acc.x * p.alpha, acc.y * p.alpha, acc.z * p.alpha, acc.w * p.alpha};
  if (p.bias) {
    const float4 b = *reinterpret_cast<const float4*>(p.bias + n);
    x[0] += b.x; x[1] += b.y; x[2] += b.z; x[3] += b.w;
  }
  if (p.rowvec) {
    const float* rv = p.rowvec + (m / p.rows_per_group) * p.rowvec_ld + n;
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] += rv[i];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = act_apply(x[i], p.act);
  if (p.residual) {
    const uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.residual) + m * p.res_ld + n);
    const uint32_t w[2] = {r.x, r.y};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (p.is_bf16) { x[2 * e] += __uint_as_float(w[e] << 16); x[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u); }
      else { const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[e])); x[2 * e] += t.x; x[2 * e + 1] += t.y; }
    }
  }
  if (p.out32) {
    float4* o = reinterpret_cast<float4*>(p.out32 + m * p.out32_ld + n);
    float4 v = make_float4(x[0], x[1], x[2], x[3]);
    if (accumulate) { const float4 c = *o; v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w; }
    *o = v;
  }
  if (p.out16) {
    uint2 u;
    u.x = pack16(x[0], x[1], p.is_bf16); u.y = pack16(x[2], x[3], p.is_bf16);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out16) + m * p.out_ld + n) = u;
  }
}

// split-K second pass: out = act(alpha * sum_s ws[s] + bias + rowvec) + residual  (+ accumulate into out32)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const GemmKP p, int accumulate) {
  pdl_grid_dependency_sync();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)p.M * p.N;
  if (idx >= total) return;
  const long long m = idx / p.N;
  const int n = (int)(idx % p.N);
  float acc = 0.f;
  for (int s = 0; s < p.split_k; ++s) acc += p.splitk_ws[(size_t)s * total + idx];
  float x = acc * p.alpha;
  if (p.bias) x += p.bias[n];
  if (p.rowvec) x += p.rowvec[(m / p.rows_per_group) * p.rowvec_ld + n];
  x = act_apply(x, p.act);
  if (p.residual) {
    const uint16_t b = reinterpret_cast<const uint16_t*>(p.residual)[m * p.res_ld + n];
    x += p.is_bf16 ? __uint_as_float((uint32_t)b << 16) : __half2float(*reinterpret_cast<const __half*>(&b));
  }
  if (p.out32) { float* o = p.out32 + m * p.out32_ld + n; *o = accumulate ? (*o + x) : x; }
  if (p.out16) {
    uint16_t* o = reinterpret_cast<uint16_t*>(p.out16) + m * p.out_ld + n;
    if (p.is_bf16) { const __nv_bfloat16 t = __float2bfloat16_rn(x); *o = *reinterpret_cast<const uint16_t*>(&t); }
    else           { const __half t = __float2half_rn(x);           *o = *reinterpret_cast<const uint16_t*>(&t); }
  }
}

template <int BN, int STAGES, int EPI>
static int launch_gemm(const CUtensorMap* maps, const GemmKP& kp, dim3 grid, cudaStream_t st) {
  using S = GemmSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    COMAT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  launch_k(gemm_tc_kernel<BN, STAGES, EPI>, grid, GEMM_THREADS, S::TOTAL, st, maps[0], maps[1], maps[2], maps[3], maps[4], kp);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

template <int BN, int STAGES, int EPI>
static int launch_gemm_persist(const CUtensorMap* maps, const GemmKP& kp, int total_tiles, cudaStream_t st) {
  using S = PersistSmem<BN, STAGES>;
  static_assert(S::TOTAL <= 232448, "persistent GEMM smem budget");
  static bool configured = false;
  if (!configured) {
    COMAT_CUDA(cudaFuncSetAttribute(gemm_tc_persist_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  launch_k(gemm_tc_persist_kernel<BN, STAGES, EPI>, grid, PERSIST_THREADS, S::TOTAL, st, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], kp);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

template <int BN, int STAGES, int EPI>
static int launch_gemm_pair(const CUtensorMap* maps, const GemmKP& kp, int total_pairs, cudaStream_t st) {
  using S = PairSmem<BN, STAGES>;
  static_assert(S::TOTAL <= 232448, "pair GEMM smem budget");
  static bool configured = false;
  if (!configured) {
    COMAT_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  const int max_clusters = num_sms() / 2;
  const int clusters = total_pairs < max_clusters ? total_pairs : max_clusters;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(PERSIST_THREADS); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = comat_pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<BN, STAGES, EPI>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], kp);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

}  // namespace comat

using namespace comat;

static int pick_bn(int N, int forced) {
  if (forced == 32 || forced == 64 || forced == 128 || forced == 160 || forced == 256) return forced;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N % 160 == 0) return 160;
  if (N % 128 == 0 || N > 1024) return 128;
  if (N % 64 == 0 && N < 128) return 64;
  return (N % 160) > (N % 128) || (N % 128 == 0) ? 128 : 160;
}

extern "C" int comat_gemm(const comat_gemm_params* g, void* stream) {
  if (!g || g->M <= 0 || g->N <= 0 || g->n_seg < 1 || g->n_seg > 2) return COMAT_ERR_INVALID;
  if (g->dtype != COMAT_F16 && g->dtype != COMAT_BF16) return COMAT_ERR_UNSUPPORTED;
  if (!g->out16 && !g->out32) return COMAT_ERR_INVALID;
  GemmKP kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = g->M; kp.N = g->N; kp.n_seg = g->n_seg;
  kp.alpha = g->alpha; kp.bias = g->bias; kp.rowvec = g->rowvec; kp.rowvec_ld = g->rowvec_ld > 0 ? g->rowvec_ld : g->N; kp.rows_per_group = g->rows_per_group > 0 ? g->rows_per_group : 1;
  kp.act = g->act; kp.residual = g->residual; kp.res_ld = g->res_ld; kp.out16 = g->out16; kp.out_ld = g->out_ld;
  kp.out32 = g->out32; kp.out32_ld = g->out32_ld; kp.is_bf16 = g->dtype == COMAT_BF16;
  int BN = pick_bn(g->N, g->force_bn);
  // 256-wide tiles need 25 % less smem operand traffic per FLOP than 160-wide ones; take them when N divides and
  // there are enough of them to fill the machine (profiles/r01_gemm_persist_vs_tile.md)
  if (g->force_bn == 0 && (g->N % 256) == 0 && (long long)((g->M + BM - 1) / BM) * (g->N / 256) >= num_sms()) BN = 256;
  const int a_mn = g->a_mn_major ? 1 : 0, b_mn = g->b_mn_major ? 1 : 0;
  if (a_mn || b_mn) {
    if (g->conv || g->n_seg != 1) return COMAT_ERR_UNSUPPORTED;
    if (b_mn && (BN % 64) != 0) BN = g->N > 64 ? 128 : 64;        // B panels are 64 columns wide
  }
  kp.a_mn = a_mn; kp.b_mn = b_mn;
  // CTA-pair kernel (cta_group::2; force_kernel == 3 or chosen here): K-major operands, tile widths whose halves are whole
  // 8-row swizzle groups.  Automatic choice: K >= 640 (shorter K is epilogue-bound, where the pair buys nothing) and a
  // tile width whose pair-tiles fill the 74 SM pairs: score = N-tile efficiency x wave efficiency x width factor
  // (wider tiles move fewer operand bytes per FLOP).  COMAT_GEMM_PAIR=0 disables the automatic choice.
  bool use_pair = g->force_kernel == 3 && !a_mn && !b_mn;
  if (use_pair && BN != 128 && BN != 160 && BN != 256) BN = (g->N % 256 == 0 || g->N > 512) ? 256 : (g->N % 160 == 0 ? 160 : 128);
  if (g->force_kernel == 0 && !a_mn && !b_mn) {
    static int pair_mode = -1;
    if (pair_mode < 0) { const char* e = getenv("COMAT_GEMM_PAIR"); pair_mode = (e && e[0] == '0') ? 0 : 1; }
    int kb_est = 0;
    for (int s2 = 0; s2 < g->n_seg; ++s2) kb_est += (g->a_k[s2] + BK - 1) / BK;
    kb_est *= g->conv ? g->n_taps : 1;
    const int tiles_m_est = (g->M + BM - 1) / BM;
    if (pair_mode == 1 && kb_est >= 10 && tiles_m_est >= 2 && g->N >= 128) {
      const int split = g->split_k > 1 ? g->split_k : 1;
      const int clusters = num_sms() / 2;
      const int cand[3] = {256, 160, 128};
      const float wfac[3] = {1.0f, 0.93f, 0.88f};
      float best = 0.f; int best_bn = 0;
      for (int c = 0; c < 3; ++c) {
        if (g->force_bn != 0 && g->force_bn != cand[c]) continue;
        const int tn = (g->N + cand[c] - 1) / cand[c];
        const int pairs = ((tiles_m_est + 1) / 2) * tn * split;
        const float wave = (float)pairs / (float)(((pairs + clusters - 1) / clusters) * clusters);
        const float score = wave * ((float)g->N / (float)(tn * cand[c])) * wfac[c];
        if (score > best) { best = score; best_bn = cand[c]; }
      }
      if (best >= 0.6f) { use_pair = true; BN = best_bn; }
    }
  }
  // kind::f16 takes ONE 16-bit format for both operands: a descriptor with a_format != b_format raises an illegal-instruction
  // fault on sm_100a (measured, profiles/r01_gpu_tests_run13.log), so mixed fp16 x bf16 operands are rejected here
  if (g->b_dtype != 0 && g->b_dtype != g->dtype) return COMAT_ERR_UNSUPPORTED;
  kp.idesc = make_idesc_f16(use_pair ? 2 * BM : BM, BN, kp.is_bf16 ? 1 : 0, a_mn, b_mn);
  CUtensorMap maps[6];
  memset(maps, 0, sizeof(maps));
  dim3 grid;
  kp.conv = g->conv ? 1 : 0;
  if (kp.conv) {
    if (g->n_taps < 1 || g->n_taps > 9 || g->H <= 0 || g->W <= 0 || g->n_img <= 0) return COMAT_ERR_INVALID;
    if ((long long)g->n_img * g->H * g->W != g->M) return COMAT_ERR_INVALID;
    kp.H = g->H; kp.W = g->W; kp.n_img = g->n_img; kp.n_taps = g->n_taps; kp.c_total = g->c_total;
    for (int t = 0; t < g->n_taps; ++t) { kp.dh[t] = g->tap_dh[t]; kp.dw[t] = g->tap_dw[t]; }
    // spatial tile: 128 output pixels = TW x TH x TN
    int TW = 1;
    while (TW < g->W && TW < 128) TW <<= 1;            // power of two >= W, capped at 128
    if (TW > 128) TW = 128;
    int TH = 128 / TW, TN = 1;
    if (TH > g->H) {                                   // small maps: several images per tile
      int th = 1;
      while (th < g->H) th <<= 1;
      TH = th < 128 / TW ? th : 128 / TW;
      TN = 128 / (TW * TH);
    }
    kp.TW = TW; kp.TH = TH; kp.TN = TN;
    kp.tiles_w = (g->W + TW - 1) / TW; kp.tiles_h = (g->H + TH - 1) / TH;
    const int tiles_n = (g->n_img + TN - 1) / TN;
    grid = dim3(kp.tiles_w * kp.tiles_h * tiles_n, (g->N + BN - 1) / BN, 1);
    kp.tiles_m = (int)grid.x;
  } else {
    kp.n_taps = 1; kp.c_total = 0;
    grid = dim3((g->M + BM - 1) / BM, (g->N + BN - 1) / BN, 1);
    kp.tiles_m = (int)grid.x;
  }
  for (int s = 0; s < g->n_seg; ++s) {
    const int K = g->a_k[s];
    if (K <= 0 || ((!a_mn || !b_mn) && (K % 8) != 0) || !g->a[s] || !g->b[s]) return COMAT_ERR_INVALID;   // K-major rows need 16-B pitch
    if ((reinterpret_cast<uintptr_t>(g->a[s]) & 15) || (reinterpret_cast<uintptr_t>(g->b[s]) & 15)) return COMAT_ERR_INVALID;
    if (kp.conv && (K % BK) != 0) return COMAT_ERR_UNSUPPORTED;   // channel slabs of 64
    kp.seg_kblocks[s] = (K + BK - 1) / BK;
    kp.seg_bkoff[s] = g->b_koff[s];
    if (kp.conv) {
      const uint64_t dims[4] = {(uint64_t)K, (uint64_t)g->W, (uint64_t)g->H, (uint64_t)g->n_img};
      const uint64_t str[3] = {(uint64_t)K * 2, (uint64_t)K * 2 * g->W, (uint64_t)K * 2 * g->W * g->H};
      const uint32_t box[4] = {(uint32_t)BK, (uint32_t)kp.TW, (uint32_t)kp.TH, (uint32_t)kp.TN};
      if (!make_tmap_16bit(&maps[s], g->a[s], 4, dims, str, box)) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
    } else if (a_mn) {                                   // A stored [K, M] row-major: panels of 64 M-columns x 64 k-rows
      if ((g->a_ld[s] % 8) != 0) return COMAT_ERR_INVALID;
      const uint64_t dims[2] = {(uint64_t)g->M, (uint64_t)K};
      const uint64_t str[1] = {(uint64_t)g->a_ld[s] * 2};
      const uint32_t box[2] = {64u, (uint32_t)BK};
      if (!make_tmap_16bit(&maps[s], g->a[s], 2, dims, str, box)) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
    } else {
      if ((g->a_ld[s] % 8) != 0) return COMAT_ERR_INVALID;
      const uint64_t dims[2] = {(uint64_t)K, (uint64_t)g->M};
      const uint64_t str[1] = {(uint64_t)g->a_ld[s] * 2};
      const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
      if (!make_tmap_16bit(&maps[s], g->a[s], 2, dims, str, box)) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
    }
    if (b_mn) {                                          // B stored [K, N] row-major
      if ((g->b_ld[s] % 8) != 0) return COMAT_ERR_INVALID;
      const uint64_t dims[2] = {(uint64_t)g->N, (uint64_t)K};
      const uint64_t str[1] = {(uint64_t)g->b_ld[s] * 2};
      const uint32_t box[2] = {64u, (uint32_t)BK};
      if (!make_tmap_16bit(&maps[2 + s], g->b[s], 2, dims, str, box)) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
    } else {
      if ((g->b_ld[s] % 8) != 0) return COMAT_ERR_INVALID;
      const uint64_t kext = kp.conv ? (uint64_t)g->n_taps * g->c_total : (uint64_t)g->b_koff[s] + K;
      const uint64_t dims[2] = {kext, (uint64_t)g->N};
      const uint64_t str[1] = {(uint64_t)g->b_ld[s] * 2};
      const uint32_t box[2] = {(uint32_t)BK, (uint32_t)(use_pair ? BN / 2 : BN)};      // pair: each CTA loads half of the B tile
      if (!make_tmap_16bit(&maps[2 + s], g->b[s], 2, dims, str, box)) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
    }
  }
  // split-K (caller supplies the fp32 partial workspace)
  {
    const int kb_total = kp.n_taps * (kp.seg_kblocks[0] + (g->n_seg > 1 ? kp.seg_kblocks[1] : 0));
    int sk = g->split_k > 1 ? g->split_k : 1;
    if (sk > kb_total) sk = kb_total;
    if (sk > 1) {
      if (!g->splitk_ws && !g->accumulate) return COMAT_ERR_WORKSPACE;
      kp.kb_per_split = (kb_total + sk - 1) / sk;
      sk = (kb_total + kp.kb_per_split - 1) / kp.kb_per_split;
      kp.split_k = sk; kp.splitk_ws = g->splitk_ws;
      grid.z = sk;
    } else { kp.split_k = 1; kp.kb_per_split = kb_total; }
    if (g->accumulate) {
      // accumulate = fp32 atomics straight into out32 (alpha only: no bias / activation / residual / 16-bit output)
      if (!g->out32 || g->out16 || g->bias || g->rowvec || g->act || g->residual) return COMAT_ERR_INVALID;
      kp.atomic_acc = 1;
    }
  }
  if (g->act == 3) {
    // fused GEGLU epilogue: half-width 16-bit output, written with direct 32-byte stores
    if (!g->out16 || g->out32 || g->residual || g->rowvec || kp.split_k > 1 || g->accumulate || (g->N & 1) || a_mn || b_mn) return COMAT_ERR_INVALID;
  }
  // TMA-store epilogue (default): needs a 16-bit output with 16-byte aligned base and row pitch
  static int epi_mode = -1;
  if (epi_mode < 0) { const char* e = getenv("COMAT_GEMM_EPILOGUE"); epi_mode = (e && !strcmp(e, "direct")) ? 0 : 1; }
  kp.tma_store = 0;
  if (epi_mode == 1 && g->act != 3 && kp.split_k == 1 && g->out16 && !g->out32 && (g->out_ld % 8) == 0 && (reinterpret_cast<uintptr_t>(g->out16) & 15) == 0 &&
      (!kp.conv || g->out_ld == g->N)) {
    bool ok;
    if (kp.conv) {
      const uint64_t dims[4] = {(uint64_t)g->N, (uint64_t)g->W, (uint64_t)g->H, (uint64_t)g->n_img};
      const uint64_t str[3] = {(uint64_t)g->out_ld * 2, (uint64_t)g->out_ld * 2 * g->W, (uint64_t)g->out_ld * 2 * g->W * g->H};
      const uint32_t box[4] = {32u, (uint32_t)kp.TW, (uint32_t)kp.TH, (uint32_t)kp.TN};
      ok = make_tmap_16bit(&maps[4], g->out16, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
    } else {
      const uint64_t dims[2] = {(uint64_t)g->N, (uint64_t)g->M};
      const uint64_t str[1] = {(uint64_t)g->out_ld * 2};
      const uint32_t box[2] = {32u, (uint32_t)BM};
      ok = make_tmap_16bit(&maps[4], g->out16, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
    }
    kp.tma_store = ok ? 1 : 0;
  }
  // residual through TMA (persistent / pair kernels, lean flavours): same geometry and panel layout as the output tile.
  // COMAT_GEMM_TMA_RES=0 keeps the per-thread (prefetched) row loads.
  {
    static int res_mode = -1;
    if (res_mode < 0) { const char* e = getenv("COMAT_GEMM_TMA_RES"); res_mode = (e && e[0] == '0') ? 0 : 1; }
    kp.tma_res = 0;
    if (res_mode == 1 && kp.tma_store && g->residual && g->act == 0 && (g->res_ld % 8) == 0 && (reinterpret_cast<uintptr_t>(g->residual) & 15) == 0 &&
        (!kp.conv || g->res_ld == g->N)) {
      bool ok;
      if (kp.conv) {
        const uint64_t dims[4] = {(uint64_t)g->N, (uint64_t)g->W, (uint64_t)g->H, (uint64_t)g->n_img};
        const uint64_t str[3] = {(uint64_t)g->res_ld * 2, (uint64_t)g->res_ld * 2 * g->W, (uint64_t)g->res_ld * 2 * g->W * g->H};
        const uint32_t box[4] = {32u, (uint32_t)kp.TW, (uint32_t)kp.TH, (uint32_t)kp.TN};
        ok = make_tmap_16bit(&maps[5], g->residual, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
      } else {
        const uint64_t dims[2] = {(uint64_t)g->N, (uint64_t)g->M};
        const uint64_t str[1] = {(uint64_t)g->res_ld * 2};
        const uint32_t box[2] = {32u, (uint32_t)BM};
        ok = make_tmap_16bit(&maps[5], g->residual, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
      }
      kp.tma_res = ok ? 1 : 0;
    }
  }
  if (g->gn_sums != nullptr) {
    // GroupNorm statistics of the output in the epilogue: needs the TMA-store epilogue (every epilogue thread walks every chunk,
    // so the warp shuffles are convergent), complete sums per CTA (no split-K: statistics taken in the reduction pass were
    // measured 3.5 ms / step SLOWER than the two-pass GroupNorm on the 8x8 / 16x16 levels that use split-K) and lane quarters
    // that stay inside one image
    const int G = g->gn_groups;
    bool ok = G > 0 && (g->N % G) == 0 && !kp.atomic_acc && !a_mn && !b_mn && g->act == 0 && kp.split_k == 1 && kp.tma_store != 0;
    const int rpi = kp.conv ? g->H * g->W : g->gn_rows_per_image;
    if (ok && kp.conv) ok = kp.TW * kp.TH >= 32;
    else if (ok) ok = rpi >= 32 && (rpi % 32) == 0 && (g->M % rpi) == 0;
    if (!ok) return COMAT_ERR_UNSUPPORTED;
    kp.gn_sums = g->gn_sums; kp.gn_G = G; kp.gn_cpg = g->N / G;
    kp.gn_rows_per_img = rpi;
    kp.gn_n_img = g->M / rpi;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = COMAT_ERR_UNSUPPORTED;
  // kernel choice (COMAT_GEMM_KERNEL=tile|persist forces one for A/B measurements): the persistent kernel wins when the
  // per-tile prologue/epilogue dominates (short K) or when there is at most one wave of tiles (deeper ring covers the
  // load latency); with long K and several waves two co-resident one-tile CTAs keep more loads in flight per SM.
  static int kernel_mode = -1;
  if (kernel_mode < 0) {
    const char* e = getenv("COMAT_GEMM_KERNEL");
    kernel_mode = (e && !strcmp(e, "tile")) ? 0 : (e && !strcmp(e, "persist")) ? 1 : 2;
  }
  const int total_tiles = (int)(grid.x * grid.y * grid.z);
  const int kb_all = kp.n_taps * (kp.seg_kblocks[0] + (g->n_seg > 1 ? kp.seg_kblocks[1] : 0));
  const bool use_persist = g->force_kernel == 2 || (g->force_kernel != 1 && (kernel_mode == 1 || (kernel_mode == 2 && (kb_all <= 24 || total_tiles <= num_sms()))));
  // epilogue flavour (compile-time specialisation of every kernel, see EPI_*)
  const int epi = kp.atomic_acc ? EPI_ATOMIC : kp.split_k > 1 ? EPI_SPLITK : g->act == 3 ? EPI_GEGLU : kp.gn_sums != nullptr ? EPI_GN :
                  (kp.tma_store && g->act == 0 && !g->out32) ? EPI_LEAN : EPI_STD;
#define COMAT_EPI_SWITCH(CALL)                                   \
  switch (epi) {                                                 \
    case EPI_STD:    rc = CALL(EPI_STD); break;                  \
    case EPI_GN:     rc = CALL(EPI_GN); break;                   \
    case EPI_SPLITK: rc = CALL(EPI_SPLITK); break;               \
    case EPI_ATOMIC: rc = CALL(EPI_ATOMIC); break;               \
    case EPI_LEAN:   rc = CALL(EPI_LEAN); break;                 \
    default:         rc = CALL(EPI_GEGLU); break;                \
  }
  if (use_pair) {
    const int pairs = (int)(((grid.x + 1) / 2) * grid.y * grid.z);
#define PAIR_CALL_128(E) launch_gemm_pair<128, 6, E>(maps, kp, pairs, st)
#define PAIR_CALL_160(E) launch_gemm_pair<160, 6, E>(maps, kp, pairs, st)
#define PAIR_CALL_256(E) launch_gemm_pair<256, 4, E>(maps, kp, pairs, st)
    switch (BN) {
      case 128: COMAT_EPI_SWITCH(PAIR_CALL_128) break;
      case 160: COMAT_EPI_SWITCH(PAIR_CALL_160) break;
      case 256: COMAT_EPI_SWITCH(PAIR_CALL_256) break;
    }
  } else if (use_persist) {
#define PERS_CALL_32(E) launch_gemm_persist<32, 8, E>(maps, kp, total_tiles, st)
#define PERS_CALL_64(E) launch_gemm_persist<64, 8, E>(maps, kp, total_tiles, st)
#define PERS_CALL_128(E) launch_gemm_persist<128, 6, E>(maps, kp, total_tiles, st)
#define PERS_CALL_160(E) launch_gemm_persist<160, 5, E>(maps, kp, total_tiles, st)
#define PERS_CALL_256(E) launch_gemm_persist<256, 3, E>(maps, kp, total_tiles, st)
    switch (BN) {
      case 32:  COMAT_EPI_SWITCH(PERS_CALL_32) break;
      case 64:  COMAT_EPI_SWITCH(PERS_CALL_64) break;
      case 128: COMAT_EPI_SWITCH(PERS_CALL_128) break;
      case 160: COMAT_EPI_SWITCH(PERS_CALL_160) break;
      case 256: COMAT_EPI_SWITCH(PERS_CALL_256) break;
    }
  } else {
#define TILE_CALL_32(E) launch_gemm<32, 4, E>(maps, kp, grid, st)
#define TILE_CALL_64(E) launch_gemm<64, 4, E>(maps, kp, grid, st)
#define TILE_CALL_128(E) launch_gemm<128, 3, E>(maps, kp, grid, st)
#define TILE_CALL_160(E) launch_gemm<160, 3, E>(maps, kp, grid, st)
#define TILE_CALL_256(E) launch_gemm<256, 4, E>(maps, kp, grid, st)
    switch (BN) {
      case 32:  COMAT_EPI_SWITCH(TILE_CALL_32) break;
      case 64:  COMAT_EPI_SWITCH(TILE_CALL_64) break;
      case 128: COMAT_EPI_SWITCH(TILE_CALL_128) break;
      case 160: COMAT_EPI_SWITCH(TILE_CALL_160) break;
      case 256: COMAT_EPI_SWITCH(TILE_CALL_256) break;
    }
  }
#undef COMAT_EPI_SWITCH
  if (rc == COMAT_OK && kp.split_k > 1 && !kp.atomic_acc) {
    const long long total = (long long)kp.M * kp.N;
    auto al = [](const void* q, uintptr_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
    const bool vec4 = (kp.N % 4) == 0 && al(kp.splitk_ws, 16) && al(kp.bias, 16) && al(kp.out32, 16) && (kp.out32_ld % 4) == 0 &&
                      al(kp.out16, 8) && (kp.out_ld % 4) == 0 && al(kp.residual, 8) && (kp.res_ld % 4) == 0 &&
                      (kp.rowvec == nullptr || (kp.rowvec_ld % 4) == 0);
    if (vec4) launch_k(splitk_reduce4_kernel, (unsigned)((total / 4 + 255) / 256), 256, 0, st, kp, g->accumulate ? 1 : 0);
    else launch_k(splitk_reduce_kernel, (unsigned)((total + 255) / 256), 256, 0, st, kp, g->accumulate ? 1 : 0);
    COMAT_CHECK_LAUNCH();
  }
  return rc;
}

// 1 when comat_gemm would accept p->gn_sums for this problem (same conditions, evaluated without launching anything):
// callers use it to decide between the fused statistics and the two-pass GroupNorm.
extern "C" int comat_gemm_gn_supported(const comat_gemm_params* g) {
  if (!g || g->gn_groups <= 0 || (g->N % g->gn_groups) != 0 || g->a_mn_major || g->b_mn_major) return 0;
  if (g->accumulate || g->act != 0 || g->n_seg < 1 || g->n_seg > 2) return 0;
  const int rpi = g->conv ? g->H * g->W : g->gn_rows_per_image;
  if (rpi <= 0 || (g->M % rpi) != 0) return 0;
  int kb_total = 0;
  for (int s = 0; s < g->n_seg; ++s) kb_total += (g->a_k[s] + BK - 1) / BK;
  kb_total *= g->conv ? g->n_taps : 1;
  int sk = g->split_k > 1 ? g->split_k : 1;
  if (sk > kb_total) sk = kb_total;
  if (sk > 1) { const int per = (kb_total + sk - 1) / sk; sk = (kb_total + per - 1) / per; }
  if (sk > 1) return 0;
  if (!g->out16 || g->out32) return 0;
  const char* e = getenv("COMAT_GEMM_EPILOGUE");
  if (e && !strcmp(e, "direct")) return 0;
  if ((g->out_ld % 8) != 0 || (reinterpret_cast<uintptr_t>(g->out16) & 15) != 0) return 0;
  if (g->conv) {
    if (g->out_ld != g->N || g->W <= 0 || g->H <= 0) return 0;
    int TW = 1;
    while (TW < g->W && TW < 128) TW <<= 1;
    int TH = 128 / TW;
    if (TH > g->H) { int th = 1; while (th < g->H) th <<= 1; TH = th < 128 / TW ? th : 128 / TW; }
    return TW * TH >= 32 ? 1 : 0;
  }
  return (rpi >= 32 && (rpi % 32) == 0) ? 1 : 0;
}
