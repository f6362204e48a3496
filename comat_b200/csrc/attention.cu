// Fused multi-head attention forward on tcgen05:  O = softmax(scale * Q K^T) V  per (sample, head), flash-style
// (online softmax over 64-key tiles; S and the running P.V product live in TMEM), with
//   * optional export of the normalised fp32 probabilities when all keys fit two tiles (UNet cross-attention, 77 text
//     tokens): this is the tensor the reference's hooked Attention.forward hands to AttentionStore
//     (attn_utils/tc_attn_utils.py:126-145, :60-68) — written once, never re-read by this kernel;
//   * the log-sum-exp per row saved for the backward pass.
// Replaces F.scaled_dot_product_attention / baddbmm+softmax+bmm of diffusers' Attention and HF BLIP's eager attention.
//
// Layout: q, k, v (n, L, H*d) 16-bit token-major (the projection GEMMs' natural output, or column slices of a fused
// q|k|v projection: row pitches are parameters), read as they lie: a [64 keys][d] tile of V lands in smem as
// 128-byte-swizzled rows, which is the canonical MN-major B operand of the P.V product (reduction over the key rows) -
// no V^T copy exists.  Head dims that are not multiples of 64 (40, 80, 160) are handled by a 3-D tensor map {d, H, rows}:
// the TMA box is 64 wide and elements past d are out-of-bounds -> zero-filled, so no padded copies of Q/K exist.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-5 softmax / epilogue
// (one thread per query row: row max / sum need no shuffles).
//
// Pipeline (r01 v6).  The UNet's 64x64-latent self-attention has d = 40: 160 tensor FLOPs per exponential, so the kernel is
// bound by the SFU (16 ex2/clk/SM), not the tensor pipe.  The first version kept ONE score tile in TMEM, so every key tile
// was a serial chain  Q.K^T -> softmax -> P.V -> next Q.K^T  and the softmax warps idled for the MMA round trip each tile
// (ncu: XU pipe 55 % busy, profiles/r01_attn_fwd40_ncu_v5.md).  Now S and P are double-buffered: Q.K^T of tile j+1 is
// issued BEFORE P.V of tile j, so scores are always waiting when the softmax warps finish a tile.  64-key tiles keep that
// within 256 TMEM columns (S0 | S1 | O) and 96 KB of smem for d <= 64, i.e. two CTAs (8 softmax warps) per SM.
#include "tc_common.cuh"

namespace comat {

constexpr int ATT_BM = 128;   // queries per CTA
constexpr int ATT_BN = 64;    // keys per tile
constexpr int ATT_THREADS = 192;
constexpr int ATT_DEFAULT_POLY = 0;   // exponential pairs (of every 8) on the FMA pipe in the long-sequence kernel; COMAT_ATTN_POLY overrides

struct AttnKP {
  int Lq, Lk, H, d;
  int n_kv_tiles;
  float scale_log2;    // scale * log2(e)
  float scale;
  void* out;           // (n, Lq, H*d) 16-bit
  float* probs;        // ((n - probs_b0)*H, Lq, Lk) fp32 or null (requires n_kv_tiles <= 2)
  int probs_b0;        // first sample whose probabilities are exported (attrcon: the conditional half of a CFG batch)
  int probs_staged;    // export through shared memory + one bulk copy per CTA (needs 16-byte aligned tiles)
  float* lse;          // (n*H, Lq) fp32 or null
  long long out_ld;    // H*d
  uint32_t idesc_qk, idesc_pv;
  int is_bf16;
  const int* kv_lens;  // per-sample number of valid keys (padding mask) or null
  int causal;          // key index <= query index (BLIP text decoder self-attention)
};

template <int D>
struct AttnCfg {
  static constexpr int NKC = (D + 63) / 64;                 // 64-wide d chunks of Q / K
  static constexpr int DN = (D + 15) / 16 * 16;             // P.V MMA N (output columns)
  static constexpr int KSTEPS_QK = (D + 15) / 16;           // 16-wide k-steps actually issued
  static constexpr int Q_BYTES = NKC * ATT_BM * 128;
  static constexpr int K_BYTES = NKC * ATT_BN * 128;
  static constexpr int V_BYTES = K_BYTES;                   // same [64 keys][NKC x 64] tile shape as K
  static constexpr int KV_STAGE = K_BYTES + V_BYTES;
  static constexpr int P_BYTES = ATT_BM * 128;              // one buffer: [128 rows][64 keys] 16-bit = one swizzle row per query
  static constexpr int STAGES = 3;                          // tile j+1 loads while tile j is in use and tile j-1 drains
  static_assert(Q_BYTES + 3 * KV_STAGE + 2 * ATT_BM * 128 + 256 <= 232448, "attention smem budget");
  static constexpr int BAR_OFF = Q_BYTES + STAGES * KV_STAGE + 2 * P_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256;               // dynamic smem is declared 1024-B aligned (checked in-kernel)
  static constexpr int TMEM_COLS = (128 + DN) <= 256 ? 256 : 512;
  static constexpr int O_COL = 128;                         // S buffers at columns [0,64) and [64,128), O tile at [128, 128+DN)
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}

template <int D, typename T>
__global__ void __launch_bounds__(ATT_THREADS, (D <= 64) ? 2 : 1)     // d <= 64: two CTAs per SM (smem 96 KB, TMEM 256 columns each)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnKP p) {
  pdl_trigger();
  using Cf = AttnCfg<D>;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // swizzled tiles need 1024-B alignment; d <= 64 leaves no slack (2 CTAs / SM)
  unsigned char* sQ = smem;
  unsigned char* sKV = smem + Cf::Q_BYTES;
  unsigned char* sP = sKV + Cf::STAGES * Cf::KV_STAGE;            // two P buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cf::BAR_OFF);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // [3]
  uint64_t* kv_empty = bars + 4;      // [3]
  uint64_t* s_full = bars + 7;        // [2] scores of tile j in S buffer j & 1
  uint64_t* p_full = bars + 9;        // [2] probabilities of tile j in P buffer j & 1 (128 arrivals)
  uint64_t* pv_done = bars + 11;      // [2] P.V of tile j complete: P buffer j & 1 reusable, O consistent
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int NT = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < Cf::STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); mbar_init(&pv_done[s], 1); }
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, Cf::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                    // everything above touched no global memory

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Cf::Q_BYTES);
      for (int c = 0; c < Cf::NKC; ++c) tma_load_3d(sQ + c * ATT_BM * 128, &tmQ, q_full, c * 64, h, b * p.Lq + m0);
      int stage = 0, phase = 0;
      for (int j = 0; j < NT; ++j) {
        mbar_wait_backoff(&kv_empty[stage], phase ^ 1);
        unsigned char* sK = sKV + stage * Cf::KV_STAGE;
        unsigned char* sV = sK + Cf::K_BYTES;
        mbar_expect_tx(&kv_full[stage], Cf::K_BYTES + Cf::V_BYTES);
        for (int c = 0; c < Cf::NKC; ++c) {
          tma_load_2d(sK + c * ATT_BN * 128, &tmK, &kv_full[stage], h * p.d + c * 64, b * p.Lk + j * ATT_BN);
          tma_load_2d(sV + c * ATT_BN * 128, &tmV, &kv_full[stage], h * p.d + c * 64, b * p.Lk + j * ATT_BN);
        }
        if (++stage == Cf::STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      // S_j = Q K_j^T into S buffer j & 1
      auto issue_qk = [&](int stage, int sbuf) {
        const uint32_t aK = smem_u32(sKV + stage * Cf::KV_STAGE);
#pragma unroll
        for (int ks = 0; ks < Cf::KSTEPS_QK; ++ks) {
          const uint32_t offq = (uint32_t)(ks / 4) * (ATT_BM * 128) + (uint32_t)(ks % 4) * 32;
          const uint32_t offk = (uint32_t)(ks / 4) * (ATT_BN * 128) + (uint32_t)(ks % 4) * 32;
          umma_f16(tmem_base + (uint32_t)(sbuf * ATT_BN), make_kmajor_sw128_desc(aQ + offq), make_kmajor_sw128_desc(aK + offk),
                   p.idesc_qk, ks > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[sbuf]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      int stage = 0, phase = 0;                  // ring position of tile j
      for (int j = 0; j < NT; ++j) {
        int nstage = stage + 1, nphase = phase;
        if (nstage == Cf::STAGES) { nstage = 0; nphase ^= 1; }
        if (j + 1 < NT) {
          // scores of the NEXT tile first: S buffer (j+1)&1 was last read by the softmax of tile j-1, whose p_full this
          // thread has already observed; the softmax warps find them ready when they finish tile j
          mbar_wait_backoff(&kv_full[nstage], nphase);
          tc_fence_after();
          issue_qk(nstage, (j + 1) & 1);
        }
        mbar_wait_backoff(&p_full[j & 1], (j >> 1) & 1);  // P_j is in smem
        tc_fence_after();
        const uint32_t aV = smem_u32(sKV + stage * Cf::KV_STAGE + Cf::K_BYTES);
        const uint32_t aPj = aP + (uint32_t)(j & 1) * Cf::P_BYTES;
#pragma unroll
        for (int ks = 0; ks < ATT_BN / 16; ++ks) {
          // V tile read MN-major: 16 key rows = 2048 B per k-step, d-panels ATT_BN*128 B apart
          umma_f16(tmem_base + Cf::O_COL, make_kmajor_sw128_desc(aPj + (uint32_t)ks * 32),
                   make_mnmajor_sw128_desc(aV + (uint32_t)ks * 2048, ATT_BN * 128), p.idesc_pv,
                   (j > 0 || ks > 0) ? 1u : 0u);      // O accumulates in TMEM across key tiles
        }
        umma_commit(&pv_done[j & 1]);
        umma_commit(&kv_empty[stage]);
        stage = nstage; phase = nphase;
      }
    }
    __syncwarp();
  } else {
    // ---------------- softmax + epilogue: thread = query row ----------------
    // O accumulates in TMEM across key tiles (the P.V MMAs run with accumulate on); the running row maximum is only
    // refreshed when some row of the warp grew by more than 2^8 (in the exp2 domain), so the TMEM read-scale-write of O
    // is rare after the first tiles and exp2 arguments stay <= 8.  The 64 scores of a row are read from TMEM once and
    // kept in registers for the max and the exponentials.
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const bool row_ok = (m0 + r) < p.Lq;
    float m_run = -INFINITY, l_run = 0.f;
    const int klen = (p.kv_lens != nullptr) ? min(p.Lk, p.kv_lens[b]) : p.Lk;
    const float sl2 = p.scale_log2;

    for (int j = 0; j < NT; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      const int ktile = min(ATT_BN, klen - j * ATT_BN);                       // warp-uniform
      const int kvalid = p.causal ? min(ktile, m0 + r - j * ATT_BN + 1) : ktile;   // per-row limit
      // warp-uniform: every row of this warp sees all keys of the tile -> no masking
      const bool full_w = (ktile == ATT_BN) && (!p.causal || (m0 + q4 * 32 - j * ATT_BN + 1 >= ATT_BN));
      float sv[ATT_BN];
#pragma unroll
      for (int c0 = 0; c0 < ATT_BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(trow + (uint32_t)(sb * ATT_BN + c0), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) sv[c0 + i] = __uint_as_float(v[i]);
      }
      tmem_ld_wait();
      if (!full_w) {
#pragma unroll
        for (int i = 0; i < ATT_BN; ++i)
          if (i >= kvalid) sv[i] = -INFINITY;
      }
      // row maximum with the three-input max of sm_100 (FMNMX3): 32 instead of 63 instructions per row and tile
      float mx[4] = {sv[0], sv[1], sv[2], sv[3]};
#pragma unroll
      for (int i = 4; i + 8 <= ATT_BN; i += 8) {
        mx[0] = fmax3(mx[0], sv[i], sv[i + 1]); mx[1] = fmax3(mx[1], sv[i + 2], sv[i + 3]);
        mx[2] = fmax3(mx[2], sv[i + 4], sv[i + 5]); mx[3] = fmax3(mx[3], sv[i + 6], sv[i + 7]);
      }
      mx[0] = fmax3(mx[0], sv[ATT_BN - 4], sv[ATT_BN - 3]); mx[1] = fmax3(mx[1], sv[ATT_BN - 2], sv[ATT_BN - 1]);
      const float mt = fmaxf(fmax3(mx[0], mx[1], mx[2]), mx[3]);
      if (j == 0) {
        m_run = mt;                               // O is not initialised yet: P.V_0 overwrites it
      } else {
        const bool grow = (mt - m_run) * sl2 > 8.f;         // false for mt = -inf or NaN-free equal maxima
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = fmaxf(m_run, mt);
          const float alpha = (m_new == -INFINITY) ? 1.f : fast_exp2((m_run - m_new) * sl2);
          l_run *= alpha;
          m_run = m_new;
          // P.V of tile j-1 may still be accumulating into O (its Q.K^T successor was issued ahead of it): wait for it.
          // P.V of tile j cannot start before this thread's p_full arrival below.
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < Cf::DN; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(trow + (uint32_t)(Cf::O_COL + c0), v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x16(trow + (uint32_t)(Cf::O_COL + c0), v);
          }
          tmem_st_wait();
        }
      }
      const float mneg = (m_run == -INFINITY) ? 0.f : m_run * sl2;
      // P buffer sb was last read by P.V of tile j-2
      if (j >= 2) mbar_wait(&pv_done[sb], ((j - 2) >> 1) & 1);
      // p = exp2(s*sl2 - m*sl2), row sum, 16-bit P into swizzled smem (K-major: one 128-byte row of 64 keys per query)
      // The exponentials are issued LA elements ahead of their consumers (row sum, 16-bit pack): with the consumer right
      // behind its ex2 the warp stalled ~20 cycles per pair on the SFU result while the SFU queue (8 issue cycles per
      // warp instruction) ran dry - XU pipe 55 % busy with two softmax warps per scheduler (profiles/r01_attn_fwd40_ncu_v5.md).
      constexpr int LA = 8;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      unsigned char* prow = sP + sb * Cf::P_BYTES + (r / 8) * 1024 + (r % 8) * 128;
#pragma unroll
      for (int i = 0; i < LA; ++i) sv[i] = fast_exp2(fmaf(sv[i], sl2, -mneg));
#pragma unroll
      for (int c0 = 0; c0 < ATT_BN; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          if (c0 + i + LA < ATT_BN) {
            sv[c0 + i + LA] = fast_exp2(fmaf(sv[c0 + i + LA], sl2, -mneg));
            sv[c0 + i + LA + 1] = fast_exp2(fmaf(sv[c0 + i + LA + 1], sl2, -mneg));
          }
          const float e0 = sv[c0 + i], e1 = sv[c0 + i + 1];
          ls[(i / 2) & 3] += e0 + e1;
          pk[i / 2] = pack2<T>(e0, e1);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cc = c0 / 8 + q;                      // 16-byte chunk (8 keys) inside the row
          *reinterpret_cast<uint4*>(prow + ((cc ^ (r % 8)) * 16)) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
      }
      l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      tc_fence_before();
      fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      mbar_arrive(&p_full[sb]);
    }
    const float m_used = m_run;
    // epilogue: O / l from TMEM
    mbar_wait(&pv_done[(NT - 1) & 1], ((NT - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    {
      T* op = reinterpret_cast<T*>(p.out) + ((size_t)b * p.Lq + m0 + r) * p.out_ld + (size_t)h * p.d;
#pragma unroll
      for (int c0 = 0; c0 < Cf::DN; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(trow + (uint32_t)(Cf::O_COL + c0), v);
        tmem_ld_wait();
        if (row_ok) {
          if constexpr (D % 8 == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              if (c0 + i < D) {
                uint4 u;
                u.x = pack2<T>(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l);
                u.y = pack2<T>(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l);
                u.z = pack2<T>(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l);
                u.w = pack2<T>(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l);
                *reinterpret_cast<uint4*>(op + c0 + i) = u;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < D) op[c0 + i] = from_f32<T>(__uint_as_float(v[i]) * inv_l);
          }
        }
      }
      if (row_ok && p.lse != nullptr) p.lse[((size_t)b * p.H + h) * p.Lq + m0 + r] = m_used * p.scale + logf(l_run);
    }
    if (p.probs != nullptr && b >= p.probs_b0) {
      // at most two key tiles: both score buffers are still in TMEM (columns = key index); write softmax(S) as fp32
      // ((n - b0)*H, Lq, Lk).  tcgen05.ld is .sync.aligned: the whole warp executes the loads, only the stores are predicated.
      // A CTA's 128 rows x Lk probabilities are ONE contiguous block of the export tensor: the rows are staged in the (now idle)
      // K/V + P shared-memory buffers - row pitch Lk = 77 words is odd, so the per-row writes are bank-conflict free - and leave
      // with one bulk copy.  (r01 wrote them straight from registers: 32 lanes x 32 different rows per store instruction, 4-byte
      // pieces of 32 sectors, 520 GB/s; profiles/r02_xattn_ncu.md.)
      float* tile_g = p.probs + (((size_t)(b - p.probs_b0) * p.H + h) * p.Lq + m0) * p.Lk;
      float* pp = tile_g + (size_t)r * p.Lk;
      const int rows_valid = min(ATT_BM, p.Lq - m0);
      constexpr int CAP = Cf::STAGES * Cf::KV_STAGE + 2 * Cf::P_BYTES;
      const bool staged = p.probs_staged && ((rows_valid * p.Lk) & 3) == 0 && ATT_BM * p.Lk * 4 <= CAP;     // CTA-uniform
      float* st = reinterpret_cast<float*>(sKV) + (size_t)r * p.Lk;
      const float mneg = m_used * p.scale_log2;
#pragma unroll 1
      for (int c0 = 0; c0 < p.Lk; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(trow + (uint32_t)c0, v);
        tmem_ld_wait();
        if (row_ok) {
          float* dst = staged ? st : pp;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.Lk) dst[c0 + i] = fast_exp2(__uint_as_float(v[i]) * p.scale_log2 - mneg) * inv_l;
        }
      }
      tc_fence_before();
      if (staged) {
        fence_proxy_async_smem();                              // generic-proxy smem writes -> visible to the bulk-copy engine
        asm volatile("bar.sync 1, 128;" ::: "memory");         // the four softmax warps
        if (warp == 2 && lane == 0) {
          bulk_s2g(tile_g, sKV, (uint32_t)(rows_valid * p.Lk * 4));
          bulk_commit();
          bulk_wait_read<0>();                                 // shared memory stays valid until the copy has read it
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cf::TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------------------
// Long-sequence variant for d <= 64 (the UNet's 64x64-latent self-attention: 4096 x 4096 per head at d = 40; BLIP ViT 577 x 577
// at d = 64): THREE CTAs per SM instead of two.
// Why (profiles/r02_attn_fwd_analysis.md): at d = 40 the kernel is bound by the exponentials (16 MUFU.EX2 / clk / SM), not the
// tensor pipe (160 tensor FLOPs per exponential).  The two-CTA kernel reached 52 % of the MUFU rate: ncu's per-instruction stall
// samples put 37 % of a softmax warp's time outside its arithmetic block (score barrier, TMEM load latency, proxy fence + 128
// mbarrier arrivals per tile) and the two warps a scheduler holds were in their arithmetic blocks at the same time, while the
// bare instruction mix sustains 86 % of the MUFU rate at the same occupancy (tools/bench_softmax_loop.cu).  More independent
// warps per scheduler is what hides those phases.
// What makes three CTAs fit:
//   * ONE score buffer (64 TMEM columns) + O: 128 columns per CTA.  The softmax warps release the buffer as soon as the scores
//     are in registers (s_empty), so Q.K^T of tile j+1 still runs underneath the exponentials of tile j;
//   * ONE P buffer and separate two-stage K and V rings (K of tile j+1 is needed at the START of tile j, V of tile j at its
//     END, so two stages each give both a lead of two tiles): 64 KB of shared memory per CTA;
//   * one mbarrier arrival per WARP (fence, __syncwarp, elected lane) instead of one per thread.
// EXP_POLY of every 8 column pairs take their exponential on the FMA pipe (packed f32x2 Cody-Waite split + degree-3 minimax
// polynomial, 7.5e-5 relative: below the 16-bit rounding of P) instead of MUFU, which moves the bound from the MUFU pipe
// towards the issue slots (17.9 vs 13.7 elements / clk / SM in the micro-benchmark).
// No probability export, no causal mask (those calls have <= 2 key tiles and use the kernel above).
// ---------------------------------------------------------------------------------------------------------------
template <int D>
struct AttnLCfg {
  static_assert(D <= 64, "one 64-wide d chunk");
  static constexpr int DN = (D + 15) / 16 * 16;
  static constexpr int KSTEPS_QK = (D + 15) / 16;
  static constexpr int Q_BYTES = ATT_BM * 128;
  static constexpr int KV_BYTES = ATT_BN * 128;              // one K or V tile
  static constexpr int P_BYTES = ATT_BM * 128;
  static constexpr int K_OFF = Q_BYTES, V_OFF = K_OFF + 2 * KV_BYTES, P_OFF = V_OFF + 2 * KV_BYTES;
  static constexpr int BAR_OFF = P_OFF + P_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256;
  static constexpr int TMEM_COLS = 128;                      // S at [0, 64), O at [64, 64 + DN)
  static constexpr int O_COL = 64;
};

__device__ __forceinline__ uint64_t pk_f32x2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void un_f32x2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// 2^x for two arguments x <= 0 on the FMA / ALU pipes: x = n + f with n = round(x) taken from the low mantissa bits of
// x + 1.5 * 2^23, 2^f by a degree-3 minimax polynomial on [-0.5, 0.5] (7.5e-5 relative), 2^n added into the exponent field
__device__ __forceinline__ void ex2_poly_pair(float x0, float x1, float& o0, float& o1) {
  const uint64_t x = pk_f32x2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const uint64_t t = fadd2(x, pk_f32x2(12582912.f, 12582912.f));
  const uint64_t n = fadd2(t, pk_f32x2(-12582912.f, -12582912.f));
  const uint64_t f = ffma2(n, pk_f32x2(-1.f, -1.f), x);
  uint64_t q = ffma2(f, pk_f32x2(0.055171653628349304f, 0.055171653628349304f), pk_f32x2(0.2426111251115799f, 0.2426111251115799f));
  q = ffma2(q, f, pk_f32x2(0.6932609677314758f, 0.6932609677314758f));
  q = ffma2(q, f, pk_f32x2(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, t0, t1;
  un_f32x2(q, q0, q1); un_f32x2(t, t0, t1);
  o0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  o1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

constexpr int ATT_L_THREADS = 256;   // warpgroup 0: TMA producer, MMA issuer (+ two idle warps); warpgroup 1: softmax
template <int D, typename T, int EXP_POLY>
__global__ void __launch_bounds__(ATT_L_THREADS, 3)
attn_fwd_long_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnKP p) {
  pdl_trigger();
  using Cf = AttnLCfg<D>;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  unsigned char* sQ = smem;
  unsigned char* sK = smem + Cf::K_OFF;
  unsigned char* sV = smem + Cf::V_OFF;
  unsigned char* sP = smem + Cf::P_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cf::BAR_OFF);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;        // [2]
  uint64_t* v_empty = bars + 7;       // [2]
  uint64_t* s_full = bars + 9;        // scores of tile j in TMEM
  uint64_t* s_empty = bars + 10;      // ... and in the softmax warps' registers (4 arrivals): the buffer may be overwritten
  uint64_t* p_full = bars + 11;       // probabilities of tile j in smem (4 arrivals)
  uint64_t* pv_done = bars + 12;      // P.V of tile j complete: P buffer reusable, O consistent
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int NT = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(s_empty, 4); mbar_init(p_full, 4); mbar_init(pv_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, Cf::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  // Register budget: the CTA is launched with 80 registers per thread (3 CTAs x 256 threads x 80 = 61440 of the SM's 65536: the
  // register file is allocated per FOUR warps, so a 192-thread CTA is charged for 256 threads anyway).  The producer warpgroup
  // returns its share and the softmax warpgroup takes it: 40 / 120 registers per thread - the 64 scores of a row stay in
  // registers.
  // (each role's code sits wholly inside the branch that opens with its setmaxnreg: ptxas allocates registers per region)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ---------------- TMA producer: K and V rings in the order the MMA warp consumes them: K0, K1, V0, K2, V1, ...
    if (lane == 0) {
      mbar_expect_tx(q_full, Cf::Q_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, 0, h, b * p.Lq + m0);
      auto load_k = [&](int j) {
        const int s = j & 1;
        mbar_wait_backoff(&k_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[s], Cf::KV_BYTES);
        tma_load_2d(sK + s * Cf::KV_BYTES, &tmK, &k_full[s], h * p.d, b * p.Lk + j * ATT_BN);
      };
      auto load_v = [&](int j) {
        const int s = j & 1;
        mbar_wait_backoff(&v_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[s], Cf::KV_BYTES);
        tma_load_2d(sV + s * Cf::KV_BYTES, &tmV, &v_full[s], h * p.d, b * p.Lk + j * ATT_BN);
      };
      load_k(0);
      for (int j = 0; j < NT; ++j) {
        if (j + 1 < NT) load_k(j + 1);
        load_v(j);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer
    if (lane == 0) {
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait_backoff(&k_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + s * Cf::KV_BYTES);
#pragma unroll
        for (int ks = 0; ks < Cf::KSTEPS_QK; ++ks)
          umma_f16(tmem_base, make_kmajor_sw128_desc(aQ + (uint32_t)ks * 32), make_kmajor_sw128_desc(aK + (uint32_t)ks * 32),
                   p.idesc_qk, ks > 0 ? 1u : 0u);
        umma_commit(s_full);
        umma_commit(&k_empty[s]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < NT; ++j) {
        if (j + 1 < NT) {
          mbar_wait_backoff(s_empty, j & 1);           // the scores of tile j are in registers
          issue_qk(j + 1);                             // runs underneath the exponentials of tile j
        }
        mbar_wait_backoff(p_full, j & 1);
        mbar_wait_backoff(&v_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aV = smem_u32(sV + (j & 1) * Cf::KV_BYTES);
#pragma unroll
        for (int ks = 0; ks < ATT_BN / 16; ++ks)
          umma_f16(tmem_base + Cf::O_COL, make_kmajor_sw128_desc(aP + (uint32_t)ks * 32),
                   make_mnmajor_sw128_desc(aV + (uint32_t)ks * 2048, ATT_BN * 128), p.idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(pv_done);
        umma_commit(&v_empty[j & 1]);
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    // ---------------- softmax + epilogue: thread = query row
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const bool row_ok = (m0 + r) < p.Lq;
    float m_run = -INFINITY, l_run = 0.f;
    const int klen = (p.kv_lens != nullptr) ? min(p.Lk, p.kv_lens[b]) : p.Lk;
    const float sl2 = p.scale_log2;
    unsigned char* prow = sP + (r / 8) * 1024 + (r % 8) * 128;

    for (int j = 0; j < NT; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float sv[ATT_BN];
#pragma unroll
      for (int c0 = 0; c0 < ATT_BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(trow + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) sv[c0 + i] = __uint_as_float(v[i]);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);               // Q.K^T of tile j+1 may overwrite the score buffer
      const int ktile = min(ATT_BN, klen - j * ATT_BN);   // warp-uniform
      if (ktile < ATT_BN) {
#pragma unroll
        for (int i = 0; i < ATT_BN; ++i)
          if (i >= ktile) sv[i] = -INFINITY;
      }
      float mx[4] = {sv[0], sv[1], sv[2], sv[3]};
#pragma unroll
      for (int i = 4; i + 8 <= ATT_BN; i += 8) {
        mx[0] = fmax3(mx[0], sv[i], sv[i + 1]); mx[1] = fmax3(mx[1], sv[i + 2], sv[i + 3]);
        mx[2] = fmax3(mx[2], sv[i + 4], sv[i + 5]); mx[3] = fmax3(mx[3], sv[i + 6], sv[i + 7]);
      }
      mx[0] = fmax3(mx[0], sv[ATT_BN - 4], sv[ATT_BN - 3]); mx[1] = fmax3(mx[1], sv[ATT_BN - 2], sv[ATT_BN - 1]);
      const float mt = fmaxf(fmax3(mx[0], mx[1], mx[2]), mx[3]);
      if (j == 0) {
        m_run = mt;                                       // O is not initialised yet: P.V_0 overwrites it
      } else {
        // single P buffer: P.V of tile j-1 has to be done before this tile's probabilities are written; the same wait makes O
        // consistent for the (rare) rescale below.  It was issued a score-load + row-maximum ago.
        mbar_wait(pv_done, (j - 1) & 1);
        const bool grow = (mt - m_run) * sl2 > 8.f;
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = fmaxf(m_run, mt);
          const float alpha = (m_new == -INFINITY) ? 1.f : fast_exp2((m_run - m_new) * sl2);
          l_run *= alpha;
          m_run = m_new;
          tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < Cf::DN; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(trow + (uint32_t)(Cf::O_COL + c0), v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x16(trow + (uint32_t)(Cf::O_COL + c0), v);
          }
          tmem_st_wait();
        }
      }
      const float mneg = (m_run == -INFINITY) ? 0.f : m_run * sl2;
      // p = exp2(s * sl2 - m * sl2): packed f32x2 scale, EXP_POLY of every 8 pairs on the FMA pipe, the others on MUFU; row sum
      // in two packed accumulators; 16-bit P into swizzled smem (K-major: one 128-byte row of 64 keys per query)
      const uint64_t sl2x2 = pk_f32x2(sl2, sl2), mnegx2 = pk_f32x2(-mneg, -mneg);
      uint64_t acc0 = pk_f32x2(0.f, 0.f), acc1 = acc0;
#pragma unroll
      for (int c0 = 0; c0 < ATT_BN; c0 += 8) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const int pi = (c0 + i) / 2;                    // pair index 0..31
          float x0, x1, e0, e1;
          un_f32x2(ffma2(pk_f32x2(sv[c0 + i], sv[c0 + i + 1]), sl2x2, mnegx2), x0, x1);
          const bool poly = (EXP_POLY == 2 && (pi & 3) == 3) || (EXP_POLY == 3 && ((pi & 7) == 1 || (pi & 7) == 4 || (pi & 7) == 6)) ||
                            (EXP_POLY == 4 && (pi & 1) == 1);
          if (poly) ex2_poly_pair(x0, x1, e0, e1);
          else { e0 = fast_exp2(x0); e1 = fast_exp2(x1); }
          if (pi & 1) acc1 = fadd2(acc1, pk_f32x2(e0, e1)); else acc0 = fadd2(acc0, pk_f32x2(e0, e1));
          pk[i / 2] = pack2<T>(e0, e1);
        }
        *reinterpret_cast<uint4*>(prow + (((c0 / 8) ^ (r % 8)) * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      {
        float a0, a1;
        un_f32x2(fadd2(acc0, acc1), a0, a1);
        l_run += a0 + a1;
      }
      tc_fence_before();             // (a rescale's tcgen05.st to O is ordered before the P.V the arrival below releases)
      fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l from TMEM
    mbar_wait(pv_done, (NT - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    T* op = reinterpret_cast<T*>(p.out) + ((size_t)b * p.Lq + m0 + r) * p.out_ld + (size_t)h * p.d;
#pragma unroll
    for (int c0 = 0; c0 < Cf::DN; c0 += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(trow + (uint32_t)(Cf::O_COL + c0), v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          if (c0 + i < D) {
            uint4 u;
            u.x = pack2<T>(__uint_as_float(v[i]) * inv_l, __uint_as_float(v[i + 1]) * inv_l);
            u.y = pack2<T>(__uint_as_float(v[i + 2]) * inv_l, __uint_as_float(v[i + 3]) * inv_l);
            u.z = pack2<T>(__uint_as_float(v[i + 4]) * inv_l, __uint_as_float(v[i + 5]) * inv_l);
            u.w = pack2<T>(__uint_as_float(v[i + 6]) * inv_l, __uint_as_float(v[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(op + c0 + i) = u;
          }
        }
      }
    }
    if (row_ok && p.lse != nullptr) p.lse[((size_t)b * p.H + h) * p.Lq + m0 + r] = m_run * p.scale + logf(l_run);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cf::TMEM_COLS); }
}

template <int D, typename T, int EXP_POLY>
static int launch_attn_long(const CUtensorMap* maps, const AttnKP& kp, dim3 grid, cudaStream_t st) {
  using Cf = AttnLCfg<D>;
  static bool configured = false;
  if (!configured) {
    COMAT_CUDA(cudaFuncSetAttribute(attn_fwd_long_kernel<D, T, EXP_POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::TOTAL));
    // three CTAs need 3 x 65 KB of shared memory: ask for the largest carve-out (the default heuristic left room for two)
    COMAT_CUDA(cudaFuncSetAttribute(attn_fwd_long_kernel<D, T, EXP_POLY>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  launch_k(attn_fwd_long_kernel<D, T, EXP_POLY>, grid, ATT_L_THREADS, Cf::TOTAL, st, maps[0], maps[1], maps[2], kp);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

template <int D, typename T>
static int launch_attn(const CUtensorMap* maps, const AttnKP& kp, dim3 grid, cudaStream_t st) {
  using Cf = AttnCfg<D>;
  static bool configured = false;
  if (!configured) {
    COMAT_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::TOTAL));
    configured = true;
  }
  launch_k(attn_fwd_kernel<D, T>, grid, ATT_THREADS, Cf::TOTAL, st, maps[0], maps[1], maps[2], kp);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

}  // namespace comat
using namespace comat;

extern "C" size_t comat_attention_workspace_bytes(int n, int Lk, int H, int d) {
  (void)n; (void)Lk; (void)H; (void)d;
  return 256;      // no scratch is needed any more (V is read in place); kept for ABI stability
}

extern "C" int comat_attention_fwd_strided(const void* q, const void* k, const void* v, void* out, float* probs, float* lse,
                                           void* workspace, int n, int Lq, int Lk, int H, int d, long long q_ld, long long k_ld,
                                           long long v_ld, float scale, int dtype, const int* kv_lens, int causal,
                                           int probs_first_sample, void* stream) {
  if (!q || !k || !v || !out || !workspace || n <= 0 || Lq <= 0 || Lk <= 0 || H <= 0) return COMAT_ERR_INVALID;
  if (probs_first_sample < 0 || probs_first_sample >= n) return COMAT_ERR_INVALID;
  if (q_ld < (long long)H * d || k_ld < (long long)H * d || v_ld < (long long)H * d || (q_ld % 8) || (k_ld % 8) || (v_ld % 8)) return COMAT_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) || (reinterpret_cast<uintptr_t>(v) & 15)) return COMAT_ERR_INVALID;
  if (d != 40 && d != 64 && d != 80 && d != 128 && d != 160 && d != 32 && d != 16) return COMAT_ERR_UNSUPPORTED;
  if (dtype != COMAT_F16 && dtype != COMAT_BF16) return COMAT_ERR_UNSUPPORTED;
  if (probs && Lk > 2 * ATT_BN) return COMAT_ERR_UNSUPPORTED;     // both score buffers must still hold the whole row
  if (((H * d) % 8) != 0 || (d % 8) != 0) return COMAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  AttnKP kp;
  memset(&kp, 0, sizeof(kp));
  kp.Lq = Lq; kp.Lk = Lk; kp.H = H; kp.d = d; kp.n_kv_tiles = (Lk + ATT_BN - 1) / ATT_BN;
  kp.scale = scale; kp.scale_log2 = scale * 1.4426950408889634f;
  kp.kv_lens = kv_lens; kp.causal = causal;
  kp.probs_staged = (probs != nullptr && (((long long)Lq * Lk) & 3) == 0 && (reinterpret_cast<uintptr_t>(probs) & 15) == 0) ? 1 : 0;
  kp.out = out; kp.probs = probs; kp.probs_b0 = probs_first_sample; kp.lse = lse; kp.out_ld = (long long)H * d; kp.is_bf16 = dtype == COMAT_BF16;
  const int fmt = kp.is_bf16 ? 1 : 0;
  const int DN = (d + 15) / 16 * 16;
  kp.idesc_qk = make_idesc_f16(ATT_BM, ATT_BN, fmt);
  kp.idesc_pv = make_idesc_f16(ATT_BM, DN, fmt, 0, 1);        // B (= V) MN-major
  CUtensorMap maps[3];
  {
    const uint64_t dq[3] = {(uint64_t)d, (uint64_t)H, (uint64_t)n * Lq};
    const uint64_t sq[2] = {(uint64_t)d * 2, (uint64_t)q_ld * 2};
    const uint32_t bq[3] = {64, 1, (uint32_t)ATT_BM};
    if (!make_tmap_16bit(&maps[0], q, 3, dq, sq, bq)) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
    // K and V: 2-D maps over (H*d, rows) - a tile is the 64-column window that STARTS at the head's first column.  For d = 40
    // the window also holds 24 columns of the next head; they meet the zero-filled columns 40..63 of the Q tile in Q.K^T and
    // land in output columns of P.V that are never stored.  (r02 first used the 3-D {d, H, rows} map for K / V too: every
    // 80-byte row of a box then ends out of bounds and the TMA unit issued 4.3 L2 requests per row - 553 per key tile - which,
    // not the exponentials, set the pace of the d = 40 kernel: profiles/r02_attn_fwd_analysis.md.)
    const uint64_t dk[2] = {(uint64_t)H * d, (uint64_t)n * Lk};
    const uint64_t sk[1] = {(uint64_t)k_ld * 2};
    const uint64_t sv[1] = {(uint64_t)v_ld * 2};
    const uint32_t bk[2] = {64, (uint32_t)ATT_BN};
    if (!make_tmap_16bit(&maps[1], k, 2, dk, sk, bk)) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
    if (!make_tmap_16bit(&maps[2], v, 2, dk, sv, bk)) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
  }
  dim3 grid((Lq + ATT_BM - 1) / ATT_BM, H, n);
  // long sequences at d = 40 / 64 without probability export / causal mask: the three-CTAs-per-SM kernel (0.370 vs 0.409 ms on the
  // 64x64-latent self-attention shape, profiles/r02_attn_fwd_analysis.md).  COMAT_ATTN_LONG=0 keeps the two-CTA kernel;
  // COMAT_ATTN_POLY = 0 | 3 picks how many of every 8 exponential pairs run on the FMA pipe (3: 0.366 ms; default 0 = all MUFU).
  static int long_mode = -1, poly_mode = -1;
  if (long_mode < 0) { const char* e = getenv("COMAT_ATTN_LONG"); long_mode = (e && e[0] == '0') ? 0 : 1; }
  if (poly_mode < 0) { const char* e = getenv("COMAT_ATTN_POLY"); poly_mode = e ? atoi(e) : ATT_DEFAULT_POLY; }
  if (long_mode && (d == 40 || d == 64) && !probs && !causal && kp.n_kv_tiles >= 3) {
#define ATT_LCASE(DD)                                                                                                              \
  case DD:                                                                                                                         \
    if (poly_mode == 3) return kp.is_bf16 ? launch_attn_long<DD, __nv_bfloat16, 3>(maps, kp, grid, st) : launch_attn_long<DD, __half, 3>(maps, kp, grid, st); \
    return kp.is_bf16 ? launch_attn_long<DD, __nv_bfloat16, 0>(maps, kp, grid, st) : launch_attn_long<DD, __half, 0>(maps, kp, grid, st);
    switch (d) {
      ATT_LCASE(40) ATT_LCASE(64)
    }
#undef ATT_LCASE
  }
#define ATT_CASE(DD)                                                                      \
  case DD:                                                                                \
    return kp.is_bf16 ? launch_attn<DD, __nv_bfloat16>(maps, kp, grid, st) : launch_attn<DD, __half>(maps, kp, grid, st);
  switch (d) {
    ATT_CASE(16) ATT_CASE(32) ATT_CASE(40) ATT_CASE(64) ATT_CASE(80) ATT_CASE(128) ATT_CASE(160)
  }
#undef ATT_CASE
  return COMAT_ERR_UNSUPPORTED;
}

extern "C" int comat_attention_fwd(const void* q, const void* k, const void* v, void* out, float* probs, float* lse, void* workspace,
                                   int n, int Lq, int Lk, int H, int d, float scale, int dtype, const int* kv_lens, int causal,
                                   void* stream) {
  const long long ld = (long long)H * d;
  return comat_attention_fwd_strided(q, k, v, out, probs, lse, workspace, n, Lq, Lk, H, d, ld, ld, ld, scale, dtype, kv_lens, causal, 0, stream);
}
