// Fused attention backward on tcgen05 (flash-style recomputation, no P materialised in HBM).
//
//   P = exp(scale * Q K^T - lse),  dP = dO V^T (+ dP_ext),  D_i = sum_j P_ij dP_ij,  dS = P o (dP - D)
//   dV = P^T dO,   dK = scale * dS^T Q,   dQ = scale * dS K
//
// dP_ext is the gradient that the attention-map loss sends into the exported cross-attention probabilities
// (the reference differentiates through the tensor its hook stored: attn_utils/tc_attn_utils.py:142-143, SURVEY "hard parts").
//
// Two launches of one templated kernel, no atomics, bit-reproducible:
//   MODE 0 (dQ)    : CTA = 128 queries, loops over key tiles   : S = Q K^T, dP = dO V^T  -> dS (smem) -> dQ += dS K
//   MODE 1 (dK,dV) : CTA = 128 keys,    loops over query tiles : S^T = K Q^T, dP^T = V dO^T -> P^T, dS^T (smem)
//                                                                 -> dV += P^T dO,  dK += dS^T Q
// No transposed copies exist anywhere: a streamed [64 rows][d] tile lands in smem as 128-byte-swizzled rows, which is at once
// the canonical K-major operand for the S / dP products (reduction over d) and the canonical MN-major operand for the
// accumulating products (reduction over the 64 rows) - the second use only flips the major bit of the instruction descriptor.
// S / dP (64 columns each) and the dQ / dK / dV accumulators live in TMEM; for d <= 64 that is <= 256 columns and <= 113 KB of
// smem, so two CTAs share an SM and one CTA's softmax-backward arithmetic overlaps the other's MMAs.  The elementwise softmax
// backward is done by 128 row-owning threads.
#include "tc_common.cuh"

namespace comat {

constexpr int AB_ROWS = 128;
constexpr int AB_ROW_WARPS = 8;                 // two row warps per TMEM lane quarter, each owning 32 of a tile's 64 columns
constexpr int AB_THREADS = 64 + 32 * AB_ROW_WARPS;

struct AttnBwdKP {
  int Lq, Lk, H, d;
  int Lq_pad, Lk_pad;
  int n_inner;
  float scale, scale_log2;
  const float* lse_pad;   // (n*H, Lq_pad), lse * log2(e), padded with +1e30
  const float* D_pad;     // (n*H, Lq_pad), padded with 0
  const float* dp_ext;    // ((n - dp_b0)*H, Lq, Lk) fp32 or null
  int dp_b0;              // first sample that has an external dP (the conditional half of a CFG batch); earlier samples: none
  void* out0;             // MODE 0: dQ (n, Lq, H*d) ; MODE 1: dK (n, Lk, H*d)
  void* out1;             // MODE 1: dV
  long long out_ld;
  uint32_t idesc_s, idesc_acc;
  const int* kv_lens;
  int causal;
  int nsplit;             // MODE 1 only: the query range is split over nsplit CTAs per key tile (cross-attention: 77 keys = ONE key tile,
                          // so n*H CTAs would walk all query tiles serially); partial dK / dV are added into fp32 scratch with atomics
  float* acc32_0;         // fp32 (n, Lk, H*d) scratch for dK (x scale) when nsplit > 1
  float* acc32_1;         // ... for dV
};

template <int D, int MODE>
struct ABCfg {
  static constexpr int NKC = (D + 63) / 64;
  static constexpr int DN = (D + 15) / 16 * 16;
  static constexpr int KSTEPS = (D + 15) / 16;
  static constexpr int BY = 64;                                    // inner tile width (keys in MODE 0, queries in MODE 1)
  static constexpr int NP = (MODE == 0) ? 1 : 2;                   // accumulators = smem A tiles produced by the row threads
  static constexpr int RES_BYTES = 2 * NKC * AB_ROWS * 128;
  static constexpr int NAT_BYTES = NKC * BY * 128;                 // one streamed tile: NKC panels of [BY rows][128 B]
  static constexpr int VEC_BYTES = (MODE == 1) ? 2 * BY * 4 : 0;   // lse / D slices of the inner tile
  static constexpr int STAGE_BYTES = ((2 * NAT_BYTES + VEC_BYTES + 1023) / 1024) * 1024;
  static constexpr int PD_BYTES = NP * AB_ROWS * 128;
  static constexpr int FIXED = RES_BYTES + PD_BYTES + 256 + 1024;
  static constexpr int COL_S = 0, COL_DP = BY, COL_ACC0 = 2 * BY, COL_ACC1 = 2 * BY + DN;
  static constexpr int TMEM_COLS = (2 * BY + NP * DN <= 256) ? 256 : 512;
  // two CTAs per SM need 2 x (TOTAL + 1 KB reserved) <= 228 KB and 2 x 256 TMEM columns
  static constexpr int OCC = (TMEM_COLS == 256 && FIXED + STAGE_BYTES <= 113 * 1024) ? 2 : 1;
  static constexpr int BUDGET = (OCC == 2) ? 113 * 1024 : 225 * 1024;
  static constexpr int STAGES = (FIXED + 2 * STAGE_BYTES <= BUDGET) ? 2 : 1;
  static constexpr int BAR_OFF = RES_BYTES + STAGES * STAGE_BYTES + PD_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

template <typename T>
__device__ __forceinline__ uint32_t ab_pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t ab_pack2<__half>(float a, float b) { const __half2 t = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&t); }
template <>
__device__ __forceinline__ uint32_t ab_pack2<__nv_bfloat16>(float a, float b) { const __nv_bfloat162 t = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&t); }

// 16-byte chunk `cc` (8 elements) of row r in a [128 x 64] K-major 128B-swizzled half tile
__device__ __forceinline__ unsigned char* sw128_chunk(unsigned char* half_base, int r, int cc) {
  return half_base + (r / 8) * 1024 + (r % 8) * 128 + ((cc ^ (r % 8)) * 16);
}

template <int D, int MODE, typename T>
__global__ void __launch_bounds__(AB_THREADS, ABCfg<D, MODE>::OCC)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmA1,   // resident natural 1: MODE0 Q   | MODE1 K      (box 128 rows)
                const __grid_constant__ CUtensorMap tmA2,   // resident natural 2: MODE0 dO  | MODE1 V
                const __grid_constant__ CUtensorMap tmB1,   // streamed natural 1: MODE0 K   | MODE1 Q      (box BY rows)
                const __grid_constant__ CUtensorMap tmB2,   // streamed natural 2: MODE0 V   | MODE1 dO
                const AttnBwdKP p) {
  pdl_trigger();
  using Cf = ABCfg<D, MODE>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sRes = smem;
  unsigned char* sStage = smem + Cf::RES_BYTES;
  unsigned char* sPD = sStage + Cf::STAGES * Cf::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cf::BAR_OFF);
  uint64_t* res_full = bars;
  uint64_t* st_full = bars + 1;     // [2]
  uint64_t* st_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* acc_full = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
  uint64_t* pd_free = bars + 9;     // accumulating MMAs of the previous tile have read the P / dS tiles in smem

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsplit = (MODE == 1 && p.nsplit > 1) ? p.nsplit : 1;
  const int x0 = (blockIdx.x / nsplit) * AB_ROWS;          // first query (MODE 0) / key (MODE 1) of this CTA
  const int h = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.H + h;
  const float* dp_ext = (p.dp_ext != nullptr && b >= p.dp_b0) ? p.dp_ext : nullptr;     // CTA-uniform
  const int bhx = (b - p.dp_b0) * p.H + h;                                               // row block of this (sample, head) in dp_ext
  const int split = blockIdx.x % nsplit;
  const int it0 = (int)((long long)p.n_inner * split / nsplit);                  // this CTA's inner tiles [it0, it0 + NI)
  const int NI = (int)((long long)p.n_inner * (split + 1) / nsplit) - it0;
  const int Lx = (MODE == 0) ? p.Lq : p.Lk;     // rows of the outer dimension
  const int Ly = (MODE == 0) ? p.Lk : p.Lq;     // inner dimension

  if (warp == 0 && lane == 0) {
    mbar_init(res_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&st_full[s], 1); mbar_init(&st_empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 32 * AB_ROW_WARPS); mbar_init(acc_full, 1); mbar_init(pd_free, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, Cf::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                                    // everything above touched no global memory

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(res_full, Cf::RES_BYTES);
      for (int c = 0; c < Cf::NKC; ++c) {
        tma_load_3d(sRes + c * AB_ROWS * 128, &tmA1, res_full, c * 64, h, b * Lx + x0);
        tma_load_3d(sRes + (Cf::NKC + c) * AB_ROWS * 128, &tmA2, res_full, c * 64, h, b * Lx + x0);
      }
      int stage = 0, phase = 0;
      for (int it = 0; it < NI; ++it) {
        mbar_wait(&st_empty[stage], phase ^ 1);
        unsigned char* st = sStage + stage * Cf::STAGE_BYTES;
        mbar_expect_tx(&st_full[stage], 2 * Cf::NAT_BYTES + Cf::VEC_BYTES);
        const int y0 = (it0 + it) * Cf::BY;
        for (int c = 0; c < Cf::NKC; ++c) {
          // streamed tiles: 2-D maps, the 64-column window that starts at the head's first column (see nat_st below)
          tma_load_2d(st + c * Cf::BY * 128, &tmB1, &st_full[stage], h * p.d + c * 64, b * Ly + y0);
          tma_load_2d(st + Cf::NAT_BYTES + c * Cf::BY * 128, &tmB2, &st_full[stage], h * p.d + c * 64, b * Ly + y0);
        }
        if (MODE == 1) {
          unsigned char* vec = st + 2 * Cf::NAT_BYTES;
          bulk_g2s(vec, p.lse_pad + (size_t)bh * p.Lq_pad + y0, Cf::BY * 4, &st_full[stage]);
          bulk_g2s(vec + Cf::BY * 4, p.D_pad + (size_t)bh * p.Lq_pad + y0, Cf::BY * 4, &st_full[stage]);
        }
        if (++stage == Cf::STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t aR1 = smem_u32(sRes), aR2 = aR1 + Cf::NKC * AB_ROWS * 128, aPD = smem_u32(sPD);
      mbar_wait(res_full, 0);
      // S = A1 B1^T and dP = A2 B2^T of the inner tile held in `stage`
      auto issue_scores = [&](int stage) {
        const uint32_t aB1 = smem_u32(sStage + stage * Cf::STAGE_BYTES), aB2 = aB1 + Cf::NAT_BYTES;
#pragma unroll
        for (int ks = 0; ks < Cf::KSTEPS; ++ks) {
          const uint32_t offa = (uint32_t)(ks / 4) * (AB_ROWS * 128) + (uint32_t)(ks % 4) * 32;
          const uint32_t offb = (uint32_t)(ks / 4) * (Cf::BY * 128) + (uint32_t)(ks % 4) * 32;
          umma_f16(tmem_base + Cf::COL_S, make_kmajor_sw128_desc(aR1 + offa), make_kmajor_sw128_desc(aB1 + offb), p.idesc_s, ks > 0);
        }
#pragma unroll
        for (int ks = 0; ks < Cf::KSTEPS; ++ks) {
          const uint32_t offa = (uint32_t)(ks / 4) * (AB_ROWS * 128) + (uint32_t)(ks % 4) * 32;
          const uint32_t offb = (uint32_t)(ks / 4) * (Cf::BY * 128) + (uint32_t)(ks % 4) * 32;
          umma_f16(tmem_base + Cf::COL_DP, make_kmajor_sw128_desc(aR2 + offa), make_kmajor_sw128_desc(aB2 + offb), p.idesc_s, ks > 0);
        }
        umma_commit(s_full);
      };
      // accumulating products of inner tile `it`: B = a streamed natural tile read MN-major (reduction over its BY rows, 16 rows =
      // 2048 B per k-step, d-panels BY*128 B apart)
      auto issue_acc = [&](int stage, int it) {
        const uint32_t aB1 = smem_u32(sStage + stage * Cf::STAGE_BYTES), aB2 = aB1 + Cf::NAT_BYTES;
        const uint32_t aAcc0 = (MODE == 0) ? aB1 : aB2;      // MODE 0: dQ += dS K      MODE 1: dV += P^T dO
#pragma unroll
        for (int ks = 0; ks < Cf::BY / 16; ++ks) {
          umma_f16(tmem_base + Cf::COL_ACC0, make_kmajor_sw128_desc(aPD + (uint32_t)ks * 32),
                   make_mnmajor_sw128_desc(aAcc0 + (uint32_t)ks * 2048, Cf::BY * 128), p.idesc_acc, (it > 0 || ks > 0) ? 1u : 0u);
        }
        if (MODE == 1) {
#pragma unroll
          for (int ks = 0; ks < Cf::BY / 16; ++ks) {
            umma_f16(tmem_base + Cf::COL_ACC1, make_kmajor_sw128_desc(aPD + (uint32_t)(AB_ROWS * 128) + (uint32_t)ks * 32),
                     make_mnmajor_sw128_desc(aB1 + (uint32_t)ks * 2048, Cf::BY * 128), p.idesc_acc,
                     (it > 0 || ks > 0) ? 1u : 0u);          // dK += dS^T Q
          }
        }
        umma_commit(&st_empty[stage]);
        umma_commit(pd_free);
      };
      int stage = 0, phase = 0;
      if constexpr (Cf::STAGES >= 2) {
        // Software-pipelined order (r02): the scores of tile it+1 are issued BEFORE the accumulating products of tile it, so the row
        // threads start on tile it+1 while the tensor core still accumulates tile it.  In r01's order (scores, wait, accumulate, next
        // scores) the row threads idled for two MMA round trips per tile.  The single S / dP TMEM buffer is free once p_full(it) has
        // fired (every row thread has read it); the P / dS smem tiles are guarded by pd_free.
        if (NI > 0) {
          mbar_wait(&st_full[0], 0);
          tc_fence_after();
          issue_scores(0);
        }
        for (int it = 0; it < NI; ++it) {
          int nstage = stage + 1, nphase = phase;
          if (nstage == Cf::STAGES) { nstage = 0; nphase ^= 1; }
          mbar_wait(p_full, it & 1);                          // row threads wrote P / dS of this tile
          tc_fence_after();
          if (it + 1 < NI) {
            mbar_wait(&st_full[nstage], nphase);
            tc_fence_after();
            issue_scores(nstage);
          }
          issue_acc(stage, it);
          stage = nstage; phase = nphase;
        }
      } else {
        for (int it = 0; it < NI; ++it) {
          mbar_wait(&st_full[stage], phase);
          tc_fence_after();
          issue_scores(stage);
          mbar_wait(p_full, it & 1);
          tc_fence_after();
          issue_acc(stage, it);
          if (++stage == Cf::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ===================== row threads: softmax backward + epilogue =====================
    // Eight warps: warps w and w + 4 share a TMEM lane quarter (rows) and split every 64-column inner tile in two
    // 32-column halves.  With four row warps one warp per scheduler carried the whole row phase and its TMEM-load / SFU
    // latencies sat on the critical path of the S -> P/dS -> accumulate chain (ncu: issue slots 29-33 % busy, XU 25-35 %,
    // profiles/r01_attn_bwd40_ncu_v4.md); four warps per scheduler (two CTAs per SM) hide them.
    const int q4 = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q4 * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int xrow = x0 + r;
    const bool row_ok = xrow < Lx;
    float lse_r = 0.f, D_r = 0.f;
    if (MODE == 0) {
      lse_r = p.lse_pad[(size_t)bh * p.Lq_pad + xrow];      // already multiplied by log2(e) by the pre-pass
      D_r = p.D_pad[(size_t)bh * p.Lq_pad + xrow];
    }
    const int klen = (p.kv_lens != nullptr) ? min(p.Lk, p.kv_lens[b]) : p.Lk;
    const int cb = half * 32;                    // this warp's first column inside a tile
    int stage = 0, phase = 0;
    for (int it = 0; it < NI; ++it) {
      mbar_wait(s_full, it & 1);
      if (MODE == 1) mbar_wait(&st_full[stage], phase);      // acquire the TMA-written lse / D slices for generic loads
      tc_fence_after();
      const int y0 = (it0 + it) * Cf::BY;
      const float* vec = reinterpret_cast<const float*>(sStage + stage * Cf::STAGE_BYTES + 2 * Cf::NAT_BYTES);
      // warp-uniform: no external dP, no padding / causal edge inside this (warp, tile) -> predicate-free fast path
      bool fast_w;
      if (MODE == 0) fast_w = (dp_ext == nullptr) && (y0 + Cf::BY <= klen) && (!p.causal || (y0 + Cf::BY - 1 <= x0 + q4 * 32));
      else           fast_w = (dp_ext == nullptr) && (x0 + q4 * 32 + 31 < klen) && (!p.causal || (x0 + q4 * 32 + 31 <= y0));
      const float sl2 = p.scale_log2;
      uint32_t vs[2][16], vd[2][16];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        tmem_ld_32x32b_x16(trow + (uint32_t)(Cf::COL_S + cb + 16 * g), vs[g]);
        tmem_ld_32x32b_x16(trow + (uint32_t)(Cf::COL_DP + cb + 16 * g), vd[g]);
      }
      tmem_ld_wait();
      unsigned char* half0 = sPD;                      // MODE 0: dS   | MODE 1: P^T     (one 64-wide K-major tile each)
      unsigned char* half1 = sPD + AB_ROWS * 128;      //              | MODE 1: dS^T
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int c0 = cb + 16 * g;
        uint32_t pk_p[8], pk_ds[8];
        if (fast_w) {
          if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float p0 = fast_exp2(fmaf(__uint_as_float(vs[g][i]), sl2, -lse_r));
              const float p1 = fast_exp2(fmaf(__uint_as_float(vs[g][i + 1]), sl2, -lse_r));
              pk_ds[i / 2] = ab_pack2<T>(p0 * (__uint_as_float(vd[g][i]) - D_r), p1 * (__uint_as_float(vd[g][i + 1]) - D_r));
            }
          } else {
            const float4* l4 = reinterpret_cast<const float4*>(vec + c0);
            const float4* d4 = reinterpret_cast<const float4*>(vec + Cf::BY + c0);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 L = l4[i / 4], Dv = d4[i / 4];
              const float p0 = fast_exp2(fmaf(__uint_as_float(vs[g][i]), sl2, -L.x));
              const float p1 = fast_exp2(fmaf(__uint_as_float(vs[g][i + 1]), sl2, -L.y));
              const float p2 = fast_exp2(fmaf(__uint_as_float(vs[g][i + 2]), sl2, -L.z));
              const float p3 = fast_exp2(fmaf(__uint_as_float(vs[g][i + 3]), sl2, -L.w));
              pk_p[i / 2] = ab_pack2<T>(p0, p1);
              pk_p[i / 2 + 1] = ab_pack2<T>(p2, p3);
              pk_ds[i / 2] = ab_pack2<T>(p0 * (__uint_as_float(vd[g][i]) - Dv.x), p1 * (__uint_as_float(vd[g][i + 1]) - Dv.y));
              pk_ds[i / 2 + 1] = ab_pack2<T>(p2 * (__uint_as_float(vd[g][i + 2]) - Dv.z), p3 * (__uint_as_float(vd[g][i + 3]) - Dv.w));
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float pv[2], dsv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = c0 + i + e;                      // inner index within the tile
              const int y = y0 + c;
              float pr, dp = __uint_as_float(vd[g][i + e]);
              if (MODE == 0) {
                const bool kv_ok = (y < klen) && (!p.causal || y <= xrow);
                pr = kv_ok ? fast_exp2(__uint_as_float(vs[g][i + e]) * sl2 - lse_r) : 0.f;
                if (dp_ext != nullptr && y < p.Lk && row_ok) dp += dp_ext[((size_t)bhx * p.Lq + xrow) * p.Lk + y];
                dsv[e] = pr * (dp - D_r);
              } else {
                const float lse_c = vec[c], D_c = vec[Cf::BY + c];
                const bool kv_ok = (xrow < klen) && (!p.causal || xrow <= y);
                pr = kv_ok ? fast_exp2(__uint_as_float(vs[g][i + e]) * sl2 - lse_c) : 0.f;   // lse pad = 1e30 -> 0 beyond Lq
                if (dp_ext != nullptr && y < p.Lq && row_ok) dp += dp_ext[((size_t)bhx * p.Lq + y) * p.Lk + xrow];
                dsv[e] = pr * (dp - D_c);
              }
              pv[e] = pr;
            }
            pk_p[i / 2] = ab_pack2<T>(pv[0], pv[1]);
            pk_ds[i / 2] = ab_pack2<T>(dsv[0], dsv[1]);
          }
        }
        if (g == 0 && it > 0) mbar_wait(pd_free, (it - 1) & 1);   // the previous tile's accumulating MMAs have read the P / dS tiles
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int cc = c0 / 8 + q;                     // 16-byte chunk (8 columns) inside the 64-wide row
          if (MODE == 0) {
            *reinterpret_cast<uint4*>(sw128_chunk(half0, r, cc)) = make_uint4(pk_ds[q * 4], pk_ds[q * 4 + 1], pk_ds[q * 4 + 2], pk_ds[q * 4 + 3]);
          } else {
            *reinterpret_cast<uint4*>(sw128_chunk(half0, r, cc)) = make_uint4(pk_p[q * 4], pk_p[q * 4 + 1], pk_p[q * 4 + 2], pk_p[q * 4 + 3]);
            *reinterpret_cast<uint4*>(sw128_chunk(half1, r, cc)) = make_uint4(pk_ds[q * 4], pk_ds[q * 4 + 1], pk_ds[q * 4 + 2], pk_ds[q * 4 + 3]);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      if (++stage == Cf::STAGES) { stage = 0; phase ^= 1; }
    }
    // ---- epilogue: accumulators -> 16-bit global.  MODE 0: the two warps of a lane quarter take alternate 16-column groups
    // of dQ (x scale).  MODE 1: half 0 writes acc0 = dV (x 1) -> out1, half 1 writes acc1 = dK (x scale) -> out0.
    mbar_wait(acc_full, 0);
    tc_fence_after();
    {
      const int o = (MODE == 0) ? 0 : half;
      const float mul = (MODE == 0 || o == 1) ? p.scale : 1.f;
      void* outp = (MODE == 0) ? p.out0 : (o == 0 ? p.out1 : p.out0);
      T* op = reinterpret_cast<T*>(outp) + ((size_t)b * Lx + xrow) * p.out_ld + (size_t)h * p.d;
      float* ap = nullptr;                                   // split-query mode: fp32 partial sums, converted by ab_cvt_kernel
      if (MODE == 1 && nsplit > 1) ap = (o == 0 ? p.acc32_1 : p.acc32_0) + ((size_t)b * Lx + xrow) * p.out_ld + (size_t)h * p.d;
#pragma unroll 1
      for (int c0 = (MODE == 0) ? half * 16 : 0; c0 < Cf::DN; c0 += (MODE == 0) ? 32 : 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(trow + (uint32_t)((o == 0 ? Cf::COL_ACC0 : Cf::COL_ACC1) + c0), v);
        tmem_ld_wait();
        if (row_ok) {
          if (ap != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              if (c0 + i + 4 <= D)
                atomicAdd(reinterpret_cast<float4*>(ap + c0 + i), make_float4(__uint_as_float(v[i]) * mul, __uint_as_float(v[i + 1]) * mul,
                                                                               __uint_as_float(v[i + 2]) * mul, __uint_as_float(v[i + 3]) * mul));
          } else {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              if (c0 + i + 8 <= D) {
                uint4 u;
                u.x = ab_pack2<T>(__uint_as_float(v[i]) * mul, __uint_as_float(v[i + 1]) * mul);
                u.y = ab_pack2<T>(__uint_as_float(v[i + 2]) * mul, __uint_as_float(v[i + 3]) * mul);
                u.z = ab_pack2<T>(__uint_as_float(v[i + 4]) * mul, __uint_as_float(v[i + 5]) * mul);
                u.w = ab_pack2<T>(__uint_as_float(v[i + 6]) * mul, __uint_as_float(v[i + 7]) * mul);
                *reinterpret_cast<uint4*>(op + c0 + i) = u;
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cf::TMEM_COLS); }
}

// D_i = sum_c dO o O (+ sum_k P dP_ext), lse (x log2 e) padded with 1e30 / D padded with 0 to Lq_pad.
// One warp per (sample, query) row: lanes stride each head's d channels (coalesced), one shuffle reduction per head.
template <typename T>
__global__ void __launch_bounds__(256) ab_prep_kernel(const T* __restrict__ o, const T* __restrict__ dO, const float* __restrict__ lse,
                                                      const float* __restrict__ probs, const float* __restrict__ dp_ext,
                                                      float* __restrict__ lse_pad, float* __restrict__ D_pad, int n, int Lq, int Lk, int H,
                                                      int d, int Lq_pad, int dp_b0) {
  pdl_grid_dependency_sync();
  __shared__ float s_acc[8][32];                       // per warp: up to 32 heads
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;   // row = b * Lq_pad + q
  if (row >= (long long)n * Lq_pad) return;
  const int q = (int)(row % Lq_pad), b = (int)(row / Lq_pad);
  if (q >= Lq) {
    for (int h = lane; h < H; h += 32) {
      lse_pad[((size_t)b * H + h) * Lq_pad + q] = 1e30f;
      D_pad[((size_t)b * H + h) * Lq_pad + q] = 0.f;
    }
    return;
  }
  const int Cc = H * d;
  const T* op = o + ((size_t)b * Lq + q) * Cc;
  const T* dp = dO + ((size_t)b * Lq + q) * Cc;
  if (((Cc | d) & 7) == 0 && ((reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(dO)) & 15) == 0) {
    // 16-byte loads: a vector of 8 channels lies inside one head (d % 8 == 0); per-head sums through shared-memory atomics
    // (r01 walked the heads one after the other with 2-byte loads: 16 us per call on the SDXL shapes)
    for (int h = lane; h < H; h += 32) s_acc[warp][h] = 0.f;
    __syncwarp();
    for (int v0 = lane; v0 < Cc / 8; v0 += 32) {
      const uint4 a = *reinterpret_cast<const uint4*>(op + v0 * 8), g = *reinterpret_cast<const uint4*>(dp + v0 * 8);
      const T* at = reinterpret_cast<const T*>(&a);
      const T* gt = reinterpret_cast<const T*>(&g);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) part = fmaf(to_f32<T>(at[i]), to_f32<T>(gt[i]), part);
      atomicAdd(&s_acc[warp][(v0 * 8) / d], part);
    }
  } else {
    for (int h = 0; h < H; ++h) {
      float part = 0.f;
      for (int c = lane; c < d; c += 32) part += to_f32<T>(op[h * d + c]) * to_f32<T>(dp[h * d + c]);
      part = warp_sum(part);
      if (lane == 0) s_acc[warp][h] = part;
    }
  }
  __syncwarp();
  for (int h = lane; h < H; h += 32) {
    const size_t bh = (size_t)b * H + h;
    float s = s_acc[warp][h];
    if (dp_ext != nullptr && b >= dp_b0) {
      const size_t bhx = (size_t)(b - dp_b0) * H + h;
      const float* pr = probs + (bhx * Lq + q) * Lk;
      const float* de = dp_ext + (bhx * Lq + q) * Lk;
      for (int k = 0; k < Lk; ++k) s += pr[k] * de[k];
    }
    lse_pad[bh * Lq_pad + q] = lse[bh * Lq + q] * 1.4426950408889634f;    // log2 units: p = exp2(s*scale*log2e - lse2)
    D_pad[bh * Lq_pad + q] = s;
  }
}

// split-query mode: fp32 partial sums of dK / dV -> 16-bit outputs
template <typename T>
__global__ void __launch_bounds__(256) ab_cvt_kernel(const float* __restrict__ a, const float* __restrict__ b2, T* __restrict__ oa,
                                                     T* __restrict__ ob, long long n) {
  pdl_grid_dependency_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  oa[i] = from_f32<T>(a[i]);
  ob[i] = from_f32<T>(b2[i]);
}

template <int D, int MODE, typename T>
static int launch_ab(const CUtensorMap* m, const AttnBwdKP& kp, dim3 grid, cudaStream_t st) {
  using Cf = ABCfg<D, MODE>;
  static bool configured = false;
  if (!configured) {
    COMAT_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<D, MODE, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::TOTAL));
    configured = true;
  }
  launch_k(attn_bwd_kernel<D, MODE, T>, grid, AB_THREADS, Cf::TOTAL, st, m[0], m[1], m[2], m[3], kp);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

template <int D, typename T>
static int run_bwd(const void* q, const void* k, const void* v, const void* dO, AttnBwdKP kp, int n,
                   void* dq, void* dk, void* dv, int fmt, cudaStream_t st, long long q_ld, long long k_ld, long long v_ld) {
  constexpr int DN = (D + 15) / 16 * 16;
  constexpr int BY = ABCfg<D, 0>::BY;
  const int H = kp.H, d = kp.d, Lq = kp.Lq, Lk = kp.Lk;
  const long long hd = (long long)H * d;
  auto nat_ld = [&](CUtensorMap* m, const void* base, int L, int rows, long long ld) {
    const uint64_t dims[3] = {(uint64_t)d, (uint64_t)H, (uint64_t)n * L};
    const uint64_t str[2] = {(uint64_t)d * 2, (uint64_t)ld * 2};
    const uint32_t box[3] = {64, 1, (uint32_t)rows};
    return make_tmap_16bit(m, base, 3, dims, str, box);
  };
  // Streamed operand tiles (K, V in the dQ pass; Q, dO in the dK / dV pass) use a 2-D map over (H*d, rows): for d = 40 the 64-wide
  // window also carries 24 columns of the next head, which meet the zero-filled columns of the RESIDENT tiles (3-D map, out of
  // bounds beyond d) in the S / dP products and land in accumulator columns >= d that the epilogue never stores.  With the 3-D
  // map every 80-byte box row ended out of bounds and cost 4.3 L2 requests (profiles/r02_attn_fwd_analysis.md).
  auto nat_st = [&](CUtensorMap* m, const void* base, int L, int rows, long long ld) {
    const uint64_t dims[2] = {(uint64_t)H * d, (uint64_t)n * L};
    const uint64_t str[1] = {(uint64_t)ld * 2};
    const uint32_t box[2] = {64, (uint32_t)rows};
    return make_tmap_16bit(m, base, 2, dims, str, box);
  };
  kp.idesc_s = make_idesc_f16(AB_ROWS, BY, fmt);
  kp.idesc_acc = make_idesc_f16(AB_ROWS, DN, fmt, 0, 1);       // B operand MN-major
  CUtensorMap m[4];
  memset(m, 0, sizeof(m));
  // ---- MODE 0: dQ
  bool ok = nat_ld(&m[0], q, Lq, AB_ROWS, q_ld) && nat_ld(&m[1], dO, Lq, AB_ROWS, hd) && nat_st(&m[2], k, Lk, BY, k_ld) && nat_st(&m[3], v, Lk, BY, v_ld);
  if (!ok) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
  kp.n_inner = (Lk + BY - 1) / BY;
  kp.out0 = dq; kp.out1 = nullptr;
  int rc = launch_ab<D, 0, T>(m, kp, dim3((Lq + AB_ROWS - 1) / AB_ROWS, H, n), st);
  if (rc) return rc;
  // ---- MODE 1: dK, dV
  ok = nat_ld(&m[0], k, Lk, AB_ROWS, k_ld) && nat_ld(&m[1], v, Lk, AB_ROWS, v_ld) && nat_st(&m[2], q, Lq, BY, q_ld) && nat_st(&m[3], dO, Lq, BY, hd);
  if (!ok) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
  kp.n_inner = (Lq + BY - 1) / BY;
  kp.out0 = dk; kp.out1 = dv;
  const int xtiles = (Lk + AB_ROWS - 1) / AB_ROWS;
  // few key tiles (cross-attention over 77 text tokens: one) and many query tiles: split the query range so the launch fills the GPU
  int nsplit = 1;
  if (kp.acc32_0 != nullptr && (long long)xtiles * H * n < num_sms() && kp.n_inner >= 8) {
    nsplit = (int)((2LL * num_sms() + (long long)xtiles * H * n - 1) / ((long long)xtiles * H * n));
    if (nsplit > kp.n_inner / 4) nsplit = kp.n_inner / 4;
    if (nsplit < 1) nsplit = 1;
  }
  kp.nsplit = nsplit;
  if (nsplit == 1) return launch_ab<D, 1, T>(m, kp, dim3(xtiles, H, n), st);
  const long long cnt = (long long)n * Lk * hd;
  if (cudaMemsetAsync(kp.acc32_0, 0, (size_t)cnt * 2 * sizeof(float), st) != cudaSuccess) { comat_set_cuda_error(-1); return COMAT_ERR_CUDA; }
  rc = launch_ab<D, 1, T>(m, kp, dim3(xtiles * nsplit, H, n), st);
  if (rc) return rc;
  launch_k(ab_cvt_kernel<T>, (unsigned)((cnt + 255) / 256), 256, 0, st, (const float*)kp.acc32_0, (const float*)kp.acc32_1, (T*)dk, (T*)dv, cnt);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

}  // namespace comat
using namespace comat;

extern "C" size_t comat_attention_bwd_workspace_bytes(int n, int Lq, int Lk, int H, int d) {
  const size_t Lqp = (size_t)(Lq + 127) / 128 * 128, Lkp = (size_t)(Lk + 127) / 128 * 128;
  (void)Lkp; (void)d;
  // padded lse (x log2 e) and D vectors + fp32 dK / dV partial sums of the split-query mode (short key sequences only)
  const size_t split = (Lk <= 1024) ? (size_t)n * Lk * H * d * 8 + 256 : 0;
  return (size_t)n * H * Lqp * 8 + 1024 + split;
}

extern "C" int comat_attention_bwd_strided(const void* q, const void* k, const void* v, const void* o, const void* dO, const float* lse,
                                           const float* probs, const float* dp_ext, void* dq, void* dk, void* dv, void* workspace, int n,
                                           int Lq, int Lk, int H, int d, long long q_ld, long long k_ld, long long v_ld, float scale,
                                           int dtype, const int* kv_lens, int causal, int dp_first_sample, void* stream) {
  if (!q || !k || !v || !o || !dO || !lse || !dq || !dk || !dv || !workspace) return COMAT_ERR_INVALID;
  if (dp_first_sample < 0 || dp_first_sample >= n) return COMAT_ERR_INVALID;
  if (q_ld < (long long)H * d || k_ld < (long long)H * d || v_ld < (long long)H * d || (q_ld % 8) || (k_ld % 8) || (v_ld % 8)) return COMAT_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) || (reinterpret_cast<uintptr_t>(v) & 15)) return COMAT_ERR_INVALID;
  if (d != 40 && d != 64 && d != 80 && d != 128 && d != 160 && d != 32 && d != 16) return COMAT_ERR_UNSUPPORTED;
  if (dtype != COMAT_F16 && dtype != COMAT_BF16) return COMAT_ERR_UNSUPPORTED;
  if (dp_ext && !probs) return COMAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int Lqp = (Lq + 127) / 128 * 128, Lkp = (Lk + 127) / 128 * 128;
  unsigned char* ws = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  float* lse_pad = reinterpret_cast<float*>(ws);
  float* D_pad = lse_pad + (size_t)n * H * Lqp;
  {
    const long long rows = (long long)n * Lqp;
    if (H > 32) return COMAT_ERR_UNSUPPORTED;
    if (dtype == COMAT_F16)
      launch_k(ab_prep_kernel<__half>, (unsigned)((rows + 7) / 8), 256, 0, st, (const __half*)o, (const __half*)dO, lse, probs, dp_ext, lse_pad, D_pad, n, Lq, Lk, H, d, Lqp, dp_first_sample);
    else
      launch_k(ab_prep_kernel<__nv_bfloat16>, (unsigned)((rows + 7) / 8), 256, 0, st, (const __nv_bfloat16*)o, (const __nv_bfloat16*)dO, lse, probs, dp_ext, lse_pad, D_pad, n, Lq, Lk, H, d, Lqp, dp_first_sample);
  }
  AttnBwdKP kp;
  memset(&kp, 0, sizeof(kp));
  kp.Lq = Lq; kp.Lk = Lk; kp.H = H; kp.d = d; kp.Lq_pad = Lqp; kp.Lk_pad = Lkp;
  kp.scale = scale; kp.scale_log2 = scale * 1.4426950408889634f;
  kp.kv_lens = kv_lens; kp.causal = causal;
  kp.lse_pad = lse_pad; kp.D_pad = D_pad; kp.dp_ext = dp_ext; kp.dp_b0 = dp_first_sample; kp.out_ld = (long long)H * d;
  if (Lk <= 1024) {                               // split-query scratch lives behind the two padded vectors (workspace_bytes sizes it)
    float* sc = D_pad + (size_t)n * H * Lqp;
    sc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sc) + 255) & ~uintptr_t(255));
    kp.acc32_0 = sc; kp.acc32_1 = sc + (size_t)n * Lk * H * d;
  }
  const int fmt = dtype == COMAT_BF16 ? 1 : 0;
#define AB_CASE(DD)                                                                                                          \
  case DD:                                                                                                                   \
    return fmt ? run_bwd<DD, __nv_bfloat16>(q, k, v, dO, kp, n, dq, dk, dv, fmt, st, q_ld, k_ld, v_ld)                       \
               : run_bwd<DD, __half>(q, k, v, dO, kp, n, dq, dk, dv, fmt, st, q_ld, k_ld, v_ld);
  switch (d) { AB_CASE(16) AB_CASE(32) AB_CASE(40) AB_CASE(64) AB_CASE(80) AB_CASE(128) AB_CASE(160) }
#undef AB_CASE
  return COMAT_ERR_UNSUPPORTED;
}

extern "C" int comat_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dO, const float* lse,
                                   const float* probs, const float* dp_ext, void* dq, void* dk, void* dv, void* workspace, int n,
                                   int Lq, int Lk, int H, int d, float scale, int dtype, const int* kv_lens, int causal, void* stream) {
  const long long ld = (long long)H * d;
  return comat_attention_bwd_strided(q, k, v, o, dO, lse, probs, dp_ext, dq, dk, dv, workspace, n, Lq, Lk, H, d, ld, ld, ld, scale, dtype,
                                     kv_lens, causal, 0, stream);
}
