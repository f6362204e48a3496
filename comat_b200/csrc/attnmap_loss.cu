// Attention-map token / pixel loss — one coalesced HBM-bound pass over every stored cross-attention map.
//
// Replaces the nested Python loops of attr_concen_utils/gsam_interface.py:204-223 calling
// attn_utils/tc_loss_utils.py:66-173 (thousands of tiny aten launches) with
//   fwd : attnmap_fwd_kernel (streams every map exactly once: algorithmic bytes = 4 B x stored elements)
//         + attnmap_finalize_kernel (tiny)
//   bwd : attnmap_bwd_kernel (writes every dP exactly once; never re-reads P).
//
// Data movement: each CTA owns 32 consecutive pixels of one (group, sample) and loops over the group's
// (map, head) slabs; a slab tile is 32 x T contiguous floats (9 856 B at T=77), fetched with one 1-D bulk async
// copy (cp.async.bulk -> UBLKCP, TMA engine) per stage into a 4-deep shared-memory ring guarded by mbarriers.
// Reductions over the 32 pixels are warp shuffles; cross-tile sums are written as partials and reduced in a
// fixed order by the finalize kernel (bit-reproducible, no float atomics in global memory).
#include "common.cuh"

namespace comat {

constexpr int TILE_PX = COMAT_ATTNMAP_TILE_PX;   // 32 pixels = one warp lane per pixel
constexpr int MAXP = COMAT_ATTNMAP_MAX_PAIRS;
constexpr int MAXW = COMAT_ATTNMAP_MAX_WORDS;
constexpr int STAGES = 4;
constexpr int FWD_THREADS = 128;
constexpr int NWARPS = FWD_THREADS / 32;
constexpr int PAIRS_PER_WARP = MAXP / NWARPS;

struct StateLayout {
  size_t part_nd, part_bce, nd, coef, pred, gb_loss, total;
  int nd_stride;    // floats per (work item) in part_nd and per (group,sample) in nd: maxMG*maxH*MAXP*2
  int coef_stride;  // maxMG*MAXP
};

__host__ __device__ inline StateLayout make_layout(int n_work, int n_groups, int B, int maxH, int maxMG, long long pred_floats) {
  StateLayout L;
  L.nd_stride = maxMG * maxH * MAXP * 2;
  L.coef_stride = maxMG * MAXP;
  size_t o = 0;
  L.part_nd = o;  o += (size_t)n_work * L.nd_stride;
  L.part_bce = o; o += (size_t)n_work * MAXW;
  L.nd = o;       o += (size_t)n_groups * B * L.nd_stride;
  L.coef = o;     o += (size_t)n_groups * B * L.coef_stride;
  L.pred = o;     o += (size_t)pred_floats;
  L.gb_loss = o;  o += (size_t)n_groups * B * 2;
  L.total = o;
  return L;
}

struct Tables {
  const int64_t* map_ptr;
  const int32_t* grp;
  const int32_t* smp;
  const int32_t* pair;
  const int32_t* word_ntok;
  const int32_t* work;
  const float* masks;
  int n_groups, B, T, maxH, maxMG, n_work;
};

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(FWD_THREADS) attnmap_fwd_kernel(Tables tb, float* __restrict__ state, StateLayout L) {
  pdl_grid_dependency_sync();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int T = tb.T;
  const int tile_floats = TILE_PX * T;
  float* ring = reinterpret_cast<float*>(smem_raw);                               // STAGES * tile_floats
  float* s_mask = ring + (size_t)STAGES * tile_floats;                            // MAXW * 32
  int* s_tok = reinterpret_cast<int*>(s_mask + MAXW * TILE_PX);                   // MAXP
  int* s_wl = s_tok + MAXP;                                                       // MAXP
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_wl + MAXP);                      // STAGES (8B aligned: offsets are x4 floats)

  const int w = blockIdx.x;
  const int g = tb.work[w * 4 + 0], b = tb.work[w * 4 + 1], tile = tb.work[w * 4 + 2];
  const int* G = tb.grp + g * 8;
  const int res = G[0], map_begin = G[1], map_end = G[2], H = G[3], mask_off = G[4], pred_base = G[6];
  const int* S = tb.smp + b * 4;
  const int pair_begin = S[0], np = S[1] - S[0], word_begin = S[2], nw = S[3] - S[2];
  if (np <= 0) return;   // sample without words / masks contributes nothing (gsam_interface.py:188-202)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HW = res * res;
  const int px0 = tile * TILE_PX;
  const int n_it = (map_end - map_begin) * H;
  const uint32_t tile_bytes = (uint32_t)tile_floats * 4u;

  if (tid < np) {
    s_tok[tid] = tb.pair[(pair_begin + tid) * 2 + 1];
    s_wl[tid] = tb.pair[(pair_begin + tid) * 2 + 0] - word_begin;
  }
  for (int i = tid; i < nw * TILE_PX; i += FWD_THREADS) {
    int wl = i / TILE_PX, p = i % TILE_PX;
    s_mask[i] = tb.masks[(size_t)mask_off + (size_t)(word_begin + wl) * HW + px0 + p];
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
    fence_proxy_async_smem();
  }
  __syncthreads();

  auto issue = [&](int it) {
    int m = map_begin + it / H, h = it % H;
    const float* src = reinterpret_cast<const float*>(tb.map_ptr[m]) + ((size_t)(b * H + h) * HW + px0) * T;
    int s = it % STAGES;
    mbar_expect_tx(&bars[s], tile_bytes);
    bulk_g2s(ring + (size_t)s * tile_floats, src, tile_bytes, &bars[s]);
  };
  if (tid == 0) {
    for (int it = 0; it < STAGES && it < n_it; ++it) issue(it);
  }

  float acc[PAIRS_PER_WARP];
#pragma unroll
  for (int k = 0; k < PAIRS_PER_WARP; ++k) acc[k] = 0.f;

  // partial region of this (group, sample): [slot][tile][2] (tile fastest -> the reducer reads it coalesced)
  const int n_tiles = G[5], work_begin = G[7];
  const size_t region = (size_t)(work_begin + b * n_tiles);
  float* part = state + L.part_nd + region * L.nd_stride;
  for (int it = 0; it < n_it; ++it) {
    const int s = it % STAGES;
    mbar_wait(&bars[s], (uint32_t)((it / STAGES) & 1));
    const float* buf = ring + (size_t)s * tile_floats + lane * T;
    const int lm = it / H, h = it % H;
    const int slot0 = (lm * tb.maxH + h) * MAXP;
#pragma unroll
    for (int k = 0; k < PAIRS_PER_WARP; ++k) {
      const int j = warp + k * NWARPS;
      if (j < np) {
        const float v = buf[s_tok[j]];
        const float mk = s_mask[s_wl[j] * TILE_PX + lane];
        acc[k] += v;
        const float num = warp_sum(v * mk);
        const float den = warp_sum(v);
        if (lane == 0)
          *reinterpret_cast<float2*>(part + ((size_t)(slot0 + j) * n_tiles + tile) * 2) = make_float2(num, den);
      }
    }
    __syncthreads();   // everyone is done with stage s
    if (tid == 0 && it + STAGES < n_it) issue(it + STAGES);
  }

  // ---- pixel loss for this tile: pred_i = sum_{p in tau_i} mean_{maps,heads} P[...,p]  (tc_loss_utils.py:131-156)
  float* s_acc = ring;   // reuse: MAXP * 32 floats (all bulk copies have completed and been consumed)
#pragma unroll
  for (int k = 0; k < PAIRS_PER_WARP; ++k) {
    const int j = warp + k * NWARPS;
    if (j < np) s_acc[j * TILE_PX + lane] = acc[k];
  }
  __syncthreads();
  const float inv_cnt = 1.0f / (float)n_it;
  float* pred_out = state + L.pred + (size_t)pred_base + (size_t)b * MAXW * HW;
  float* pb = state + L.part_bce + region * MAXW;
  for (int wl = warp; wl < nw; wl += NWARPS) {
    float p = 0.f;
    for (int j = 0; j < np; ++j)
      if (s_wl[j] == wl) p += s_acc[j * TILE_PX + lane];
    p *= inv_cnt;
    pred_out[(size_t)wl * HW + px0 + lane] = p;
    const float mk = s_mask[wl * TILE_PX + lane];
    const float l1 = fmaxf(logf(p), -100.f);
    const float l0 = fmaxf(logf(fmaxf(1.f - p, 0.f)), -100.f);
    const float bce = -(mk * l1 + (1.f - mk) * l0);
    const float sum = warp_sum(bce);
    if (lane == 0) pb[(size_t)wl * n_tiles + tile] = sum;
  }
}

// ------------------------------------------------------------------------------------------------ finalize
// Stage A: one warp per (group, sample, map, head, pair) slot sums its tile partials (coalesced float2 reads, fixed-order
// shuffle tree -> bit-reproducible).  Stage B: one CTA per (group, sample): token-loss terms, backward coefficients, pixel sum;
// the last CTA adds the per-(group, sample) results in index order.
__global__ void __launch_bounds__(256) attnmap_reduce_kernel(Tables tb, float* __restrict__ state, StateLayout L) {
  pdl_grid_dependency_sync();
  const int slots = tb.maxMG * tb.maxH * MAXP;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)tb.n_groups * tb.B * slots) return;
  const int gb = (int)(wid / slots), slot = (int)(wid % slots);
  const int g = gb / tb.B, b = gb % tb.B;
  const int* G = tb.grp + g * 8;
  const int nm = G[2] - G[1], H = G[3], n_tiles = G[5], work_begin = G[7];
  const int np = tb.smp[b * 4 + 1] - tb.smp[b * 4 + 0];
  const int j = slot % MAXP, h = (slot / MAXP) % tb.maxH, lm = slot / (MAXP * tb.maxH);
  if (j >= np || h >= H || lm >= nm) return;
  const float2* p = reinterpret_cast<const float2*>(state + L.part_nd + (size_t)(work_begin + b * n_tiles) * L.nd_stride) +
                    (size_t)slot * n_tiles;
  float num = 0.f, den = 0.f;
  for (int t = lane; t < n_tiles; t += 32) { const float2 v = p[t]; num += v.x; den += v.y; }
  num = warp_sum(num); den = warp_sum(den);
  if (lane == 0) *reinterpret_cast<float2*>(state + L.nd + (size_t)gb * L.nd_stride + (size_t)slot * 2) = make_float2(num, den);
}

__global__ void __launch_bounds__(256) attnmap_finalize_kernel(Tables tb, float* __restrict__ state, StateLayout L,
                                                               float* __restrict__ loss2, unsigned int* counter) {
  pdl_grid_dependency_sync();
  __shared__ float s_red[256];
  __shared__ bool s_last;
  const int gb = blockIdx.x;
  const int g = gb / tb.B, b = gb % tb.B;
  const int* G = tb.grp + g * 8;
  const int res = G[0], nm = G[2] - G[1], H = G[3], n_tiles = G[5], work_begin = G[7];
  const int* S = tb.smp + b * 4;
  const int pair_begin = S[0], np = S[1] - S[0], nw = S[3] - S[2];
  const int tid = threadIdx.x;
  float tok_local = 0.f, pix_local = 0.f;
  if (np > 0) {
    const float* nd = state + L.nd + (size_t)gb * L.nd_stride;
    float* coef = state + L.coef + (size_t)gb * L.coef_stride;
    const float invW = 1.f / (float)nw, invB = 1.f / (float)tb.B;
    for (int i = tid; i < nm * np; i += blockDim.x) {          // token loss terms (tc_loss_utils.py:104-125)
      const int j = i % np, lm = i / np;
      float fm = 0.f;
      for (int h = 0; h < H; ++h) {
        const size_t off = ((size_t)(lm * tb.maxH + h) * MAXP + j) * 2;
        fm += nd[off] / nd[off + 1];
      }
      fm /= (float)H;
      const int wg = tb.pair[(pair_begin + j) * 2 + 0];
      const float inv_nt = 1.f / (float)tb.word_ntok[wg];
      const float d = 1.f - fm;
      tok_local += d * d * inv_nt * invW;
      coef[lm * MAXP + j] = -2.f * d * inv_nt * invW * invB / (float)H;
    }
    const float* pb = state + L.part_bce + (size_t)(work_begin + b * n_tiles) * MAXW;     // [wl][tile]
    for (int i = tid; i < nw * n_tiles; i += blockDim.x) pix_local += pb[i];              // (tc_loss_utils.py:153-167)
    pix_local *= invW / (float)(res * res);
  }
  s_red[tid] = tok_local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_red[tid] += s_red[tid + o];
    __syncthreads();
  }
  const float tok = s_red[0];
  __syncthreads();
  s_red[tid] = pix_local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_red[tid] += s_red[tid + o];
    __syncthreads();
  }
  const float pix = s_red[0];
  if (tid == 0) {
    state[L.gb_loss + (size_t)gb * 2 + 0] = tok;
    state[L.gb_loss + (size_t)gb * 2 + 1] = pix;
    __threadfence();
    const unsigned int done = atomicAdd(counter, 1u);
    s_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && tid == 0) {
    __threadfence();
    float a = 0.f, c = 0.f;
    const volatile float* gl = state + L.gb_loss;
    for (int i = 0; i < (int)gridDim.x; ++i) {   // fixed order -> reproducible
      a += gl[i * 2 + 0];
      c += gl[i * 2 + 1];
    }
    loss2[0] = a / (float)tb.B;   // gsam_interface.py:225-226
    loss2[1] = c / (float)tb.B;
    *counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int BWD_THREADS = 128;
__global__ void __launch_bounds__(BWD_THREADS) attnmap_bwd_kernel(Tables tb, const float* __restrict__ state, StateLayout L,
                                                                  const float* __restrict__ grad2,
                                                                  const int64_t* __restrict__ map_grad_ptr) {
  pdl_grid_dependency_sync();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int T = tb.T;
  const int tile_floats = TILE_PX * T;
  float* obuf = reinterpret_cast<float*>(smem_raw);            // 2 * tile_floats
  float* s_mask = obuf + 2 * (size_t)tile_floats;              // MAXW*32
  float* s_dbce = s_mask + MAXW * TILE_PX;                     // MAXW*32
  int* s_tok = reinterpret_cast<int*>(s_dbce + MAXW * TILE_PX);
  int* s_wl = s_tok + MAXP;

  const int w = blockIdx.x;
  const int g = tb.work[w * 4 + 0], b = tb.work[w * 4 + 1], tile = tb.work[w * 4 + 2];
  const int* G = tb.grp + g * 8;
  const int res = G[0], map_begin = G[1], map_end = G[2], H = G[3], mask_off = G[4], pred_base = G[6];
  const int* S = tb.smp + b * 4;
  const int pair_begin = S[0], np = S[1] - S[0], word_begin = S[2], nw = S[3] - S[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HW = res * res, px0 = tile * TILE_PX;
  const int n_it = (map_end - map_begin) * H;
  const uint32_t tile_bytes = (uint32_t)tile_floats * 4u;
  const int gb = g * tb.B + b;

  for (int i = tid; i < 2 * tile_floats; i += BWD_THREADS) obuf[i] = 0.f;
  if (np > 0) {
    const float g_pix = grad2[1];
    if (tid < np) {
      s_tok[tid] = tb.pair[(pair_begin + tid) * 2 + 1];
      s_wl[tid] = tb.pair[(pair_begin + tid) * 2 + 0] - word_begin;
    }
    const float* pred = state + L.pred + (size_t)pred_base + (size_t)b * MAXW * HW;
    const float cpix = g_pix / ((float)HW * (float)n_it * (float)nw * (float)tb.B);
    for (int i = tid; i < nw * TILE_PX; i += BWD_THREADS) {
      const int wl = i / TILE_PX, p = i % TILE_PX;
      const float mk = tb.masks[(size_t)mask_off + (size_t)(word_begin + wl) * HW + px0 + p];
      const float x = pred[(size_t)wl * HW + px0 + p];
      s_mask[i] = mk;
      // aten binary_cross_entropy_backward: (x - t) / max((1 - x) * x, 1e-12)
      s_dbce[i] = cpix * (x - mk) / fmaxf((1.f - x) * x, 1e-12f);
    }
  }
  __syncthreads();
  const float g_tok = grad2[0];
  const float* nd = state + L.nd + (size_t)gb * L.nd_stride;
  const float* coef = state + L.coef + (size_t)gb * L.coef_stride;

  for (int it = 0; it < n_it; ++it) {
    const int lm = it / H, h = it % H;
    float* buf = obuf + (size_t)(it & 1) * tile_floats;
    if (it >= 2) {
      if (tid == 0) bulk_wait_read<1>();   // the store that last used this buffer has finished reading smem
      __syncthreads();
    }
    if (np > 0) {
      // clear the touched columns, then accumulate (pairs of different words may share a token column)
      for (int j = warp; j < np; j += BWD_THREADS / 32) buf[lane * T + s_tok[j]] = 0.f;
      __syncthreads();
      for (int j = warp; j < np; j += BWD_THREADS / 32) {
        const size_t off = ((size_t)(lm * tb.maxH + h) * MAXP + j) * 2;
        const float num = nd[off], den = nd[off + 1];
        const float c = g_tok * coef[lm * MAXP + j];
        const int wl = s_wl[j];
        const float mk = s_mask[wl * TILE_PX + lane];
        const float val = c * (mk * den - num) / (den * den) + s_dbce[wl * TILE_PX + lane];
        atomicAdd(&buf[lane * T + s_tok[j]], val);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      float* dst = reinterpret_cast<float*>(map_grad_ptr[map_begin + lm]) + ((size_t)(b * H + h) * HW + px0) * T;
      bulk_s2g(dst, buf, tile_bytes);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all<0>();
}

// ------------------------------------------------------------------------------------------------ mask resize
// torchvision Resize(antialias=True) on a bool mask followed by `> 0` (tc_loss_utils.py:88-94): the output pixel is
// 1 iff any input pixel with a non-zero separable triangle-filter weight is set.  Weights follow aten's
// _compute_indices_weights_aa (bilinear, align_corners=False): support = max(scale,1), taps in
// [int(center - support + .5), int(center + support + .5)), w = 1 - |(j + 0.5 - center) / max(scale,1)| if < 1.
__device__ __forceinline__ void aa_window(int o, float scale, int in_size, int& lo, int& hi, float& center, float& inv) {
  const float support = scale >= 1.f ? scale : 1.f;
  inv = scale >= 1.f ? 1.f / scale : 1.f;
  center = scale * ((float)o + 0.5f);
  lo = max((int)(center - support + 0.5f), 0);
  hi = min((int)(center + support + 0.5f), in_size);
}
__global__ void mask_resize_any_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int n, int in_h, int in_w,
                                       int res) {
  pdl_grid_dependency_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * res * res) return;
  const int x = idx % res, y = (idx / res) % res, i = idx / (res * res);
  const float sy = (float)in_h / (float)res, sx = (float)in_w / (float)res;
  int y0, y1, x0, x1;
  float cy, cx, iy, ix;
  aa_window(y, sy, in_h, y0, y1, cy, iy);
  aa_window(x, sx, in_w, x0, x1, cx, ix);
  const uint8_t* src = in + (size_t)i * in_h * in_w;
  bool any = false;
  for (int yy = y0; yy < y1 && !any; ++yy) {
    const float wy = 1.f - fabsf(((float)yy - cy + 0.5f) * iy);
    if (!(wy > 0.f)) continue;
    for (int xx = x0; xx < x1; ++xx) {
      const float wx = 1.f - fabsf(((float)xx - cx + 0.5f) * ix);
      if (wx > 0.f && src[(size_t)yy * in_w + xx]) {
        any = true;
        break;
      }
    }
  }
  out[idx] = any ? 1.f : 0.f;
}

static inline Tables make_tables(const comat_attnmap_plan* p) {
  Tables t;
  t.map_ptr = p->map_ptr; t.grp = p->grp; t.smp = p->smp; t.pair = p->pair; t.word_ntok = p->word_ntok;
  t.work = p->work; t.masks = p->masks;
  t.n_groups = p->n_groups; t.B = p->n_samples; t.T = p->tokens; t.maxH = p->max_heads;
  t.maxMG = p->max_maps_per_group; t.n_work = p->n_work;
  return t;
}
static inline size_t fwd_smem(int T) {
  return (size_t)STAGES * TILE_PX * T * 4 + MAXW * TILE_PX * 4 + 2 * MAXP * 4 + STAGES * 8 + 16;
}

static inline size_t bwd_smem(int T) { return (size_t)2 * TILE_PX * T * 4 + 2 * MAXW * TILE_PX * 4 + 2 * MAXP * 4; }

}  // namespace comat

using namespace comat;

extern "C" size_t comat_attnmap_loss_state_floats(const comat_attnmap_plan* p) {
  if (!p) return 0;
  return make_layout(p->n_work, p->n_groups, p->n_samples, p->max_heads, p->max_maps_per_group, p->pred_floats).total;
}

static int check_plan(const comat_attnmap_plan* p) {
  if (!p || !p->map_ptr || !p->grp || !p->smp || !p->work || !p->masks) return COMAT_ERR_INVALID;
  if (p->n_work <= 0 || p->n_groups <= 0 || p->n_samples <= 0 || p->tokens <= 0) return COMAT_ERR_INVALID;
  if ((size_t)(STAGES * TILE_PX * p->tokens * 4) > 200 * 1024) return COMAT_ERR_UNSUPPORTED;
  return COMAT_OK;
}

extern "C" int comat_attnmap_loss_fwd(const comat_attnmap_plan* p, float* loss2, float* state, size_t state_floats,
                                      unsigned int* counter, void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!loss2 || !state || !counter) return COMAT_ERR_INVALID;
  StateLayout L = make_layout(p->n_work, p->n_groups, p->n_samples, p->max_heads, p->max_maps_per_group, p->pred_floats);
  if (state_floats < L.total) return COMAT_ERR_WORKSPACE;
  Tables tb = make_tables(p);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = fwd_smem(p->tokens);
  static size_t configured = 0;
  if (smem > configured) {
    COMAT_CUDA(cudaFuncSetAttribute(attnmap_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  launch_k(attnmap_fwd_kernel, p->n_work, FWD_THREADS, smem, st, tb, state, L);
  COMAT_CHECK_LAUNCH();
  {
    const long long warps = (long long)p->n_groups * p->n_samples * p->max_maps_per_group * p->max_heads * MAXP;
    launch_k(attnmap_reduce_kernel, (unsigned)((warps * 32 + 255) / 256), 256, 0, st, tb, state, L);
  }
  launch_k(attnmap_finalize_kernel, p->n_groups * p->n_samples, 256, 0, st, tb, state, L, loss2, counter);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

extern "C" int comat_attnmap_loss_bwd(const comat_attnmap_plan* p, const float* grad2, const float* state,
                                      const int64_t* map_grad_ptr, void* stream) {
  int rc = check_plan(p);
  if (rc) return rc;
  if (!grad2 || !state || !map_grad_ptr) return COMAT_ERR_INVALID;
  StateLayout L = make_layout(p->n_work, p->n_groups, p->n_samples, p->max_heads, p->max_maps_per_group, p->pred_floats);
  Tables tb = make_tables(p);
  const size_t smem = bwd_smem(p->tokens);
  static size_t configured = 0;
  if (smem > configured) {
    COMAT_CUDA(cudaFuncSetAttribute(attnmap_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  launch_k(attnmap_bwd_kernel, p->n_work, BWD_THREADS, smem, (cudaStream_t)stream, tb, state, L, grad2, map_grad_ptr);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}

extern "C" int comat_mask_resize_any(const uint8_t* in, float* out, int n, int in_h, int in_w, int res, void* stream) {
  if (!in || !out || n <= 0 || in_h <= 0 || in_w <= 0 || res <= 0) return COMAT_ERR_INVALID;
  const int total = n * res * res;
  launch_k(mask_resize_any_kernel, (total + 255) / 256, 256, 0, (cudaStream_t)stream, in, out, n, in_h, in_w, res);
  COMAT_CHECK_LAUNCH();
  return COMAT_OK;
}
