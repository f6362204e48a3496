/* comat_b200 — C ABI of the B200-native CoMat training hot path.
 *
 * The reference (CaraJ7/CoMat) is pure Python and has no FFI: its seam is the Python object surface listed in
 * SURVEY.md section 8b.  This header is the boundary *underneath* that surface: every entry point replaces the
 * arithmetic of one reference function (cited per declaration, paths relative to the reference tree) and is what
 * a ctypes / cffi binding on the reference side would bind.
 *
 * Conventions
 *   - plain C: raw device pointers + sizes, no torch / C++ types;
 *   - the caller allocates every output and workspace; nothing is freed or retained by the callee;
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as void*); no host synchronisation,
 *     no global mutable state, re-entrant per stream;
 *   - return 0 on success, a negative comat_status otherwise (comat_strerror() names it); CUDA launch errors are
 *     returned as COMAT_ERR_CUDA and the cudaError_t is retrievable with comat_last_cuda_error();
 *   - there is NO CPU fallback: built for sm_100a only.
 */
#ifndef COMAT_B200_H
#define COMAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  COMAT_OK = 0,
  COMAT_ERR_INVALID = -1,   /* bad argument (shape, alignment, null pointer)            */
  COMAT_ERR_UNSUPPORTED = -2, /* shape/dtype combination not built                        */
  COMAT_ERR_CUDA = -3,      /* a CUDA runtime/driver call failed                        */
  COMAT_ERR_WORKSPACE = -4  /* workspace too small                                      */
} comat_status;

typedef enum { COMAT_F32 = 0, COMAT_F16 = 1, COMAT_BF16 = 2 } comat_dtype;

int comat_version(void);
const char* comat_strerror(int status);
int comat_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * Attention-map token / pixel loss.
 * Replaces: attn_utils/tc_loss_utils.py:66-173 (get_grounding_loss_by_layer) batched over the loops of
 *           attr_concen_utils/gsam_interface.py:140-228 (get_mask_loss): all samples x timesteps x layers of one
 *           training step in ONE launch (+ a small finalize launch).
 *
 * Device tables (int32 unless noted), built once per step by the host mirror (comat_b200/attn_loss.py):
 *   map_ptr   int64[n_maps]      device address of each stored probability map, fp32, layout (B*H, res*res, T)
 *   grp       int32[n_groups*8]  {res, map_begin, map_end, H, mask_off, n_tiles, pred_off_base, work_begin}
 *   smp       int32[B*4]         {pair_begin, pair_end, word_begin, word_end}
 *   pair      int32[n_pairs*2]   {global word id, token column}
 *   word_ntok int32[n_words]     tokens per word (|tau_i|)
 *   work      int32[n_work*4]    {group, sample, tile, _}; ordered (group, sample, tile), every sample present,
 *                                one CTA each (32 pixels per tile)
 *   masks     f32[...]           binarised masks, per group at grp.mask_off: (n_words, res*res)
 * Outputs: loss[2] = {token_loss, pixel_loss} (already divided by B, gsam_interface.py:225-226);
 *          saved state for backward in `state` (size from comat_attnmap_loss_state_floats).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t n_maps, n_groups, n_samples, n_words, n_pairs, n_work, tokens;
  int32_t max_heads, max_maps_per_group, _pad;
  int64_t pred_floats;            /* sum over groups of B*MAX_WORDS*res*res */
  const int64_t* map_ptr;
  const int32_t* grp;
  const int32_t* smp;
  const int32_t* pair;
  const int32_t* word_ntok;
  const int32_t* work;
  const float* masks;
} comat_attnmap_plan;

#define COMAT_ATTNMAP_TILE_PX 32
#define COMAT_ATTNMAP_MAX_PAIRS 32
#define COMAT_ATTNMAP_MAX_WORDS 16

size_t comat_attnmap_loss_state_floats(const comat_attnmap_plan* plan_counts);
int comat_attnmap_loss_fwd(const comat_attnmap_plan* plan, float* loss2, float* state, size_t state_floats,
                           unsigned int* counter, void* stream);
/* dP for every map (same layout as the map, fp32, dense).  map_grad_ptr: int64[n_maps] device addresses.
 * grad2 = d(total)/d{token_loss,pixel_loss} (device pointer, 2 floats). */
int comat_attnmap_loss_bwd(const comat_attnmap_plan* plan, const float* grad2, const float* state,
                           const int64_t* map_grad_ptr, void* stream);

/* tc_loss_utils.py:88-94: Resize((res,res), antialias=True) of a bool mask then `> 0` == windowed "any".
 * in: u8 (n, in_h, in_w)  out: f32 (n, res, res) of 0/1. */
int comat_mask_resize_any(const uint8_t* in, float* out, int n, int in_h, int in_w, int res, void* stream);


/* ------------------------------------------------------------------------------------------------------------
 * Tensor-core GEMM / implicit-GEMM convolution (tcgen05 + TMEM + TMA), 16-bit in, fp32 accumulate:
 *     out[m,n] = act(alpha * sum_k A[m,k] * B[n,k] + bias[n] + rowvec[m / rows_per_group, n]) + residual[m,n]
 * Replaces the library calls behind diffusers Linear / Conv2d / LoRACompatibleLinear on the UNet, VAE and BLIP
 * paths (reference call sites: TrainableSDPipeline.py:144-150 `self.unet(...)`, :220 `self.vae.decode`,
 * concept_mat_utils/caption_blip.py:57 `self.model(**inputs)`, LoRA: training_utils/pipeline.py:94-115).
 *   plain mode : A = [M, a_k] row-major with leading dimension a_ld.
 *   conv mode  : A = NHWC activation (n_img, H, W, a_k) dense; k-blocks iterate (tap, segment, 64-channel slab),
 *                tap t reads the input at spatial offset (tap_dh[t], tap_dw[t]); out-of-range pixels are zero
 *                (TMA out-of-bounds fill == the convolution's zero padding).  M must equal n_img*H*W, a_k % 64 == 0.
 *   n_seg = 2  : two K-segments accumulate into one output — [x | x.down^T] x [W | up] (LoRA) or
 *                [hidden | skip] x W (torch.cat fused away).  b_koff[s] = column of B where segment s starts
 *                (conv: inside each tap's block of c_total channels).
 *   B          : [N, b_ld] row-major (K-major) 16-bit weights.
 * All base pointers 16-byte aligned, a_ld/b_ld/a_k multiples of 8.  act: 0 none, 1 SiLU, 2 GELU(erf).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t M, N;
  int32_t dtype;            /* COMAT_F16 or COMAT_BF16 (A, B, residual, out16) */
  int32_t n_seg;
  const void* a[2];
  int64_t a_ld[2];
  int32_t a_k[2];
  const void* b[2];
  int64_t b_ld[2];
  int32_t b_koff[2];
  int32_t conv, n_img, H, W, n_taps, c_total;
  int32_t tap_dh[9], tap_dw[9];
  float alpha;
  const float* bias;
  const float* rowvec;
  int32_t rows_per_group;
  int32_t act;
  const void* residual;
  int64_t res_ld;
  void* out16;
  int64_t out_ld;
  float* out32;
  int64_t out32_ld;
  int32_t force_bn;         /* 0 = auto; else 32/64/128/160/256 (tuning / tests) */
  int32_t split_k;          /* >1: split the K loop over grid.z; needs splitk_ws = split_k*M*N floats; a second pass
                               reduces the partials in a fixed order and applies the epilogue */
  int32_t accumulate;       /* out32 += alpha * A.B with vector fp32 atomics from every (K-split) CTA: gradient accumulation over
                               K-splits and over calls in one kernel; fp32 output only, no bias / act / residual; splitk_ws unused */
  float* splitk_ws;
  int64_t rowvec_ld;        /* row pitch of rowvec in floats (0 = N): lets all ResBlock time-embedding projections of a UNet call
                               come from ONE GEMM whose output is sliced per block */
  int32_t a_mn_major;       /* 1: a[0] is stored [K, M] row-major (M contiguous) instead of [M, K]; a_ld = its row pitch, a_k = K.
                               out = A^T-stored . B: the LoRA weight-gradient GEMMs (d up = dy^T t, d down = u^T x of
                               training_utils/pipeline.py:94-115's backward) read dy / t / x as they lie, no transposed copies.
                               Plain mode, one K segment only. */
  int32_t b_mn_major;       /* 1: b[0] is stored [K, N] row-major (N contiguous) */
  int32_t b_dtype;          /* 0 or equal to dtype.  (A descriptor whose B format differs from A's faults on sm_100a, so mixed
                               fp16 x bf16 operands return COMAT_ERR_UNSUPPORTED.) */
  int32_t force_kernel;     /* 0 = auto; 1 = one-tile-per-CTA kernel; 2 = persistent kernel; 3 = CTA-pair (cta_group::2) kernel
                               (tuning table / tests) */
  float* gn_sums;           /* optional: GroupNorm statistics of the OUTPUT, accumulated by the epilogue with fp32 atomics into
                               gn_sums[(image * gn_groups + group) * 2 + {0, 1}] += {sum, sum of squares} of the fp32 results
                               (ZEROED by the caller).  The GroupNorm that consumes `out` (diffusers ResnetBlock2D.norm1 / norm2,
                               Transformer2DModel.norm, conv_norm_out) then runs comat_groupnorm_fwd_from_sums: no statistics pass
                               over HBM.  Conv mode: image = n; plain mode: image = row / gn_rows_per_image.
                               COMAT_ERR_UNSUPPORTED when comat_gemm_gn_supported() is 0 for the problem. */
  int32_t gn_groups;        /* number of channel groups (N % gn_groups == 0) */
  int32_t gn_rows_per_image;/* plain mode only: rows per image, a multiple of 32 */
} comat_gemm_params;

int comat_gemm(const comat_gemm_params* p, void* stream);
/* 1 if comat_gemm accepts p->gn_sums for this problem, else 0 (p->gn_sums itself is not read): 16-bit TMA-store epilogue, no
 * split-K / accumulate / GEGLU, and a 32-row accumulator quarter never straddles two images (conv tiles with >= 32 pixels per
 * image, plain GEMMs with gn_rows_per_image % 32 == 0). */
int comat_gemm_gn_supported(const comat_gemm_params* p);

/* ------------------------------------------------------------------------------------------------------------
 * HBM-bound normalisation / activation / rearrangement kernels (16-bit activations, fp32 statistics).
 * Replace aten group_norm / silu / layer_norm / gelu / interpolate / cat launched by diffusers' ResnetBlock2D,
 * Transformer2DModel, BasicTransformerBlock, GEGLU, Upsample2D/Downsample2D.
 * Activations are row-major [rows, C]; images are NHWC (n, H*W, C).  dtype = COMAT_F16 / COMAT_BF16.
 * Base weights are frozen on this path (training_utils/pipeline.py:66-71): no gamma/beta gradients.
 * ------------------------------------------------------------------------------------------------------------ */
size_t comat_groupnorm_workspace_floats(int n, int HW, int G);
/* y = [silu](GroupNorm(x)); saves (mean, rstd) per (n, group) in mean_rstd[n*G*2] */
int comat_groupnorm_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mean_rstd, float* ws,
                        int n, int HW, int C, int G, float eps, int silu, int dtype, void* stream);
/* same result from statistics a producing comat_gemm accumulated (comat_gemm_params.gn_sums): sums[n*G*2] = (sum, sum of
 * squares) per (image, group); one pass over x (read once, written once) */
int comat_groupnorm_fwd_from_sums(const void* x, void* y, const float* gamma, const float* beta, float* mean_rstd,
                                  const float* sums, int n, int HW, int C, int G, float eps, int silu, int dtype, void* stream);
int comat_groupnorm_bwd(const void* x, const void* dy, void* dx, const float* gamma, const float* beta,
                        const float* mean_rstd, float* ws, int n, int HW, int C, int G, int silu, int dtype, void* stream);
int comat_layernorm_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mean_rstd, long long rows,
                        int C, float eps, int dtype, void* stream);
int comat_layernorm_bwd(const void* x, const void* dy, void* dx, const float* gamma, const float* mean_rstd, long long rows,
                        int C, int dtype, void* stream);
/* GEGLU: hg = [hidden | gate] (rows, 2*Ch) -> out = hidden * gelu(gate) (rows, Ch); bwd writes d[hidden | gate] */
int comat_geglu_fwd(const void* hg, void* out, long long rows, int Ch, int dtype, void* stream);
int comat_geglu_bwd(const void* hg, const void* dy, void* dhg, long long rows, int Ch, int dtype, void* stream);
/* op: 0 silu(x) 1 silu'(x)*y 2 gelu(x) 3 gelu'(x)*y 4 x+y 5 alpha*x 6 alpha*x+beta*y ; numel % 8 == 0 */
int comat_elementwise(const void* x, const void* y, void* out, long long numel, int op, float alpha, float beta, int dtype,
                      void* stream);
/* mode 0: nearest x2 upsample (n,H,W,C)->(n,2H,2W,C); 1: its backward (H,W = small dims);
 * mode 2: space-to-depth (n,H,W,C)->(n,H/2,W/2,4C) (channel block = (y&1)*2+(x&1)); 3: inverse (H,W = large dims) */
int comat_spatial(const void* in, void* out, int n, int H, int W, int C, int mode, int dtype, void* stream);
/* 16-bit (R, Cc) -> (Cc, ld_out >= R) */
int comat_transpose16(const void* in, void* out, int R, int Cc, int ld_out, void* stream);
/* row softmax (mode 0: out = softmax(x)) and its backward (mode 1: out = x * (dp - sum(x*dp)), x = probabilities); 16-bit,
 * cols % 8 == 0, cols <= 8192.  The softmax between the two GEMMs of an unfused attention (VAE mid-block, d = 512). */
int comat_softmax_rows(const void* x, const void* dp, void* out, long long rows, int cols, int mode, int dtype, void* stream);
int comat_copy2d16(const void* src, void* dst, long long rows, int cols, long long ld_src, long long ld_dst, void* stream);
/* fp32 NCHW latents -> 16-bit NHWC zero-padded to Cpad channels (x scale), and back (first Cout of ld channels) */
int comat_latent_to_nhwc(const float* in, void* out, int n, int Cin, int HW, int Cpad, float scale, int dtype, void* stream);
int comat_nhwc_to_nchw_f32(const void* in, float* out, int n, int Cout, int HW, int ld, float scale, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused clip + AdamW over ONE flat fp32 buffer (all LoRA A/B matrices), consuming the all-reduced gradient.
 * Replaces accelerator.clip_grad_norm_ + torch.optim.AdamW.step (training_script.py:661-664, :692-694).
 *   comat_grad_sumsq : out[0] = sum g^2 (device scalar, no host sync); partial = >= 1024 floats of scratch.
 *   comat_adamw_clip : g' = g * grad_scale * min(1, max_norm / (sqrt(sumsq) * grad_scale + 1e-6)); then torch-AdamW math
 *                      (decoupled weight decay, bias correction by the DEVICE step counter).  max_norm <= 0 disables clipping.
 *                      Overflow guard (accelerate's GradScaler behaviour, training_script.py:659-663 under mixed_precision=fp16):
 *                      a non-finite sumsq skips the update - p, m, v and counters[0] stay untouched, counters[1] += 1,
 *                      counters[2] = 1.  p, g, m, v 16-byte aligned; state = 4 floats of scratch, counters = 3 ints
 *                      {steps taken, steps skipped, last step skipped}, both device memory owned by the caller.  Hyper-parameters
 *                      are doubles: they are rounded to fp32 where torch.optim.AdamW rounds its Python floats (1 - beta2, lr / bias
 *                      correction ...), so the update matches torch's to fp32 rounding.
 * ------------------------------------------------------------------------------------------------------------ */
int comat_grad_sumsq(const float* g, long long n, float* partial, float* out, void* stream);
int comat_adamw_clip(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2,
                     double eps, double weight_decay, float max_norm, float grad_scale, const float* sumsq,
                     float* state, int* counters, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused multi-head attention forward (tcgen05, flash-style):  out = softmax(scale * q k^T) v  per (sample, head).
 * Replaces the hooked Attention.forward arithmetic (attn_utils/tc_attn_utils.py:126-145: baddbmm + softmax + bmm) and
 * F.scaled_dot_product_attention inside diffusers' Attention / HF BLIP attention.
 *   q (n, Lq, H*d), k/v (n, Lk, H*d), out (n, Lq, H*d): 16-bit token-major.  d in {16, 32, 40, 64, 80, 128, 160}.
 *   probs : optional fp32 (n*H, Lq, Lk) export of the normalised probabilities (Lk <= 128) — the tensor AttentionStore
 *           clones for the attention-map loss (tc_attn_utils.py:60-68); null to skip.
 *   lse   : optional fp32 (n*H, Lq) log-sum-exp (saved for backward); null to skip.
 *   workspace: comat_attention_workspace_bytes() bytes (V^T staging).
 * ------------------------------------------------------------------------------------------------------------ */
size_t comat_attention_workspace_bytes(int n, int Lk, int H, int d);
/* kv_lens: optional int32[n] valid-key counts (padding mask); causal != 0: key index <= query index. */
int comat_attention_fwd(const void* q, const void* k, const void* v, void* out, float* probs, float* lse, void* workspace,
                        int n, int Lq, int Lk, int H, int d, float scale, int dtype, const int* kv_lens, int causal, void* stream);
/* Same, with explicit row pitches (elements) for q, k and v: the operands may be column slices of a wider matrix, e.g.
 * the (tokens, 3*H*d) output of one fused q|k|v projection GEMM (to_q/to_k/to_v of tc_attn_utils.py:113-121 share their
 * input for self-attention) or the cached k|v projection of the text context.  Pitches must be multiples of 8. */
int comat_attention_fwd_strided(const void* q, const void* k, const void* v, void* out, float* probs, float* lse,
                                void* workspace, int n, int Lq, int Lk, int H, int d, long long q_ld, long long k_ld,
                                long long v_ld, float scale, int dtype, const int* kv_lens, int causal,
                                int probs_first_sample, void* stream);
/* probs_first_sample = b0: only samples b >= b0 export their probabilities and `probs` is ((n - b0)*H, Lq, Lk).  The attrcon step
 * (AttrConcenTrainableSDPipeline.py:239-279) captures the maps of the CONDITIONAL half of the classifier-free-guidance batch; with
 * b0 = n/2 both halves run as one UNet call instead of the reference's two half-batch calls. */

/* ------------------------------------------------------------------------------------------------------------
 * Fused classifier-free guidance + DDPM ancestral step on the fp32 latent chain.
 * Replaces TrainableSDPipeline.py:155-167 (chunk, guidance combine, scheduler.step) when guidance_rescale == 0:
 *   out = c_x * x + c_eps * (e_u + s (e_c - e_u)) + sigma * noise ;  eps = [e_u ; e_c] (2n floats) when cfg, else n floats.
 * (c_eps, c_x, sigma) = comat_b200.scheduler.DDPMScheduler.step_coefficients(t).
 * ------------------------------------------------------------------------------------------------------------ */
int comat_cfg_ddpm_step_fwd(const float* eps, const float* x, const float* noise, float* out, long long n, float guidance,
                            float c_eps, float c_x, float sigma, int cfg, void* stream);
int comat_cfg_ddpm_step_bwd(const float* grad_out, float* d_eps, float* dx, long long n, float guidance, float c_eps,
                            float c_x, int cfg, void* stream);

/* Fused attention backward (tcgen05, recomputation; two launches: dQ, then dK+dV; bit-reproducible except for short key sequences
 * - Lk <= 1024 with >= 8 query tiles, i.e. the UNet's cross-attention - where the dK/dV launch splits the query range over several
 * CTAs per key tile and adds the partial sums with fp32 atomics).
 * Inputs as the forward plus o, dO (n, Lq, H*d) 16-bit and the forward's lse.  `probs` + `dp_ext` (both fp32 (n*H, Lq, Lk),
 * Lk <= 128) inject the gradient of the exported probabilities (attention-map loss) into the softmax backward; pass nulls
 * otherwise.  Outputs dq / dk / dv in the layouts of q / k / v. */
size_t comat_attention_bwd_workspace_bytes(int n, int Lq, int Lk, int H, int d);
int comat_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dO, const float* lse,
                        const float* probs, const float* dp_ext, void* dq, void* dk, void* dv, void* workspace, int n,
                        int Lq, int Lk, int H, int d, float scale, int dtype, const int* kv_lens, int causal, void* stream);
/* Same with explicit row pitches (elements) for q, k and v (column slices of a fused q|k|v / k|v projection output);
 * o, dO and the three outputs are contiguous (n, L, H*d). */
int comat_attention_bwd_strided(const void* q, const void* k, const void* v, const void* o, const void* dO, const float* lse,
                                const float* probs, const float* dp_ext, void* dq, void* dk, void* dv, void* workspace, int n,
                                int Lq, int Lk, int H, int d, long long q_ld, long long k_ld, long long v_ld, float scale,
                                int dtype, const int* kv_lens, int causal, int dp_first_sample, void* stream);
/* dp_first_sample = b0: `probs` / `dp_ext` are ((n - b0)*H, Lq, Lk) and belong to samples b >= b0 (see probs_first_sample). */

/* ------------------------------------------------------------------------------------------------------------
 * Discriminator head of the fidelity GAN, fused (SURVEY 8b `gan_head_bce`): per-pixel Linear(C, 1) over the D UNet's noise
 * prediction + BCEWithLogits, mean over n*HW logits.  Replaces permute -> nn.Linear(4, 1) -> nn.BCEWithLogitsLoss of
 * training_utils/gan_sdxl.py:31-34, :84-89 (side 'G': n_zero = 0, all targets 1) and :118-132 (side 'D': n_zero = n/2, the
 * generated half has target 0).  eps (n, C, H, W) fp32 NCHW, C <= 8; w (C), b (1) fp32.
 *   fwd: loss_sum[0] += sum of the per-logit losses (caller zeroes it and divides by n*HW).
 *   bwd: d_eps = dL/d eps (may be null), dw / db ACCUMULATED (fp32 atomics; may be null); gout = device scalar dL_total/dL.
 * ------------------------------------------------------------------------------------------------------------ */
int comat_gan_head_bce_fwd(const float* eps, const float* w, const float* b, float* loss_sum, int n, int C, int HW, int n_zero,
                           void* stream);
int comat_gan_head_bce_bwd(const float* eps, const float* w, const float* b, const float* gout, float* d_eps, float* dw, float* db,
                           int n, int C, int HW, int n_zero, void* stream);

/* fp32 -> two bf16 tensors with  alpha * src ~= hi + lo  (hi = bf16(alpha*src), lo = bf16(alpha*src - hi)): ~16 mantissa bits at
 * fp32 range.  Used to feed the fp32-accumulated full-size LoRA gradient products  G = dy^T x  to the 16-bit tensor-core GEMMs
 * that project them onto the LoRA factors (d up = G down^T, d down = up^T G; training_utils/pipeline.py:94-115 backward). */
int comat_split_f32_bf16x2(const float* src, void* hi, void* lo, long long n, float alpha, void* stream);

/* Separable table-driven 2-D resampling with a fused per-channel affine (fp32 NCHW):
 *   out[b,c,oy,ox] = scale[c] * sum_{ky<yc[oy]} sum_{kx<xc[ox]} wy[oy*KY+ky] * wx[ox*KX+kx] * in[b,c,ys[oy]+ky,xs[ox]+kx] + shift[c]
 * With aten's bicubic-antialias taps this is Resize(384, BICUBIC, antialias) + Normalize of caption_blip.py:33-36,45;
 * with the transposed tables it is its backward.  scale / shift may be null. */
int comat_resample2d(const float* in, float* out, const int* ys, const int* yc, const float* wy, const int* xs, const int* xc,
                     const float* wx, const float* scale, const float* shift, int B, int C, int IH, int IW, int OH, int OW,
                     int KY, int KX, void* stream);

/* Cross-entropy with label smoothing, mean over rows whose label != ignore_index — the caption NLL of
 * HF BlipTextLMHeadModel (modeling_blip_text.py:764-775) behind concept_mat_utils/caption_blip.py:57.
 *   fwd : logits fp32 (R, ld >= V), labels int64 (R) -> row_stats (R,2) = {lse, row loss}, out2 = {mean loss, #valid rows}
 *   bwd : dlogits16 (R, Vpad) 16-bit, zero padded = grad_out[0] / #valid * (softmax - (1-eps) onehot - eps/V) */
int comat_ce_label_smooth_fwd(const float* logits, const long long* labels, float* row_stats, float* out2, int R, int V,
                              long long ld, float eps, long long ignore_index, void* stream);
int comat_ce_label_smooth_bwd(const float* logits, const long long* labels, const float* row_stats, const float* out2,
                              const float* grad_out, void* dlogits16, int R, int V, int Vpad, long long ld, float eps,
                              long long ignore_index, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COMAT_B200_H */
