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

#ifdef __cplusplus
}
#endif
#endif /* COMAT_B200_H */
