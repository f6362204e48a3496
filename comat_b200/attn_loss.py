"""Host mirror of the reference's attention-map loss interface, backed by the fused CUDA kernels.

Mirrors (same names, argument meaning, return values):
  * ``get_grounding_loss_by_layer``  — attn_utils/tc_loss_utils.py:66-173
  * ``get_mask_loss``                — attr_concen_utils/gsam_interface.py:140-228 (GSAM masks and the spaCy/CLIP
    token-index lists are inputs here: SURVEY.md 2.1 #8/#10 mark their producers out of scope)
  * ``update_nouns_attributes``      — gsam_interface.py:232-261 (pure host logic)

All samples x timesteps x layers of a step are reduced by ONE forward launch + one tiny finalize launch
(C ABI: comat_attnmap_loss_fwd / _bwd, include/comat_b200.h); the backward writes each dP exactly once.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib

TILE_PX, MAX_PAIRS, MAX_WORDS = 32, 32, 16

_INVALID_NOUNS = set(
    "scene surface area atmosphere noise place kitchen dream interior exterior meal background bathroom room scent "
    "street hillside mountain sky sea ocean lost language skill one night day morning space environment conditions "
    "field shore restroom party grass snow meadow water shadow waves song cycle sunlight mysteries wall salon range "
    "cry speech tone thing about activity air advertisement airport also".split())


def update_nouns_attributes(nouns, attributes):
    """gsam_interface.py:232-261: drop nouns occurring more than once, then stop-listed nouns (also plural-stripped)."""
    keep = [(n, a) for n, a in zip(nouns, attributes) if nouns.count(n) == 1]
    keep = [(n, a) for n, a in keep if n not in _INVALID_NOUNS and n[:-1] not in _INVALID_NOUNS]
    return [n for n, _ in keep], [a for _, a in keep]


def words_from_subtrees(subtree_indices, idx_to_wp):
    """gsam_interface.py:163-196: (modifier..., noun) groups -> (noun strings, per-noun attribute token lists)."""
    nouns, attrs = [], []
    for st in subtree_indices:
        if len(st) < 1:
            continue
        noun_idx = st[-1] if isinstance(st[-1], list) else [st[-1]]
        nouns.append("".join(idx_to_wp[i] for i in noun_idx))
        a = []
        for e in st[:-1]:
            a.extend(e if isinstance(e, list) else [e])
        a.extend(noun_idx)
        attrs.append(a)
    if nouns:
        nouns, attrs = update_nouns_attributes(nouns, attrs)
    return nouns, attrs


def mask_resize_any(masks_u8: torch.Tensor, res: int) -> torch.Tensor:
    """(n,H,W) uint8/bool -> (n,res,res) float 0/1; tc_loss_utils.py:88-94 (antialias Resize of a bool, then > 0)."""
    _lib.require_cuda(masks_u8)
    m = masks_u8.to(torch.uint8).contiguous()
    n, h, w = m.shape
    out = torch.empty(n, res, res, dtype=torch.float32, device=m.device)
    _lib.check(_lib.lib().comat_mask_resize_any(_lib.ptr(m), _lib.ptr(out), n, h, w, res, _lib.stream_ptr()), "mask_resize_any")
    _lib.count_launch()
    return out


class AttnMapLossPlan:
    """Device tables for one fused loss evaluation (format documented in include/comat_b200.h)."""

    def __init__(self, groups: List[dict], words_per_sample: List[List[List[int]]],
                 masks_per_sample: List[Optional[List[torch.Tensor]]], device, tokens: int = 77):
        # groups: [{"res": r, "maps": [tensor (B*H, r, r, T) ...]}]
        self.device = device
        B = len(words_per_sample)
        self.B = B
        smp, pair, word_ntok, mask_list = [], [], [], []
        for b in range(B):
            words = words_per_sample[b] if masks_per_sample[b] is not None else []
            if len(words) > MAX_WORDS:
                raise _lib.ComatError(f"sample {b}: {len(words)} words > {MAX_WORDS}")
            pb, wb = len(pair), len(word_ntok)
            for i, toks in enumerate(words):
                if len(toks) == 0:
                    raise _lib.ComatError("empty token list for a word")
                word_ntok.append(len(toks))
                for p in toks:
                    if not (0 <= int(p) < tokens):
                        raise _lib.ComatError(f"token index {p} outside [0,{tokens})")
                    pair.append((wb + i, int(p)))
                mask_list.append(masks_per_sample[b][i].reshape(masks_per_sample[b][i].shape[-2:]))
            if len(pair) - pb > MAX_PAIRS:
                raise _lib.ComatError(f"sample {b}: {len(pair) - pb} (word, token) pairs > {MAX_PAIRS}")
            smp.append((pb, len(pair), wb, len(word_ntok)))
        self.n_words, self.n_pairs = len(word_ntok), len(pair)
        self.empty = self.n_pairs == 0 or len(groups) == 0
        self.maps: List[torch.Tensor] = []
        if self.empty:
            return
        masks_u8 = torch.stack([m.to(device=device, dtype=torch.uint8) for m in mask_list])        # (n_words, Hm, Wm)
        res_set = sorted({g["res"] for g in groups})
        mask_off, off, parts = {}, 0, []
        for r in res_set:
            mask_off[r] = off
            parts.append(mask_resize_any(masks_u8, r).reshape(-1))
            off += self.n_words * r * r
        self.masks = torch.cat(parts)
        grp, work, map_ptrs = [], [], []
        pred_off, max_h, max_mg = 0, 1, 1
        for gi, g in enumerate(groups):
            r = g["res"]
            maps = []
            for m in g["maps"]:
                if m.dtype != torch.float32 or not m.is_cuda:
                    raise _lib.ComatError("attention maps must be fp32 CUDA tensors (SURVEY A.5)")
                mc = m.contiguous()
                if mc.data_ptr() % 16:
                    mc = mc.clone()
                if mc.shape[0] % B or mc.shape[1] != r or mc.shape[2] != r or mc.shape[3] != tokens:
                    raise _lib.ComatError(f"map shape {tuple(m.shape)} incompatible with B={B}, res={r}, T={tokens}")
                maps.append(mc)
            H = maps[0].shape[0] // B
            if any(m.shape[0] // B != H for m in maps):
                raise _lib.ComatError("maps of one group must share the head count")
            if (r * r) % TILE_PX:
                raise _lib.ComatError(f"res^2={r * r} must be a multiple of {TILE_PX}")
            n_tiles = r * r // TILE_PX
            mb = len(self.maps)
            self.maps.extend(maps)
            grp.append((r, mb, mb + len(maps), H, mask_off[r], n_tiles, pred_off, len(work)))
            for b in range(B):
                for t in range(n_tiles):
                    work.append((gi, b, t, 0))
            pred_off += B * MAX_WORDS * r * r
            max_h, max_mg = max(max_h, H), max(max_mg, len(maps))
        self.n_groups, self.n_work = len(groups), len(work)
        i32 = lambda x: torch.tensor(x, dtype=torch.int32).reshape(-1)
        table = torch.cat([i32(grp), i32(smp), i32(pair), i32(word_ntok), i32(work)]).to(device, non_blocking=True)
        o = 0
        self._tables = table
        sizes = [len(grp) * 8, B * 4, self.n_pairs * 2, self.n_words, self.n_work * 4]
        offs = []
        for s in sizes:
            offs.append(o)
            o += s
        self._off = offs
        # the gradient of every map lives in ONE pre-allocated buffer whose pointer table is uploaded here together with the map
        # pointers: the backward is then a single kernel launch - no allocation, no host-to-device copy (r01: 20 empty_like + a
        # pageable pointer-table upload around a 74 us kernel cost 300 us, profiles/r01_attnmap_bench_v3.json)
        sizes_el = [m.numel() for m in self.maps]
        self.grad_flat = torch.empty(sum(sizes_el), dtype=torch.float32, device=device)
        self.grad_views, goff = [], 0
        for m, k in zip(self.maps, sizes_el):
            self.grad_views.append(self.grad_flat[goff:goff + k].view(m.shape))
            goff += k
        ptrs = torch.tensor([m.data_ptr() for m in self.maps] + [g.data_ptr() for g in self.grad_views], dtype=torch.int64)
        self._ptrs = ptrs.to(device, non_blocking=True)
        self.map_ptr, self.grad_ptr = self._ptrs[:len(self.maps)], self._ptrs[len(self.maps):]
        p = _lib.AttnmapPlan()
        p.n_maps, p.n_groups, p.n_samples, p.n_words = len(self.maps), self.n_groups, B, self.n_words
        p.n_pairs, p.n_work, p.tokens = self.n_pairs, self.n_work, tokens
        p.max_heads, p.max_maps_per_group, p.pred_floats = max_h, max_mg, pred_off
        base = table.data_ptr()
        p.map_ptr = self.map_ptr.data_ptr()
        p.grp, p.smp, p.pair, p.word_ntok, p.work = (base + 4 * x for x in offs)
        p.masks = self.masks.data_ptr()
        self.c = p
        self.state_floats = int(_lib.lib().comat_attnmap_loss_state_floats(C.byref(p)))
        self.algorithmic_bytes = sum(m.numel() * 4 for m in self.maps)


_counters: Dict[torch.device, torch.Tensor] = {}


def _counter(device):
    if device not in _counters:
        _counters[device] = torch.zeros(1, dtype=torch.int32, device=device)
    return _counters[device]


class _AttnMapLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: AttnMapLossPlan, *maps):
        dev = plan.device
        loss2 = torch.empty(2, dtype=torch.float32, device=dev)
        state = torch.empty(plan.state_floats, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().comat_attnmap_loss_fwd(C.byref(plan.c), _lib.ptr(loss2), _lib.ptr(state), plan.state_floats,
                                                     _lib.ptr(_counter(dev)), _lib.stream_ptr()), "attnmap_loss_fwd")
        _lib.count_launch(3)
        ctx.plan, ctx.state = plan, state
        ctx.shapes = [m.shape for m in maps]
        return loss2

    @staticmethod
    def backward(ctx, g2):
        plan = ctx.plan
        g2 = g2.contiguous().float()
        views, grad_ptr = plan.grad_views, plan.grad_ptr
        if getattr(ctx, "backward_ran", False):
            # a second backward through the SAME forward (retain_graph): the first call's gradients may still be alive, so this
            # one gets its own buffer + pointer table (slow path; the training step differentiates each forward once).  A plan
            # reused by a later forward hands out the same buffer again: its gradients are valid until that plan's next backward.
            flat = torch.empty_like(plan.grad_flat)
            views, goff = [], 0
            for m in plan.maps:
                views.append(flat[goff:goff + m.numel()].view(m.shape))
                goff += m.numel()
            grad_ptr = torch.tensor([g.data_ptr() for g in views], dtype=torch.int64).to(plan.device)
        ctx.backward_ran = True
        _lib.check(_lib.lib().comat_attnmap_loss_bwd(C.byref(plan.c), _lib.ptr(g2), _lib.ptr(ctx.state), _lib.ptr(grad_ptr),
                                                     _lib.stream_ptr()), "attnmap_loss_bwd")
        _lib.count_launch()
        return (None, *[g.reshape(s) for g, s in zip(views, ctx.shapes)])


def fused_attnmap_loss(plan: AttnMapLossPlan, maps_for_grad: Sequence[torch.Tensor]):
    """returns tensor[2] = (token_loss, pixel_loss); ``maps_for_grad`` are the autograd inputs matching plan.maps."""
    return _AttnMapLossFn.apply(plan, *maps_for_grad)


def get_mask_loss(attn_map: Dict[str, Dict[str, List[torch.Tensor]]], words_per_sample, masks_per_sample,
                  train_layer_ls: Sequence[str], tokens: int = 77):
    """Fused equivalent of GsamSegModel.get_mask_loss (gsam_interface.py:140-228) for pre-computed masks.

    attn_map:           {str(timestep): {'up_16': [ (B*heads,res,res,T) fp32 ...], ...}} as stored by the pipeline
    words_per_sample:   per sample, the surviving nouns' attribute token lists ([[tok...], ...])
    masks_per_sample:   per sample, one (1,1,H,W) bool mask per surviving noun, or None when the sample is skipped
    returns (token_loss, pixel_loss) — both already divided by B (:225-226).
    """
    some = next(iter(next(iter(attn_map.values())).values()))[0] if attn_map else None
    if some is None:
        raise _lib.ComatError("empty attention dict")
    groups = []
    for tkey in attn_map:
        for layer in train_layer_ls:
            groups.append({"res": int(layer.split("_")[1]), "maps": attn_map[tkey][layer]})
    plan = AttnMapLossPlan(groups, words_per_sample, masks_per_sample, some.device, tokens)
    if plan.empty:
        z = some.new_zeros(())
        return z, z.clone()
    flat = [m for g in groups for m in g["maps"]]
    out = fused_attnmap_loss(plan, flat)
    return out[0], out[1]


def get_grounding_loss_by_layer(_gt_seg_list, word_token_idx_ls, res, input_attn_map_ls, is_training_sd21=False):
    """Drop-in for tc_loss_utils.py:66-173 (one sample, one resolution).  input_attn_map_ls: [(heads,res,res,T)]."""
    if is_training_sd21:
        res = int(1.5 * res)
    if len(word_token_idx_ls) == 0:
        return {"token_loss": 0, "pixel_loss": 0}                      # :77-81
    dev = input_attn_map_ls[0].device
    T = input_attn_map_ls[0].shape[-1]
    plan = AttnMapLossPlan([{"res": res, "maps": list(input_attn_map_ls)}], [list(word_token_idx_ls)],
                           [list(_gt_seg_list)], dev, T)
    out = fused_attnmap_loss(plan, list(input_attn_map_ls))
    return {"token_loss": out[0], "pixel_loss": out[1]}


class SegModel:
    """Call-compatible stand-in for ``GsamSegModel`` (attr_concen_utils/gsam_interface.py:22-53, :140-228) as training_script.py:631-637
    uses it:  ``seg_model.get_mask_loss(images, prompt, all_subtree_indices, attn_map_idx_to_wp_all, attn_map)
    -> (token_loss, pixel_loss, grounding_loss_dict)``.  The noun / attribute bookkeeping is the reference's (:163-196, :232-261);
    the mask producer (Grounded-SAM, an external model) is injected as ``mask_fn(image (3,H,W), nouns) -> [ (1,1,H,W) bool per noun ]
    | None``; every (sample, timestep, layer) term then goes through ONE fused kernel launch instead of the Python triple loop."""

    def __init__(self, train_layer_ls: Sequence[str], mask_fn, tokens: int = 77):
        self.train_layer_ls, self.mask_fn, self.tokens = list(train_layer_ls), mask_fn, tokens

    update_nouns_attributes = staticmethod(update_nouns_attributes)

    def get_mask(self, image, nouns):
        return self.mask_fn(image, nouns)

    def get_mask_loss(self, images, prompt, all_subtree_indices, attn_map_idx_to_wp_all, attn_map):
        images = images.detach()
        words, masks = [], []
        for idx, subtree_indices in enumerate(all_subtree_indices):
            nouns, attrs = words_from_subtrees(subtree_indices, attn_map_idx_to_wp_all[idx])
            m = self.get_mask(images[idx], nouns) if nouns else None                # :188-202: no nouns / no mask -> sample skipped
            words.append(attrs if m is not None else [])
            masks.append(list(m) if m is not None else None)
        token_loss, pixel_loss = get_mask_loss(attn_map, words, masks, self.train_layer_ls, self.tokens)
        # the reference also returns per-(timestep, layer) sums that nothing reads (training_script.py:633 discards them); the fused
        # kernel reduces all terms on the device, so the dictionary carries the two totals only
        return token_loss, pixel_loss, {"token/total": token_loss.detach(), "pixel/total": pixel_loss.detach()}
