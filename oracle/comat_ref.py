"""Restatement (plain PyTorch, fp32-friendly) of the CoMat-owned arithmetic on the hot path.

ORACLE / TEST INFRASTRUCTURE — see oracle/__init__.py.  Every function cites the reference
file:line it follows.  Pinned against the reference's own modules by
``oracle/pin_against_reference.py`` (goldens in tests/golden/).

Differences from the reference that are *interface only* (arithmetic identical):
  * noise is injected (``noises[i]``) instead of drawn from the global RNG;
  * GSAM masks and spaCy/CLIP token-index lists are inputs (out of scope);
  * BLIP token ids are inputs (no tokenizer vocabulary on disk).
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from torchvision.transforms import InterpolationMode
from torchvision.transforms import functional as TVF

from .sd_modules import Attention, rescale_noise_cfg

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


# --------------------------------------------------------------------------------------------
# attention capture  (attn_utils/tc_attn_utils.py:17-216)
# --------------------------------------------------------------------------------------------
class AttentionStore:
    """tc_attn_utils.py:17-94 — counts hooked calls, clones cross-attn probabilities of the places named by
    ``train_layer_ls`` ('mid_8', 'up_16', ...), swaps step_store -> attention_store after a full UNet pass."""

    def __init__(self, train_layer_ls: Sequence[str]):
        self.train_layer_place = sorted({s.split("_")[0] for s in train_layer_ls})
        self.num_att_layers = -1
        self.reset()

    @staticmethod
    def _empty():
        return {f"{p}_{k}": [] for k in ("cross", "self") for p in ("down", "mid", "up")}

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0
        self.step_store = self._empty()
        self.attention_store = {}

    def __call__(self, probs, is_cross: bool, place: str):
        if is_cross and place in self.train_layer_place:          # :60-68
            self.step_store[f"{place}_cross"].append(probs.clone())
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers:             # :37-41, :70-74
            self.cur_att_layer = 0
            self.attention_store = self.step_store
            self.step_store = self._empty()
        return probs

    def get_average_attention(self):                              # :76-79 (no averaging happens)
        return {k: list(v) for k, v in self.attention_store.items()}


class _CaptureProcessor:
    """Same arithmetic as the hooked forward tc_attn_utils.py:104-161 (explicit softmax(QK^T)V, P handed to the
    controller iff it requires grad)."""

    def __init__(self, controller, place):
        self.controller, self.place = controller, place

    def __call__(self, attn: Attention, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        is_cross = encoder_hidden_states is not None
        residual = hidden_states
        nd = hidden_states.ndim
        if nd == 4:
            b, c, h, w = hidden_states.shape
            hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
        if attn.group_norm is not None:
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        q = attn.head_to_batch_dim(attn.to_q(hidden_states))
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = attn.head_to_batch_dim(attn.to_k(ctx))
        v = attn.head_to_batch_dim(attn.to_v(ctx))
        probs = attn.get_attention_scores(q, k, None)
        if probs.requires_grad:                                   # :142-143
            probs = self.controller(probs, is_cross, self.place)
        out = attn.batch_to_head_dim(torch.bmm(probs, v))
        out = attn.to_out[0](out)
        if nd == 4:
            out = out.transpose(-1, -2).reshape(b, c, h, w)
        if attn.residual_connection:
            out = out + residual
        return out / attn.rescale_output_factor


def register_attention_control(unet, controller) -> int:
    """tc_attn_utils.py:96-196: every ``Attention`` under a top-level child whose name contains down/up/mid."""
    count = 0
    for name, child in unet.named_children():
        place = "down" if "down" in name else "up" if "up" in name else "mid" if "mid" in name else None
        if place is None:
            continue
        for m in child.modules():
            if m.__class__.__name__ == "Attention":
                m.processor = _CaptureProcessor(controller, place)
                count += 1
    controller.num_att_layers = count
    return count


def get_cross_attn_map_from_unet(store: AttentionStore, reses=(64, 32, 16, 8), poses=("down", "mid", "up")):
    """tc_attn_utils.py:198-216: regroup stored (B*heads, HW, T) by res=sqrt(HW) -> '{pos}_{res}': [(B*heads,res,res,T)]."""
    maps = store.get_average_attention()
    out = {}
    for pos in poses:
        for res in reses:
            sel = [m.reshape(-1, res, res, m.shape[-1]) for m in maps[f"{pos}_cross"] if m.shape[1] == res * res]
            if sel:
                out[f"{pos}_{res}"] = sel
    return out


# --------------------------------------------------------------------------------------------
# attention-map token / pixel loss   (attn_utils/tc_loss_utils.py:66-173)
# --------------------------------------------------------------------------------------------
def resize_mask(mask_bool: torch.Tensor, res: int) -> torch.Tensor:
    """tc_loss_utils.py:88-94: torchvision Resize(antialias=True) of a (1,1,H,W) mask, then ``> 0`` -> float
    (1,res,res).  (On a bool tensor torchvision round-trips through float and casts back: a dilation.)"""
    m = TVF.resize(mask_bool, [res, res], antialias=True)
    return (m.squeeze(0) > 0.0).float()


def grounding_loss_by_layer(masks: List[torch.Tensor], word_token_idx_ls: List[List[int]], res: int,
                            attn_maps: List[torch.Tensor]):
    """tc_loss_utils.py:66-173.  attn_maps: list of (heads,res,res,T); masks: list of (1,1,H,W) bool per word."""
    W = len(word_token_idx_ls)
    if W == 0:
        return {"token_loss": 0, "pixel_loss": 0}                 # :77-81
    m = [resize_mask(x, res) for x in masks]                      # (1,res,res)
    token = 0.0
    for amap in attn_maps:                                        # :104-125  (summed over maps)
        h = amap.shape[0]
        for i, toks in enumerate(word_token_idx_ls):
            obj = 0.0
            for p in toks:
                a = amap[..., p]
                frac = (a * m[i]).reshape(h, -1).sum(-1) / a.reshape(h, -1).sum(-1)
                obj = obj + (1.0 - frac.mean()) ** 2
            token = token + obj / len(toks)
    token = token / W
    avg = torch.stack([a.reshape(-1, res, res, a.shape[-1]).mean(0) for a in attn_maps], 0)   # :131-139
    avg = (avg.sum(0) / avg.shape[0]).unsqueeze(0)
    pixel = 0.0
    for i, toks in enumerate(word_token_idx_ls):                  # :144-167
        pred = torch.stack([avg[..., p] for p in toks], 0).sum(0)
        pixel = pixel + F.binary_cross_entropy(pred, m[i])
    return {"token_loss": token, "pixel_loss": pixel / W}


def mask_loss(attn_dict: Dict[str, Dict[str, List[torch.Tensor]]], words_per_sample: List[List[List[int]]],
              masks_per_sample: List[Optional[List[torch.Tensor]]], train_layer_ls: Sequence[str], ref_tensor):
    """attr_concen_utils/gsam_interface.py:140-228 with nouns/attribute lists already reduced to
    ``words_per_sample[b] = [[token positions of word0], ...]`` and masks injected (None = sample skipped,
    :188-202).  Divides by B = number of samples, skipped or not (:225-226)."""
    bs = len(words_per_sample)
    token = ref_tensor.new_zeros(())
    pixel = ref_tensor.new_zeros(())
    for b in range(bs):
        words, masks = words_per_sample[b], masks_per_sample[b]
        if not words or masks is None:
            continue
        for tkey in attn_dict:
            for layer in train_layer_ls:
                res = int(layer.split("_")[1])
                maps = [x.reshape(bs, x.shape[0] // bs, *x.shape[1:])[b] for x in attn_dict[tkey][layer]]  # :158-159
                d = grounding_loss_by_layer(masks, words, res, maps)
                token = token + d["token_loss"]
                pixel = pixel + d["pixel_loss"]
    return token / bs, pixel / bs


def words_from_subtrees(subtree_indices, idx_to_wp, update_fn=None):
    """gsam_interface.py:163-196: (modifier..., noun) groups -> (noun strings, attribute token lists)."""
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
    if update_fn is not None and nouns:
        nouns, attrs = update_fn(nouns, attrs)
    return nouns, attrs


# --------------------------------------------------------------------------------------------
# BLIP concept-matching reward   (concept_mat_utils/caption_blip.py:43-59)
# --------------------------------------------------------------------------------------------
def blip_preprocess(images: torch.Tensor) -> torch.Tensor:
    """caption_blip.py:33-36,45: per-image Resize((384,384), BICUBIC, antialias) + Normalize(CLIP mean/std)."""
    out = []
    for im in images:
        x = TVF.resize(im, [384, 384], interpolation=InterpolationMode.BICUBIC, antialias=True)
        out.append(TVF.normalize(x, CLIP_MEAN, CLIP_STD))
    return torch.stack(out)


def blip_labels(input_ids: torch.Tensor, prompt_length: int, pad_id: int = 0) -> torch.Tensor:
    """caption_blip.py:51-54."""
    labels = input_ids.masked_fill(input_ids == pad_id, -100)
    labels[:, :prompt_length] = -100
    return labels


def blip_score(blip_model, images, input_ids, attention_mask, prompt_length: int = 4, pad_id: int = 0,
               preprocess=True):
    """reward = -(mean CE over un-ignored shifted tokens of the whole batch) (caption_blip.py:56-58)."""
    pix = blip_preprocess(images) if preprocess else images
    labels = blip_labels(input_ids, prompt_length, pad_id)
    out = blip_model(pixel_values=pix, input_ids=input_ids, attention_mask=attention_mask, labels=labels)
    return -out.loss


def make_blip(large=True, label_smoothing=0.1, seed=0, layers=None, dtype=torch.float32):
    """Random-init HF BlipForConditionalGeneration at blip-image-captioning-large geometry, with a
    non-degenerate init (default vision initializer_range 1e-10 gives ~0 image gradients).  ``label_smoothing``
    defaults to 0.1 = the transformers==4.31.0 pin (requirements.txt:1); 5.x reads it from the config."""
    from transformers import BlipConfig, BlipForConditionalGeneration
    if large:
        vis = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                   image_size=384, patch_size=16, initializer_range=0.02)
        txt = dict(hidden_size=768, encoder_hidden_size=1024, intermediate_size=3072, num_hidden_layers=12,
                   num_attention_heads=12, vocab_size=30524, max_position_embeddings=512)
    else:
        vis = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                   image_size=384, patch_size=16, initializer_range=0.02)
        txt = dict(hidden_size=64, encoder_hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                   num_attention_heads=2, vocab_size=30524, max_position_embeddings=512)
    if layers is not None:
        vis["num_hidden_layers"], txt["num_hidden_layers"] = layers
    txt.update(label_smoothing=label_smoothing, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
               bos_token_id=30522, pad_token_id=0, sep_token_id=102)
    vis.update(attention_dropout=0.0)
    cfg = BlipConfig(vision_config=vis, text_config=txt)
    cfg.label_smoothing = label_smoothing
    torch.manual_seed(seed)
    m = BlipForConditionalGeneration(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.startswith("vision_model") and p.ndim >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    m.eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m.to(dtype)


# --------------------------------------------------------------------------------------------
# rollout   (TrainableSDPipeline.py:20-225, AttrConcenTrainableSDPipeline.py:38-279)
# --------------------------------------------------------------------------------------------
def select_training_steps(S: int, K: int, rng: random.Random, n_attrcon: int = 2):
    """training_script.py:563-566, :589-590 (random.choices = with replacement)."""
    interval = S // K
    max_start = S - interval * (K - 1) - 1
    start = rng.randint(0, max_start)
    steps = list(range(start, S, interval))
    attrcon = rng.choices(steps, k=min(n_attrcon, len(steps)))
    return steps, attrcon


def rollout(unet, vae, scheduler, prompt_embeds, negative_prompt_embeds, latents, noises, S: int,
            training_timesteps: Sequence[int], guidance_scale=7.5, guidance_rescale=0.0,
            attrcon_train_steps: Optional[Sequence[int]] = None, controller: Optional[AttentionStore] = None,
            added_cond_kwargs=None, sdxl=False, return_latents=False, decode=True):
    """SD1.5: TrainableSDPipeline.py:132-225 with the trainer's constants (bp_on_trained=True, early_exit=False,
    double_laststep=False, fast_training=False, detach_gradient=True; training_script.py:558-567).
    SDXL (``sdxl=True``): TrainableSDPipeline.py:799-846 — UNet input always detached (:809) and the image is
    returned un-rescaled when ``return_latents`` (:838-840).
    ``noises[i]`` is the DDPM variance noise of step i.  Returns (image, latents, attn_dict)."""
    cfg = guidance_scale > 1.0
    T = list(training_timesteps)
    embeds = torch.cat([negative_prompt_embeds, prompt_embeds]) if cfg else prompt_embeds
    scheduler.set_timesteps(S)
    attn_dict = {}
    prev = torch.is_grad_enabled()
    try:
        for i, t in enumerate(scheduler.timesteps):
            torch.set_grad_enabled(len(T) == 0 or i > min(T))                      # :133
            x_in = torch.cat([latents] * 2) if cfg else latents
            torch.set_grad_enabled(i in T)                                         # :138
            detach = sdxl or not (i in T)                                          # :140-145 / :809
            x_in = x_in.detach() if detach else x_in
            kw = dict(encoder_hidden_states=embeds, return_dict=False)
            if added_cond_kwargs is not None:
                kw["added_cond_kwargs"] = added_cond_kwargs
            if i in T and attrcon_train_steps is not None and i in attrcon_train_steps and controller is not None:
                # AttrConcenTrainableSDPipeline.py:239-279: cond half with capture, then uncond half
                h = x_in.shape[0] // 2
                kc, ku = dict(kw), dict(kw)
                kc["encoder_hidden_states"], ku["encoder_hidden_states"] = embeds[h:], embeds[:h]
                if added_cond_kwargs is not None:                                  # SDXL :459-463
                    kc["added_cond_kwargs"] = {k: v[h:] for k, v in added_cond_kwargs.items()}
                    ku["added_cond_kwargs"] = {k: v[:h] for k, v in added_cond_kwargs.items()}
                controller.reset()
                n_c = unet(x_in[h:], t, **kc)[0]
                attn_dict[str(int(t))] = get_cross_attn_map_from_unet(controller)
                controller.reset()
                n_u = unet(x_in[:h], t, **ku)[0]
                controller.reset()
                eps = torch.cat([n_u, n_c], 0)
            else:
                eps = unet(x_in, t, **kw)[0]
            eps = eps.to(embeds.dtype)
            if cfg:
                e_u, e_c = eps.chunk(2)
                eps = e_u + guidance_scale * (e_c - e_u)                           # :155-157
                if guidance_rescale > 0.0:
                    eps = rescale_noise_cfg(eps, e_c, guidance_rescale)
            torch.set_grad_enabled(len(T) == 0 or i >= min(T))                     # :163
            latents = scheduler.step(eps, t, latents, variance_noise=noises[i]).prev_sample
        torch.set_grad_enabled(True)
        image = None
        if decode:
            image = vae.decode(latents.to(vae.dtype) / vae.config.scaling_factor, return_dict=False)[0]
            if not (sdxl and return_latents):
                image = image / 2 + 0.5                                            # :223 ; SDXL quirk :838-840
    finally:
        torch.set_grad_enabled(prev)
    return image, latents, attn_dict


# --------------------------------------------------------------------------------------------
# GAN discriminator   (training_utils/gan_sdxl.py:50-132)
# --------------------------------------------------------------------------------------------
def d_forward(d_unet, d_head, scheduler, latents_fake, null_embed, S: int, side: str, latents_real=None):
    """side='G': BCEWithLogits(head(D_unet(z_fake, t=timesteps[-1], null)), 1)  (:52-89);
    side='D': input cat[z_fake.detach(), z_real], targets [0..,1..]  (:92-132)."""
    scheduler.set_timesteps(S)
    t = scheduler.timesteps[-1]
    if side == "G":
        x, cond = latents_fake, null_embed
    else:
        x = torch.cat([latents_fake.detach(), latents_real])
        cond = torch.cat([null_embed, null_embed])
    eps = d_unet(x, t, encoder_hidden_states=cond, return_dict=False)[0]
    pred = d_head(eps.permute(0, 2, 3, 1).float())
    target = torch.ones_like(pred)
    if side == "D":
        target[: target.shape[0] // 2] = 0
    return F.binary_cross_entropy_with_logits(pred, target)


# --------------------------------------------------------------------------------------------
# one G train-step loss   (training_script.py:556-651)
# --------------------------------------------------------------------------------------------
def g_step_loss(unet, vae, scheduler, blip_model, batch, cfgd, controller=None, d_unet=None, d_head=None):
    """Assemble L = -reward*w + w_gan*G_loss + w_tok*token + w_pix*pixel  (training_script.py:618,625,639-640).
    ``batch`` carries every random draw (latents, noises, offsets, steps) so both paths see the same inputs.
    Returns dict of scalars + 'image' + 'latents'."""
    res = cfgd.get("resolution", 512)
    image, lat, attn_dict = rollout(
        unet, vae, scheduler, batch["prompt_embeds"], batch["null_embeds"], batch["latents"], batch["noises"],
        cfgd["S"], batch["training_steps"], cfgd.get("cfg_scale", 7.5), cfgd.get("cfg_rescale", 0.0),
        batch.get("attrcon_steps"), controller, return_latents=d_unet is not None)
    off = res // 224                                                               # :606-611
    size = res - off
    ox, oy = batch["crop"]
    crop = image[:, :, ox:ox + size, oy:oy + size]
    reward = blip_score(blip_model, crop, batch["blip_ids"], batch["blip_mask"], cfgd.get("prompt_length", 4))
    out = {"Blip": reward, "image": image, "latents": lat}
    loss = -(cfgd.get("blip_weight", 1.0) * reward)
    if d_unet is not None:
        g = d_forward(d_unet, d_head, scheduler, lat, batch["gan_null_embeds"], cfgd["S"], "G")
        out["G_loss"] = g
        loss = loss + cfgd.get("gan_loss_weight", 1.0) * g
    if controller is not None and attn_dict:
        tok, pix = mask_loss(attn_dict, batch["words"], batch["masks"], cfgd["train_layer_ls"], image.detach())
        out["token_loss"], out["pixel_loss"] = tok, pix
        loss = loss + cfgd.get("mask_token_loss_weight", 1e-3) * tok + cfgd.get("mask_pixel_loss_weight", 5e-5) * pix
    out["loss"] = loss
    out["attn_dict"] = attn_dict
    return out


# --------------------------------------------------------------------------------------------
# prompt encoding  
# --------------------------------------------------------------------------------------------
def make_clip_text(which="clip_l", tiny=True, seed=7, dtype=torch.float32, device="cpu", layers=None):
    """random-init HF CLIP text tower (transformers is the un-vendored third-party layer here, used directly like BLIP):
    ``clip_l`` -> CLIPTextModel (quick_gelu), ``bigg`` -> CLIPTextModelWithProjection (gelu)."""
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    if which == "clip_l":
        kw = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, hidden_act="quick_gelu",
                  projection_dim=768)
        cls = CLIPTextModel
    else:
        kw = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20, hidden_act="gelu",
                  projection_dim=1280)
        cls = CLIPTextModelWithProjection
    if tiny:
        kw.update(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, projection_dim=64)
    if layers is not None:
        kw["num_hidden_layers"] = layers
    cfg = CLIPTextConfig(vocab_size=49408, max_position_embeddings=77, eos_token_id=2, bos_token_id=0, pad_token_id=1, **kw)
    torch.manual_seed(seed)
    with torch.device(device):
        m = cls(cfg)
    return m.eval().requires_grad_(False).to(dtype)


@torch.no_grad()
def encode_prompt_sd(text_encoder, tokenizer, prompt, num_images_per_prompt: int, do_cfg: bool, negative_prompt=None,
                     negative_prompt_embeds=None, clip_skip: Optional[int] = None):
    """TrainableSDPipeline.py:227-424 for a frozen, non-DDP encoder whose config has no ``use_attention_mask``
    (the SD1.5 CLIP-L): tokenise to model_max_length (:291-297), ``text_encoder(ids)[0]`` (:324-327) or the
    clip_skip branch with the final LayerNorm re-applied (:328-340), repeat per image (:353-356), '' as the
    default negative prompt tokenised to the prompt's length (:359-404), repeated the same way (:406-413)."""
    prompts = [prompt] if isinstance(prompt, str) else list(prompt)
    ids = tokenizer(prompts, padding="max_length", max_length=tokenizer.model_max_length, truncation=True,
                    return_tensors="pt").input_ids
    if clip_skip is None:
        pe = text_encoder(ids, attention_mask=None)[0]
    else:
        out = text_encoder(ids, attention_mask=None, output_hidden_states=True)
        pe = text_encoder.text_model.final_layer_norm(out[-1][-(clip_skip + 1)])
    b, L, _ = pe.shape
    pe = pe.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, L, -1)
    npe = negative_prompt_embeds
    if do_cfg and npe is None:
        if negative_prompt is None:
            uncond = [""] * b
        elif isinstance(negative_prompt, str):
            uncond = [negative_prompt]
        else:
            uncond = list(negative_prompt)
        nids = tokenizer(uncond, padding="max_length", max_length=L, truncation=True, return_tensors="pt").input_ids
        npe = text_encoder(nids, attention_mask=None)[0]
    if do_cfg:
        npe = npe.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, npe.shape[1], -1)
    return pe, npe, ids


@torch.no_grad()
def encode_prompt_sdxl(text_encoder, text_encoder_2, tokenizer, tokenizer_2, prompt, num_images_per_prompt: int, do_cfg: bool,
                       negative_prompt=None, force_zeros_for_empty_prompt: bool = True, clip_skip: Optional[int] = None):
    """diffusers (0.22-0.25) ``StableDiffusionXLPipeline.encode_prompt`` - inherited un-overridden by the reference's
    TrainableSDXLPipeline (TrainableSDPipeline.py:427; called at training_script.py:521,573) and NOT on disk: restated from
    the published behaviour, **parity unpinned**.  Per encoder: ``out = enc(ids, output_hidden_states=True)``,
    ``pooled = out[0]`` (the projection encoder's ``text_embeds`` wins), ``hidden_states[-2]`` (or ``-(clip_skip + 2)``),
    concatenated on the feature axis; negatives are zeros when ``negative_prompt is None and force_zeros_for_empty_prompt``,
    else the encoding of '' / the negative prompt."""
    prompts = [prompt] if isinstance(prompt, str) else list(prompt)
    b = len(prompts)

    def enc_pair(texts, max_len, skip):
        outs, pooled = [], None
        for tok, enc in ((tokenizer, text_encoder), (tokenizer_2, text_encoder_2)):
            ids = tok(texts, padding="max_length", max_length=max_len or tok.model_max_length, truncation=True,
                      return_tensors="pt").input_ids
            out = enc(ids, output_hidden_states=True)
            pooled = out[0]
            outs.append(out.hidden_states[-2] if skip is None else out.hidden_states[-(skip + 2)])
        return torch.cat(outs, -1), pooled

    pe, pp = enc_pair(prompts, None, clip_skip)
    npe = npp = None
    if do_cfg and negative_prompt is None and force_zeros_for_empty_prompt:
        npe, npp = torch.zeros_like(pe), torch.zeros_like(pp)
    elif do_cfg:
        neg = negative_prompt or ""
        neg = b * [neg] if isinstance(neg, str) else list(neg)
        npe, npp = enc_pair(neg, pe.shape[1], None)
    L = pe.shape[1]
    pe = pe.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, L, -1)
    pp = pp.repeat(1, num_images_per_prompt).view(b * num_images_per_prompt, -1)
    if do_cfg:
        npe = npe.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, L, -1)
        npp = npp.repeat(1, num_images_per_prompt).view(b * num_images_per_prompt, -1)
    return pe, npe, pp, npp


def add_text_lora_hooks(model):
    """text-encoder LoRA as the reference gets it from diffusers' ``_modify_text_encoder`` (training_script.py:227-255; un-vendored):
    ``y = W x + up(down(x))`` on q_proj / k_proj / v_proj / out_proj of every CLIP block.  Implemented as forward hooks on the HF
    module for every Linear that carries a ``lora_layer`` attribute (comat_b200.text_encoder.install_text_lora puts them there);
    returns the hook handles."""
    handles = []
    for lyr in model.text_model.encoder.layers:
        for lin in (lyr.self_attn.q_proj, lyr.self_attn.k_proj, lyr.self_attn.v_proj, lyr.self_attn.out_proj):
            if getattr(lin, "lora_layer", None) is not None:
                handles.append(lin.register_forward_hook(
                    lambda m, inp, out: out + m.lora_layer.up(m.lora_layer.down(inp[0].to(m.lora_layer.down.weight.dtype))).to(out.dtype)))
    return handles
