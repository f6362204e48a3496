"""Drop-in pipelines: same class names and ``forward`` keyword surface as the reference
(TrainableSDPipeline.py:16-225 / :427-846, AttrConcenTrainableSDPipeline.py:28-279,
AttrConcenTrainableSDXLPipeline.py:20-496), running the in-step DDPM rollout on the B200 executors.

The pipelines are written from the reference's *behaviour* (gradient windows, CFG, DDPM step, attrcon split,
SDXL quirks) — not its code — and add two keyword-only conveniences the parity tests need: ``noises`` (pre-drawn DDPM
variance noise per step instead of the global RNG) and ``added_cond_kwargs`` passthrough.  Text encoding,
spaCy parsing and GSAM masks are out of scope: prompts enter as ``prompt_embeds``.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch

from . import engine as E
from .modules import EngineUNet, EngineVAE
from .scheduler import DDPMScheduler, fused_cfg_ddpm_step, rescale_noise_cfg


class TrainableSDPipeline:
    """SD1.5 differentiable sampler (TrainableSDPipeline.py:16-225)."""

    is_sdxl = False

    def __init__(self, vae: EngineVAE, unet: EngineUNet, scheduler: Optional[DDPMScheduler] = None, text_encoder=None,
                 tokenizer=None, text_encoder_2=None, tokenizer_2=None, force_zeros_for_empty_prompt: bool = True, **_ignored):
        self.vae, self.unet = vae, unet
        self.scheduler = scheduler or DDPMScheduler()
        self.text_encoder, self.tokenizer = text_encoder, tokenizer
        self.text_encoder_2, self.tokenizer_2 = text_encoder_2, tokenizer_2              # SDXL only
        self.config = SimpleNamespace(force_zeros_for_empty_prompt=force_zeros_for_empty_prompt)
        self.controller: Optional[E.AttnCapture] = None
        self.attn_dict: Dict[str, Dict[str, List[torch.Tensor]]] = {}

    @classmethod
    def from_pretrained(cls, model_path, **kw):
        """``PIPELINE.from_pretrained(model_path, revision=..., torch_type=...[, vae=..., unet=...])`` (training_utils/pipeline.py:19-39)
        over a local diffusers-layout directory - see ``comat_b200.loading``."""
        from .loading import from_pretrained
        return from_pretrained(cls, model_path, **kw)

    @property
    def _execution_device(self):
        return self.unet.device

    def to(self, *a, **k):
        return self

    # -- prompt encoding (TrainableSDPipeline.py:227-424): CLIP-L on the B200 executor (text_encoder.EngineCLIPText)
    @staticmethod
    def _tokenize(tokenizer, text, max_length, device):
        t = tokenizer(text, padding="max_length", max_length=max_length, truncation=True, return_tensors="pt")
        return t.input_ids.to(device), getattr(t, "attention_mask", None)

    @staticmethod
    def _mask_for(text_encoder, mask, device):
        cfg = getattr(text_encoder, "config", None)                      # :312-322: only when the config asks for it
        if mask is not None and getattr(cfg, "use_attention_mask", False):
            return mask.to(device)
        return None

    def _need_encoder(self, what):
        if self.text_encoder is None or self.tokenizer is None:
            raise NotImplementedError(f"{what}: this pipeline was built without text_encoder / tokenizer - pass the "
                                      "pre-computed embeddings instead (TrainableSDPipeline.py:227-424)")

    def encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance, negative_prompt=None,
                      prompt_embeds=None, negative_prompt_embeds=None, lora_scale=None, clip_skip=None):
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, (list, tuple)):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        text_mask = None
        if prompt_embeds is None:
            self._need_encoder("encode_prompt(prompt=...)")
            ids, text_mask = self._tokenize(self.tokenizer, prompt, self.tokenizer.model_max_length, device)
            mask = self._mask_for(self.text_encoder, text_mask, device)
            if clip_skip is None:
                prompt_embeds = self.text_encoder(ids, attention_mask=mask)[0]
            else:                                                                        # :329-340
                out = self.text_encoder(ids, attention_mask=mask, output_hidden_states=True)
                prompt_embeds = self.text_encoder.text_model.final_layer_norm(out[-1][-(clip_skip + 1)])
        prompt_embeds = prompt_embeds.to(device=device, dtype=torch.float32)
        b, L, D = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, L, D)
        if do_classifier_free_guidance and negative_prompt_embeds is None:               # :357-407
            if negative_prompt is None:
                uncond = [""] * batch_size
            elif prompt is not None and type(prompt) is not type(negative_prompt):
                raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got {type(negative_prompt)} != {type(prompt)}.")
            elif isinstance(negative_prompt, str):
                uncond = [negative_prompt]
            elif batch_size != len(negative_prompt):
                raise ValueError(f"`negative_prompt` has batch size {len(negative_prompt)}, but `prompt` has batch size {batch_size}.")
            else:
                uncond = list(negative_prompt)
            self._need_encoder("encode_prompt without negative_prompt_embeds")
            ids, m = self._tokenize(self.tokenizer, uncond, L, device)
            # the reference reuses the *prompt's* mask here (:390-399, `text_inputs.attention_mask`); kept when there is one
            mask = self._mask_for(self.text_encoder, text_mask if text_mask is not None else m, device)
            negative_prompt_embeds = self.text_encoder(ids, attention_mask=mask)[0]
        if do_classifier_free_guidance:
            negative_prompt_embeds = negative_prompt_embeds.to(device=device, dtype=torch.float32)
            Ln = negative_prompt_embeds.shape[1]
            negative_prompt_embeds = negative_prompt_embeds.repeat(1, num_images_per_prompt, 1).view(batch_size * num_images_per_prompt, Ln, -1)
        return prompt_embeds, negative_prompt_embeds

    def _encode_for_forward(self, prompt, device, n_per, cfg, negative_prompt, prompt_embeds, negative_prompt_embeds, kw):
        return self.encode_prompt(prompt, device, n_per, cfg, negative_prompt, prompt_embeds=prompt_embeds,
                                  negative_prompt_embeds=negative_prompt_embeds)

    def prepare_latents(self, batch_size, num_channels, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels, height // 8, width // 8)
        if latents is None:
            latents = torch.randn(shape, generator=generator, dtype=dtype, device=generator.device if generator is not None else device)
        return latents.to(device) * self.scheduler.init_noise_sigma

    # ------------------------------------------------------------------------------------------------
    def _unet(self, x, t, embeds, added):
        return self.unet(x, t, encoder_hidden_states=embeds, added_cond_kwargs=added, return_dict=False)[0]

    # True (default): the attrcon step runs ONE UNet call over the whole classifier-free-guidance batch and the attention kernel
    # exports the probabilities of the conditional half only.  Per-sample arithmetic (GroupNorm, attention) is unchanged, so
    # the result equals the reference's two half-batch calls + torch.cat (AttrConcenTrainableSDPipeline.py:239-279), without
    # their second pass over the weights and at twice the GEMM M.  False: the reference's literal two-call structure.
    merge_attrcon_call = True

    def _attrcon_forward(self, latents, t, prompt_embeds, added=None, t_host=None):
        """AttrConcenTrainableSDPipeline.py:239-279: conditional half with attention capture, unconditional half without."""
        h = latents.shape[0] // 2
        key = str(int(t) if t_host is None else t_host)
        if self.merge_attrcon_call and hasattr(self.controller, "sample_from"):
            self.unet.capture = self.controller
            self.controller.sample_from = h
            try:
                eps = self._unet(latents, t, prompt_embeds, added)
                self.attn_dict[key] = self.controller.attn_dict()[0]
            finally:
                self.unet.capture = None
                self.controller.sample_from = 0
            return eps
        split = (lambda d, s: None if d is None else {k: v[s] for k, v in d.items()})
        self.unet.capture = self.controller
        try:
            n_c = self._unet(latents[h:], t, prompt_embeds[h:], split(added, slice(h, None)))
            maps, _ = self.controller.attn_dict()
            self.attn_dict[key] = maps
        finally:
            self.unet.capture = None
        n_u = self._unet(latents[:h], t, prompt_embeds[:h], split(added, slice(0, h)))
        return torch.cat([n_u, n_c], 0)

    def forward(self, prompt=None, height: int = 512, width: int = 512, training_timesteps: Sequence[int] = (),
                early_exit: bool = False, detach_gradient: bool = True, train_text_encoder: bool = False,
                double_laststep: bool = False, bp_on_trained: bool = False, fast_training: bool = False,
                num_inference_steps: int = 50, guidance_scale: float = 7.5, negative_prompt=None,
                num_images_per_prompt: int = 1, eta: float = 0.0, generator=None, latents=None, prompt_embeds=None,
                negative_prompt_embeds=None, output_type: str = "image", callback=None, callback_steps: int = 1,
                cross_attention_kwargs=None, guidance_rescale: float = 0.0, return_latents: bool = False,
                return_timestamped_latents: bool = False, return_early: bool = False, D_timesteps=None, batch=None,
                attrcon_train_steps=None, *, noises=None, added_cond_kwargs=None, **sdxl_kwargs):
        if double_laststep or fast_training or return_timestamped_latents:
            raise NotImplementedError("double_laststep / fast_training / return_timestamped_latents are dead branches under "
                                      "the trainer's constants (training_script.py:558-567; SURVEY A.2)")
        batch_size = len(prompt) if isinstance(prompt, (list, tuple)) else (1 if isinstance(prompt, str) else prompt_embeds.shape[0])
        device = self._execution_device
        cfg = guidance_scale > 1.0
        T = list(training_timesteps)
        prev = torch.is_grad_enabled()
        try:
            torch.set_grad_enabled(bool(train_text_encoder))                            # :72 / :728
            prompt_embeds, negative_prompt_embeds = self._encode_for_forward(
                prompt, device, num_images_per_prompt, cfg, negative_prompt, prompt_embeds, negative_prompt_embeds, sdxl_kwargs)
            embeds = torch.cat([negative_prompt_embeds, prompt_embeds]) if cfg else prompt_embeds
            torch.set_grad_enabled(False)
            added = self._added_cond(batch_size * num_images_per_prompt, height, width, cfg, added_cond_kwargs, sdxl_kwargs)
            self.scheduler.set_timesteps(num_inference_steps, device=device)
            timesteps = self.scheduler.timesteps
            latents = self.prepare_latents(batch_size * num_images_per_prompt, 4, height, width, torch.float32, device,
                                           generator, latents)
            attr = T if attrcon_train_steps is None else list(attrcon_train_steps)
            use_attr = self.controller is not None and attrcon_train_steps is not None
            ts_host = self.scheduler._ts_host          # host copy: no device sync per sampler step
            for i, t in enumerate(timesteps):
                torch.set_grad_enabled(len(T) == 0 or i > min(T))                       # TrainableSDPipeline.py:133
                x_in = torch.cat([latents] * 2) if cfg else latents
                torch.set_grad_enabled(i in T)                                          # :138
                detach = detach_gradient and not (i in T and bp_on_trained)             # :140-145
                if self.is_sdxl:
                    detach = True                                                       # :809
                x_in = x_in.detach() if detach else x_in
                if i in T and (bp_on_trained or self.is_sdxl) and use_attr and i in attr:   # AttrConcen...:159-167 / SDXL :404-412
                    eps = self._attrcon_forward(x_in, t, embeds, added, t_host=ts_host[i])
                else:
                    eps = self._unet(x_in, t, embeds, added)
                fused = eps.is_cuda and guidance_rescale == 0.0 and not (early_exit and len(T) > 0 and i == max(T))
                if fused:
                    # one launch: guidance combine + DDPM step (:155-167); the reference's grad window for both is
                    # `i >= min(T)` except that the combine of step min(T) happens under `i in T` — identical here
                    torch.set_grad_enabled(len(T) == 0 or i >= min(T))
                    latents = fused_cfg_ddpm_step(self.scheduler, eps, ts_host[i], latents, guidance_scale, cfg,
                                                  noise=None if noises is None else noises[i], generator=generator)
                    if callback is not None and i % callback_steps == 0:
                        callback(i, t, latents)
                    continue
                if cfg:
                    e_u, e_c = eps.chunk(2)
                    eps = e_u + guidance_scale * (e_c - e_u)                            # :155-157
                    if guidance_rescale > 0.0:
                        eps = rescale_noise_cfg(eps, e_c, guidance_rescale)             # :159-161
                torch.set_grad_enabled(len(T) == 0 or i >= min(T))                      # :163
                out = self.scheduler.step(eps, ts_host[i], latents, generator=generator,
                                          variance_noise=None if noises is None else noises[i])
                latents = out.prev_sample
                if callback is not None and i % callback_steps == 0:
                    callback(i, t, latents)
                if len(T) > 0 and i == max(T) and early_exit:
                    latents = out.pred_original_sample
                    break
            torch.set_grad_enabled(True)
            if output_type == "latent":
                return latents
            image = self.vae.decode(latents / self.vae.config.scaling_factor, return_dict=False)[0]
            if self.is_sdxl and return_latents:
                return image, latents                                                   # un-rescaled: :838-840 (quirk kept)
            image = image / 2 + 0.5                                                     # :223 (unclamped)
            return (image, latents) if return_latents else image
        finally:
            torch.set_grad_enabled(prev)

    def _added_cond(self, n, height, width, cfg, added_cond_kwargs, sdxl_kwargs):
        return added_cond_kwargs

    @torch.no_grad()
    def __call__(self, prompt=None, height: int = 512, width: int = 512, num_inference_steps: int = 50, guidance_scale: float = 7.5,
                 negative_prompt=None, num_images_per_prompt: int = 1, eta: float = 0.0, generator=None, latents=None,
                 prompt_embeds=None, negative_prompt_embeds=None, output_type: str = "pil", return_dict: bool = True,
                 guidance_rescale: float = 0.0, noises=None, **sdxl_kwargs):
        """Plain sampling - the inherited diffusers ``__call__`` as the reference uses it: GAN ground-truth latents
        (tools/gan_gt_generate.py:171-180, 50 steps, cfg 7.5, ``output_type='latent'``) and validation images
        (training_script.py:456-489).  Same rollout as ``forward`` with no training timesteps; every UNet call is a no-grad
        forward (CUDA-graphed when ``unet.use_graphs``).  ``.images``: latents, or the decoded image clamped to [0, 1] as a
        tensor (``'pt'``), HWC numpy (``'np'``) or PIL images (``'pil'``)."""
        if output_type not in ("latent", "pt", "np", "pil"):
            raise ValueError(f"output_type {output_type!r}")
        out = self.forward(prompt=prompt, height=height, width=width, training_timesteps=(), num_inference_steps=num_inference_steps,
                           guidance_scale=guidance_scale, negative_prompt=negative_prompt, num_images_per_prompt=num_images_per_prompt,
                           eta=eta, generator=generator, latents=latents, prompt_embeds=prompt_embeds,
                           negative_prompt_embeds=negative_prompt_embeds, output_type="latent" if output_type == "latent" else "image",
                           guidance_rescale=guidance_rescale, noises=noises, **sdxl_kwargs)
        if output_type != "latent":
            out = out.clamp(0, 1)
            if output_type in ("np", "pil"):
                out = out.permute(0, 2, 3, 1).float().cpu().numpy()
            if output_type == "pil":
                from PIL import Image
                out = [Image.fromarray((im * 255).round().astype("uint8")) for im in out]
        return SimpleNamespace(images=out) if return_dict else (out,)


class TrainableSDXLPipeline(TrainableSDPipeline):
    """SDXL twin (TrainableSDPipeline.py:427-846): UNet input always detached, pooled-text + time-id conditioning."""

    is_sdxl = True

    def _encode_pair(self, texts, texts_2, max_length, device, clip_skip):
        """both encoders over one prompt list each: penultimate hidden states concatenated on the feature axis, pooled vector =
        encoder 2's projected EOS token (diffusers StableDiffusionXLPipeline.encode_prompt, un-vendored)."""
        if self.text_encoder_2 is None or self.tokenizer_2 is None:
            raise NotImplementedError("SDXL encode_prompt: this pipeline was built without text_encoder_2 / tokenizer_2 - "
                                      "pass prompt_embeds and pooled_prompt_embeds instead")
        pairs = [(texts_2, self.tokenizer_2, self.text_encoder_2)]
        if self.tokenizer is not None and self.text_encoder is not None:
            pairs.insert(0, (texts, self.tokenizer, self.text_encoder))
        embeds, pooled = [], None
        for txt, tok, enc in pairs:
            ids, _ = self._tokenize(tok, txt, max_length or tok.model_max_length, device)
            out = enc(ids, output_hidden_states=True)
            pooled = out[0]                                  # only the last (projection) encoder's survives
            embeds.append(out.hidden_states[-2] if clip_skip is None else out.hidden_states[-(clip_skip + 2)])
        return torch.cat(embeds, -1), pooled

    def encode_prompt(self, prompt, prompt_2=None, device=None, num_images_per_prompt: int = 1,
                      do_classifier_free_guidance: bool = True, negative_prompt=None, negative_prompt_2=None,
                      prompt_embeds=None, negative_prompt_embeds=None, pooled_prompt_embeds=None,
                      negative_pooled_prompt_embeds=None, lora_scale=None, clip_skip=None):
        """-> (prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds), the 4-tuple
        training_script.py:521 unpacks."""
        device = device or self._execution_device
        prompt = [prompt] if isinstance(prompt, str) else prompt
        batch_size = len(prompt) if prompt is not None else prompt_embeds.shape[0]
        if prompt_embeds is None:
            prompt_2 = prompt_2 or prompt
            prompt_2 = [prompt_2] if isinstance(prompt_2, str) else prompt_2
            prompt_embeds, pooled_prompt_embeds = self._encode_pair(list(prompt), list(prompt_2), None, device, clip_skip)
        zero_out = negative_prompt is None and self.config.force_zeros_for_empty_prompt
        if do_classifier_free_guidance and negative_prompt_embeds is None and zero_out:
            negative_prompt_embeds = torch.zeros_like(prompt_embeds)
            negative_pooled_prompt_embeds = None if pooled_prompt_embeds is None else torch.zeros_like(pooled_prompt_embeds)
        elif do_classifier_free_guidance and negative_prompt_embeds is None:
            negative_prompt = negative_prompt or ""
            negative_prompt_2 = negative_prompt_2 or negative_prompt
            neg = batch_size * [negative_prompt] if isinstance(negative_prompt, str) else list(negative_prompt)
            neg_2 = batch_size * [negative_prompt_2] if isinstance(negative_prompt_2, str) else list(negative_prompt_2)
            if prompt is not None and batch_size != len(neg):
                raise ValueError(f"`negative_prompt` has batch size {len(neg)}, but `prompt` has batch size {batch_size}.")
            negative_prompt_embeds, negative_pooled_prompt_embeds = self._encode_pair(neg, neg_2, prompt_embeds.shape[1], device, None)
        prompt_embeds = prompt_embeds.to(device=device, dtype=torch.float32)
        b, L, D = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, L, D)
        if do_classifier_free_guidance:
            negative_prompt_embeds = negative_prompt_embeds.to(device=device, dtype=torch.float32)
            negative_prompt_embeds = negative_prompt_embeds.repeat(1, num_images_per_prompt, 1).view(batch_size * num_images_per_prompt, L, -1)
        if pooled_prompt_embeds is not None:
            pooled_prompt_embeds = pooled_prompt_embeds.to(device=device, dtype=torch.float32).repeat(1, num_images_per_prompt).view(
                b * num_images_per_prompt, -1)
        if do_classifier_free_guidance and negative_pooled_prompt_embeds is not None:
            negative_pooled_prompt_embeds = negative_pooled_prompt_embeds.to(device=device, dtype=torch.float32).repeat(
                1, num_images_per_prompt).view(b * num_images_per_prompt, -1)
        return prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds

    def _encode_for_forward(self, prompt, device, n_per, cfg, negative_prompt, prompt_embeds, negative_prompt_embeds, kw):
        pe, npe, pp, npp = self.encode_prompt(
            prompt, kw.get("prompt_2"), device, n_per, cfg, negative_prompt, kw.get("negative_prompt_2"), prompt_embeds,
            negative_prompt_embeds, kw.get("pooled_prompt_embeds"), kw.get("negative_pooled_prompt_embeds"))
        kw["pooled_prompt_embeds"], kw["negative_pooled_prompt_embeds"] = pp, npp
        return pe, npe

    def _get_add_time_ids(self, original_size, crops_coords_top_left, target_size, dtype=torch.float32):
        add_time_ids = list(original_size + crops_coords_top_left + target_size)        # :428-449
        want = self.unet.ref.add_embedding.linear_1.in_features
        have = self.unet.config.addition_time_embed_dim * len(add_time_ids) + self._pooled_dim
        if want != have:
            raise ValueError(f"Model expects an added time embedding vector of length {want}, but a vector of {have} was created.")
        return torch.tensor([add_time_ids], dtype=dtype)

    def _added_cond(self, n, height, width, cfg, added_cond_kwargs, kw):
        if added_cond_kwargs is not None:
            return added_cond_kwargs
        pooled, neg_pooled = kw.get("pooled_prompt_embeds"), kw.get("negative_pooled_prompt_embeds")
        if pooled is None:
            raise NotImplementedError("pass pooled_prompt_embeds with prompt_embeds, or build the pipeline with text_encoder_2 / tokenizer_2")
        self._pooled_dim = pooled.shape[-1]
        osz = kw.get("original_size") or (height, width)
        tsz = kw.get("target_size") or (height, width)
        ids = self._get_add_time_ids(tuple(osz), tuple(kw.get("crops_coords_top_left", (0, 0))), tuple(tsz)).to(pooled.device)
        ids = ids.repeat(n, 1)
        if cfg:
            pooled = torch.cat([neg_pooled, pooled], 0)
            ids = torch.cat([ids, ids], 0)
        return {"text_embeds": pooled.float(), "time_ids": ids.float()}


class AttrConcenTrainableSDPipeline(TrainableSDPipeline):
    """Adds cross-attention capture on ``attrcon_train_steps`` (AttrConcenTrainableSDPipeline.py:28-279).  The spaCy
    noun/attribute alignment (:281-338) is out of scope: token-index lists are inputs to the loss."""


class AttrConcenTrainableSDXLPipeline(TrainableSDXLPipeline):
    """SDXL twin (AttrConcenTrainableSDXLPipeline.py:20-496)."""


def register_attention_control(pipeline_or_unet, controller: E.AttnCapture):
    """Product-side equivalent of attn_utils/tc_attn_utils.py:96-196: instead of monkey-patching every Attention.forward,
    the cross-attention kernel exports P for the controller's places.  Returns the hooked attention-layer count."""
    unet = getattr(pipeline_or_unet, "unet", pipeline_or_unet)
    n = 0
    for blocks in (unet.engine.down, [(unet.engine.mid[0], unet.engine.mid[1], None)], unet.engine.up):
        for _, attns, _ in blocks:
            for t in attns or []:
                n += 2 * len(t.blocks)
    controller.num_att_layers = n
    if hasattr(pipeline_or_unet, "unet"):
        pipeline_or_unet.controller = controller
    return n


AttentionStore = E.AttnCapture


def get_cross_attn_map_from_unet(attention_store: E.AttnCapture, is_training_sd21=False, reses=(64, 32, 16, 8),
                                 poses=("down", "mid", "up")):
    """tc_attn_utils.py:198-216."""
    if is_training_sd21:
        reses = [int(1.5 * r) for r in reses]
    return attention_store.attn_dict(reses, poses)[0]
