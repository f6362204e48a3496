"""One CoMat optimiser step — the body of ``Trainer.train``'s hot loop (training_script.py:556-694) on the B200 path.

Behavioural mirror (same step selection, loss assembly, clip / AdamW hyper-parameters, D update), with the host
synchronisations removed: no ``.item()`` / ``accelerator.gather`` per scalar — logs are returned as
device tensors — and the DDP bucketed all-reduce (C2/C3) replaced by ONE NCCL all-reduce of the flat LoRA-gradient
buffer per optimiser, folded with 1/world into the fused clip+AdamW kernel.
"""
from __future__ import annotations

import gc
import random
from typing import Dict, Optional

import torch

from . import attn_loss
from .optim import FlatAdamW


class CoMatTrainer:
    def __init__(self, args, pipeline, caption_model, D=None, rng: Optional[random.Random] = None, process_group=None,
                 manual_gc_interval: int = 25, attr_provider=None):
        """``manual_gc_interval`` > 0: the trainer takes over Python's cyclic GC - automatic collection is switched off while it
        steps and a full collection runs every that many steps, right after a step has been enqueued (so it overlaps the GPU's
        backlog).  A step builds ~10^5 short-lived Python objects (tape nodes, ctypes structs); the automatic generation-2
        collections they trigger pause the host for 30-80 ms at arbitrary points and starve the GPU (measured: 714 -> 644
        ms/step, profiles/r01_step_kernels_v7_two_steps.md).  Tensors are released by reference counting (the tape is cleared
        explicitly), so device memory does not depend on the collector.  0 leaves the interpreter's GC settings alone."""
        self.manual_gc_interval = int(manual_gc_interval)
        self.args, self.pipeline, self.caption_model, self.D = args, pipeline, caption_model, D
        # ``attr_provider(prompts, images in [0,1]) -> (words, masks)``: the Grounded-SAM + spaCy seam (training_script.py:627-637
        # segments the image generated IN this step); used when the batch does not carry ``words`` / ``masks`` itself
        self.attr_provider = attr_provider
        if getattr(args, "tune_text_encoder", False):
            raise NotImplementedError("--tune_text_encoder: full text-encoder weight gradients are not implemented (LoRA only)")
        self.null_embed = self.pooled_null_embed = self.gan_null_embed = None
        # --train_text_encoder_lora (training_script.py:227-255): the text LoRA joins G_parameters - one AdamW, one joint gradient-norm
        # clip (:661 clips self.G_parameters, which includes them).  Host logic checked on emulated ops; not yet run on the B200.
        self.train_text = bool(getattr(args, "train_text_encoder_lora", False))
        self.text_parameters = []
        if self.train_text:
            enc = getattr(pipeline, "text_encoder", None)
            self.text_parameters = list(enc.lora_parameters()) if hasattr(enc, "lora_parameters") else []
            if pipeline.is_sdxl or not self.text_parameters:
                raise NotImplementedError("--train_text_encoder_lora needs an SD1.5 pipeline whose text encoder was built after "
                                          "comat_b200.text_encoder.install_text_lora(model, rank)")
            if args.textenc_lora_lr is not None and args.textenc_lora_lr != args.learning_rate:
                raise NotImplementedError("--textenc_lora_lr different from --learning_rate needs a per-group rate in the fused AdamW kernel")
        self.rng = rng or random.Random(args.seed)
        self.G_parameters = list(pipeline.unet.lora_parameters()) + self.text_parameters   # training_utils/pipeline.py:123-143, :172-181
        self.optimizer = FlatAdamW(self.G_parameters, lr=args.learning_rate, betas=(args.adam_beta1, args.adam_beta2),
                                   weight_decay=args.adam_weight_decay, eps=args.adam_epsilon,
                                   max_grad_norm=args.max_grad_norm, process_group=process_group)
        self.D_optimizer = None
        if D is not None:
            self.D_parameters = D.get_trainable_parameters()
            self.D_optimizer = FlatAdamW(self.D_parameters, lr=args.learning_rate_D, betas=(args.adam_beta1_D, args.adam_beta2_D),
                                         weight_decay=args.adam_weight_decay, eps=args.adam_epsilon,
                                         max_grad_norm=args.max_grad_norm_D, process_group=process_group)
            D.unet.refresh_lora()
        pipeline.unet.refresh_lora()
        # both optimisers own flat gradient buffers that are zeroed before each backward: let the UNet executors accumulate
        # the LoRA weight gradients of the K back-propagated sampler steps into them directly
        for u in [pipeline.unet] + ([D.unet] if D is not None else []):
            if hasattr(u, "direct_lora_grads"):
                u.direct_lora_grads = True
        self.attrcon = "attrcon" in args.pretrain_model_name
        self.train_layer_ls = getattr(args, "train_layer_ls", None) or (
            ["mid_8", "up_16", "up_32", "up_64"] if "sdxl" not in args.pretrain_model_name else ["mid_16", "up_16", "up_32"])
        self.global_step = 0
        # The tail of each update (projection of the accumulated LoRA product gradients: ~640 tiny GEMMs per UNet; gradient
        # all-reduce; clip + AdamW; refresh of the 16-bit operand images) touches only persistent buffers and is independent
        # of what the main stream does next: the generator's update runs beside the discriminator step, the discriminator's
        # beside the next step's rollout.  It is issued on a side stream and joined (event wait) right before the updated
        # network is used again.  No tensor is allocated on the side stream (caching-allocator safety).
        self.overlap_updates = bool(self.G_parameters and self.G_parameters[0].is_cuda)
        self._side_stream = None
        self._ev_G = self._ev_D = None
        # fp16 loss scale of the explicit backward passes, managed as GradScaler does under ``--mixed_precision fp16``
        # (training_script.py:111 -> accelerate): the fused AdamW skips an update whose gradient holds an inf / NaN (device-side,
        # csrc/optim.cu); the skip counter is read back one step late without a sync, the scale is halved after an overflow and
        # doubled again after ``scale_growth_interval`` clean steps
        self.scale_growth_interval, self._clean_steps = 2000, 0
        self.scale_min, self.scale_max = 1.0, 65536.0
        self._scaled = None

    # ---- dynamic loss scale
    def _scaled_modules(self):
        """executors that back-propagate in fp16 under a loss scale (bf16 executors carry grad_scale == 1 and are left alone)."""
        if self._scaled is None:
            blip = getattr(getattr(self.caption_model, "blip_model", None), "model", None)
            mods = [self.pipeline.unet, getattr(self.pipeline, "vae", None), getattr(self.pipeline, "text_encoder", None), blip,
                    self.D.unet if self.D is not None else None]
            self._scaled = [m for m in mods if isinstance(getattr(m, "grad_scale", None), float) and m.grad_scale != 1.0]
        return self._scaled

    def loss_scale(self):
        mods = self._scaled_modules()
        return mods[0].grad_scale if mods else 1.0

    def _update_loss_scale(self):
        overflow, polled = False, False
        for opt in (self.optimizer, self.D_optimizer):
            r = opt.poll_overflow() if opt is not None and hasattr(opt, "poll_overflow") else None
            if r is not None:
                polled = True
                overflow |= r[0] > 0
        if not polled:
            return
        factor = 1.0
        if overflow:
            factor, self._clean_steps = 0.5, 0
        else:
            self._clean_steps += 1
            if self._clean_steps >= self.scale_growth_interval:
                factor, self._clean_steps = 2.0, 0
        if factor != 1.0:
            for m in self._scaled_modules():
                m.grad_scale = float(min(self.scale_max, max(self.scale_min, m.grad_scale * factor)))

    # ---- side-stream plumbing
    def _run_update(self, fn, which):
        if not self.overlap_updates:
            fn()
            return
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        self._side_stream.wait_stream(main)                  # after the backward pass that produced the gradients
        with torch.cuda.stream(self._side_stream):
            fn()
            ev = torch.cuda.Event()
            ev.record(self._side_stream)
        setattr(self, which, ev)

    def _join(self, which):
        ev = getattr(self, which)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            setattr(self, which, None)

    def sync(self):
        """make the current stream wait for both outstanding updates (end of training / end of a timed region)."""
        self._join("_ev_G")
        self._join("_ev_D")

    def _g_update(self):
        if hasattr(self.pipeline.unet, "finalize_lora_grads"):
            self.pipeline.unet.finalize_lora_grads()                                 # accumulated dy^T x -> d up, d down (once per step)
        handle = self.optimizer.all_reduce()                                         # the only data-path collective
        self.optimizer.step(handle)                                                  # :661-663 (clip folded in)
        self._refresh(self.pipeline.unet)
        if self.train_text:
            self.pipeline.text_encoder.refresh_lora()

    @staticmethod
    def _refresh(unet):
        try:
            unet.refresh_lora(rebuild_folded=True)      # EngineUNet: folded weights rebuilt here, off the critical path
        except TypeError:
            unet.refresh_lora()

    def _d_update(self):
        if hasattr(self.D.unet, "finalize_lora_grads"):
            self.D.unet.finalize_lora_grads()
        h = self.D_optimizer.all_reduce()
        self.D_optimizer.step(h)
        self._refresh(self.D.unet)

    # -- training_script.py:513-525: the '' embeddings, encoded ONCE before the loop (pipelines built with text encoders)
    @torch.no_grad()
    def prepare_null_embeds(self):
        a, pipe = self.args, self.pipeline
        dev = pipe._execution_device
        if not a.do_classifier_free_guidance:
            return
        if a.gan_loss and self.D is not None and getattr(self.D, "D_sd_pipeline", None) is not None:
            self.gan_null_embed, _ = self.D.encode_prompt("", dev, a.train_batch_size, do_classifier_free_guidance=False)
        if pipe.is_sdxl:
            self.null_embed, _, self.pooled_null_embed, _ = pipe.encode_prompt(
                "", device=dev, num_images_per_prompt=a.train_batch_size, do_classifier_free_guidance=False)
        else:
            self.null_embed = pipe.encode_prompt("", dev, a.train_batch_size, do_classifier_free_guidance=False)[0]

    def _null(self, batch, key, attr):
        v = batch.get(key)
        if v is None and self.train_text and key == "null_embeds":
            # :569-573: with a trained text encoder the '' embedding is re-encoded every step, inside the autograd graph
            return self.pipeline.encode_prompt("", self.pipeline._execution_device, self.args.train_batch_size,
                                               do_classifier_free_guidance=False)[0]
        if v is None:
            if getattr(self, attr) is None:
                self.prepare_null_embeds()
            v = getattr(self, attr)
        return v

    # -- training_script.py:563-566, :589-590
    def select_steps(self):
        a = self.args
        interval = a.total_step // a.K
        max_start = a.total_step - interval * (a.K - 1) - 1
        start = self.rng.randint(0, max_start)
        steps = list(range(start, a.total_step, interval))
        attr = self.rng.choices(steps, k=min(a.attrcon_train_steps, len(steps))) if self.attrcon else None
        return steps, attr

    def g_losses(self, batch: Dict) -> Dict[str, torch.Tensor]:
        """forward part of the G step: rollout -> reward / GAN / attention losses -> total ``loss``."""
        a, pipe = self.args, self.pipeline
        steps, attr = batch.get("training_steps"), batch.get("attrcon_steps")
        if steps is None:
            steps, attr = self.select_steps()
        kwargs = dict(prompt=batch.get("text"), prompt_embeds=batch.get("prompt_embeds"), height=a.resolution, width=a.resolution,
                      training_timesteps=steps, detach_gradient=True, train_text_encoder=self.train_text,
                      num_inference_steps=a.total_step, guidance_scale=a.cfg_scale, guidance_rescale=a.cfg_rescale,
                      negative_prompt_embeds=self._null(batch, "null_embeds", "null_embed") if a.do_classifier_free_guidance else None,
                      early_exit=False, return_latents=bool(a.gan_loss), latents=batch.get("init_latents"),
                      noises=batch.get("noises"))
        if self.attrcon:
            kwargs["attrcon_train_steps"] = attr
        if pipe.is_sdxl:
            kwargs.update(pooled_prompt_embeds=batch.get("pooled_prompt_embeds"),
                          negative_pooled_prompt_embeds=self._null(batch, "pooled_null_embeds", "pooled_null_embed")
                          if a.do_classifier_free_guidance else None)
            out = pipe.forward(**kwargs)
        else:
            out = pipe.forward(bp_on_trained=True, double_laststep=False, fast_training=False, **kwargs)
        image, training_latents = out if a.gan_loss else (out, None)
        off = a.resolution // 224                                                   # :606-611
        ox, oy = batch.get("crop") or (self.rng.randint(0, off), self.rng.randint(0, off))
        size = a.resolution - off
        rewards = self.caption_model(image[:, :, ox:ox + size, oy:oy + size], batch.get("text"), batch=batch.get("blip"))
        logs = {k: v.detach() for k, v in rewards.items()}
        loss = -rewards["total"].mean()                                              # :618
        if a.gan_loss:
            self._join("_ev_D")                                                      # the discriminator's previous update is in
            g = self.D.D_sd_pipeline_forward(training_latents, side="G", negative_prompt_embeds=self._null(batch, "gan_null_embeds", "gan_null_embed"),
                                             num_inference_steps=a.total_step)
            loss = loss + a.gan_loss_weight * g                                      # :620-625
            logs["G_loss"] = g.detach()
        if self.attrcon and pipe.attn_dict:
            words, masks = batch.get("words"), batch.get("masks")
            if words is None or masks is None:
                if self.attr_provider is None:
                    raise KeyError("attrcon step without batch['words'] / batch['masks'] and without an attr_provider")
                with torch.no_grad():
                    words, masks = self.attr_provider(batch["text"], image.detach().clamp(0, 1))          # :631-632
            tok, pix = attn_loss.get_mask_loss(pipe.attn_dict, words, masks, self.train_layer_ls)
            loss = loss + a.mask_token_loss_weight * tok + a.mask_pixel_loss_weight * pix     # :639-640
            logs["token_loss"], logs["pixel_loss"] = tok.detach(), pix.detach()
            pipe.attn_dict = {}                                                      # :642
        if image.requires_grad:
            norm_holder = {}

            def record_grad(grad):                                                   # :644-651, without the host sync
                n = grad.norm(2)
                norm_holder["reward_norm"] = n
                return grad / (n / 1e4) if a.norm_grad else grad
            image.register_hook(record_grad)
            logs["_norm_holder"] = norm_holder
        logs["loss"] = loss
        logs["_image"], logs["_latents"] = image, training_latents
        return logs

    def _gc_before_step(self):
        if self.manual_gc_interval > 0 and gc.isenabled():
            gc.collect()
            gc.freeze()          # weights, executors, captured graphs: long-lived, keep them out of every later scan
            gc.disable()

    def close(self):
        """end of training: join the outstanding updates and give the process its garbage collector back (``_gc_before_step`` froze
        the long-lived objects and disabled the automatic collector for the duration of the loop)."""
        self.sync()
        if self.manual_gc_interval > 0 and not gc.isenabled():
            gc.unfreeze()
            gc.enable()

    def _gc_after_step(self):
        if self.manual_gc_interval > 0 and self.global_step % self.manual_gc_interval == 0:
            gc.collect()

    def train_step(self, batch: Dict, accum_steps: int = 1, first: bool = True, last: bool = True) -> Dict[str, torch.Tensor]:
        """one micro-batch.  ``accum_steps`` > 1 = ``accelerator.accumulate`` (training_script.py:556, :679): the gradients of
        consecutive calls add up from the one flagged ``first`` (buffers zeroed) to the one flagged ``last`` (optimisers step),
        each loss scaled by 1/accum_steps as ``accelerator.backward`` does.  Identical to the reference at ``accum_steps = 1`` (both
        shipped scripts).  For A > 1 this is the *intended* accumulation: the reference calls ``optimizer.zero_grad()`` before
        ``backward`` on every micro-step (:658-659), and accelerate's wrapped ``zero_grad`` - un-vendored; a no-op only while
        ``sync_gradients`` is False - therefore clears the accumulated gradients on the very micro-step that steps, so the
        reference applies loss/A of the LAST micro-batch alone."""
        a = self.args
        self._gc_before_step()
        self._join("_ev_G")                                                          # the generator's previous update is in
        if first:
            self._update_loss_scale()
            for u in (self.pipeline.unet, getattr(self.D, "unet", None)):
                if hasattr(u, "new_step"):
                    u.new_step()                                                     # captured taped-call graphs are all free again
        logs = self.g_losses(batch)
        loss = logs["loss"]
        if first:
            self.optimizer.zero_grad()                                               # :658
        (loss if accum_steps == 1 else loss / accum_steps).backward()                # :659
        if last:
            self._run_update(self._g_update, "_ev_G")                                # :661-664, beside the discriminator step
        out = {k: v for k, v in logs.items() if not k.startswith("_")}
        out["step_loss"] = loss.detach()
        out.update(logs.get("_norm_holder", {}))
        if a.gan_loss:                                                               # :679-694
            d_loss = self.D.D_sd_pipeline_forward(logs["_latents"].detach(), side="D", negative_prompt_embeds=self._null(batch, "gan_null_embeds", "gan_null_embed"),
                                                  num_inference_steps=a.total_step, batch={"latents": batch["real_latents"]})
            if first:
                self.D_optimizer.zero_grad()
            (d_loss if accum_steps == 1 else d_loss / accum_steps).backward()
            if last:
                self._run_update(self._d_update, "_ev_D")                            # :689-694, beside the next step's rollout
            out["D_loss"] = d_loss.detach()
        if last:
            self.global_step += 1
            self._gc_after_step()
        return out
