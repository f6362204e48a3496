"""Fidelity-preservation GAN: the discriminator is a second (LoRA'd) SD1.5 UNet + a per-pixel Linear(4,1) head.

Mirrors ``D_sd`` (training_utils/gan_sdxl.py:6-155) and ``load_discriminator`` (training_utils/gan_sd_model.py:8-14):
``D_sd_pipeline_forward(training_latents, side='G'|'D', **kwargs)`` with the same kwargs the trainer passes
(``negative_prompt_embeds``, ``num_inference_steps``, ``batch``).  The UNet runs on the B200 executor.
"""
from __future__ import annotations

import torch

from . import ops
from .modules import EngineUNet
from .scheduler import DDPMScheduler


class D_sd(torch.nn.Module):
    def __init__(self, unet: EngineUNet, mlp: torch.nn.Module = None, scheduler: DDPMScheduler = None, pipeline=None):
        super().__init__()
        self.unet = unet
        self.D_sd_pipeline = pipeline          # optional TrainableSD(XL)Pipeline carrying the D side's text encoder(s) (gan_sdxl.py:13)
        self.mlp = mlp if mlp is not None else torch.nn.Sequential(torch.nn.Linear(4, 1))    # gan_sdxl.py:31-34 (fp32)
        self.mlp.to(device=unet.device, dtype=torch.float32)
        self.ori_scheduler = scheduler or DDPMScheduler()
        self.D_parameters = None

    def get_trainable_parameters(self):
        self.D_parameters = list(self.unet.lora_parameters()) + list(self.mlp.parameters())   # gan_sdxl.py:37-40
        return self.D_parameters

    def set_D_sd_pipeline_lora(self, requires_grad=True):
        for p in self.D_parameters or self.get_trainable_parameters():
            p.requires_grad = requires_grad

    def get_D_gt_noise(self, device, **kwargs):
        return kwargs["batch"]["latents"].to(device, dtype=torch.float32)                     # gan_sdxl.py:46-48

    def D_sd_pipeline_forward(self, training_latents, side="G", **kwargs):
        device = training_latents.device
        self.ori_scheduler.set_timesteps(kwargs["num_inference_steps"], device=device)
        t = self.ori_scheduler.timesteps[-1]                                                  # "its own step" (=1)
        null = kwargs["negative_prompt_embeds"]
        if side == "G":                                                                       # gan_sdxl.py:52-89
            self.set_D_sd_pipeline_lora(False)
            x, cond = training_latents, null
        elif side == "D":                                                                     # gan_sdxl.py:92-132
            self.set_D_sd_pipeline_lora(True)
            with torch.no_grad():
                real = self.get_D_gt_noise(device, **kwargs)
            x = torch.cat([training_latents.detach(), real])
            cond = torch.cat([null, null])
        else:
            raise ValueError(side)
        eps = self.unet(x, t, encoder_hidden_states=cond, return_dict=False)[0]
        head = self.mlp[0] if isinstance(self.mlp, torch.nn.Sequential) and len(self.mlp) == 1 else self.mlp
        if not (isinstance(head, torch.nn.Linear) and head.out_features == 1 and head.in_features == eps.shape[1] and head.bias is not None):
            raise NotImplementedError("D_sd head: the per-pixel nn.Linear(4, 1) of gan_sdxl.py:31-34 is the one built here "
                                      "(--gan_unet_lastlayer_cls swaps in a conv head, refused in load_discriminator)")
        # permute -> Linear(4, 1) -> BCEWithLogits (:84-89 / :118-132) as one fused kernel; targets: generated half 0 on the D side
        return ops.gan_head_bce(eps, head.weight, head.bias, x.shape[0] // 2 if side == "D" else 0)

    @torch.no_grad()
    def encode_prompt(self, prompt, device, batch_size, do_classifier_free_guidance=False):
        """gan_sdxl.py:134-155: the D pipeline's embedding of ``prompt`` (the trainer passes '' once, training_script.py:516)
        -> (null_embed, pooled_null_embed | None); the text encoder is released afterwards ("only called once")."""
        pipe = self.D_sd_pipeline
        if pipe is None:
            raise NotImplementedError("D_sd was built without a D_sd_pipeline (text encoder): pass gan_null_embeds in the batch")
        if pipe.is_sdxl:
            null, _, pooled, _ = pipe.encode_prompt(prompt, device=device, num_images_per_prompt=batch_size,
                                                    do_classifier_free_guidance=do_classifier_free_guidance)
        else:
            null = pipe.encode_prompt(prompt, device, batch_size, do_classifier_free_guidance=do_classifier_free_guidance)[0]
            pooled = None
        pipe.text_encoder = None               # the reference parks it on the CPU and deletes it (:151-152, training_script.py:529-530)
        return null, pooled


def load_discriminator(args, unet_or_dtype, device=None):
    """gan_sd_model.py:8-14 keeps the reference quirk: the arch string has 'gan' stripped and only 'sd_1_5' resolves
    (the default 'gan_sd_1_5' -> '_sd_1_5' resolves to nothing -> None).

    Two call forms: ``load_discriminator(args, unet)`` with a ready ``EngineUNet`` (synthetic weights, tests), and the reference's
    ``load_discriminator(args, weight_dtype, device)`` (training_script.py:290), which builds the D pipeline from the local
    diffusers directory ``args.gan_pretrain_model`` / ``args.pretrain_model`` like ``D_sd.__init__`` (gan_sdxl.py:7-35: the reference
    hard-codes the Hub id runwayml/stable-diffusion-v1-5; there is no Hub here)."""
    if isinstance(unet_or_dtype, torch.dtype):
        from .pipelines import TrainableSDPipeline
        path = getattr(args, "gan_pretrain_model", None) or args.pretrain_model
        d_pipe = TrainableSDPipeline.from_pretrained(path, dtype=unet_or_dtype, device=device or "cuda", lora_rank=args.lora_rank)
        D = load_discriminator(args, d_pipe.unet)
        if D is not None:
            d_pipe.unet = d_pipe.vae = None            # D keeps the UNet; the D pipeline only lends its text encoder (:134-155)
            D.D_sd_pipeline = d_pipe
        return D
    unet = unet_or_dtype
    if getattr(args, "gan_unet_lastlayer_cls", False):
        # gan_sdxl.py:27-30 swaps the D UNet's conv_out for a 1-channel classifier conv; only the Linear(4,1) head (:31-34,
        # what both shipped scripts train) is built here - refuse rather than silently train a different discriminator
        raise NotImplementedError("--gan_unet_lastlayer_cls: the conv_out classification head is not implemented")
    if getattr(args, "condition_discriminator", False):
        raise NotImplementedError("--condition_discriminator crashes in the reference (gan_sdxl.py:60, self.pipeline undefined)")
    name = args.gan_model_arch.replace("gan", "")
    if name == "sd_1_5":
        return D_sd(unet)
    if "sdxl" in name:
        # gan_sd_model.py:13-14 would build D_sdxl, whose constructor cannot run (gan_sdxl.py:161 calls super().__init__() without
        # the required arguments -> TypeError); both shipped scripts use the SD1.5 discriminator
        raise NotImplementedError("an SDXL discriminator (D_sdxl) is broken in the reference and not implemented here; use gansd_1_5")
    return None
