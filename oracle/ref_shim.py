"""Import shim that lets the reference's OWN files run on top of the restated third-party layer.

ORACLE / TEST INFRASTRUCTURE, build-container only: it needs /root/reference, which does not exist on the
GPU box.  Used by oracle/pin_against_reference.py (golden generation) and by tests marked ``needs_reference``.

``install()`` registers stand-ins for the packages the reference imports but this image lacks
(SURVEY Appendix E): ``diffusers`` (classes from oracle/sd_modules.py behind diffusers-shaped names),
``spacy``, ``ultralytics``, ``groundingdino``, ``seg_model...`` and puts /root/reference on sys.path.
Only third-party arithmetic is restated; CoMat's pipelines/losses execute verbatim.
"""
from __future__ import annotations

import inspect
import os
import sys
import types

import torch

from . import sd_modules as sdm

REFERENCE_ROOT = os.environ.get("COMAT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "attn_utils"))


class _PipelineBase:
    """Minimal diffusers.DiffusionPipeline surface used by the reference pipelines."""

    def __init__(self, vae=None, text_encoder=None, tokenizer=None, unet=None, scheduler=None, safety_checker=None,
                 feature_extractor=None, requires_safety_checker=False, **kw):
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.unet, self.scheduler = unet, scheduler
        self.safety_checker, self.feature_extractor = safety_checker, feature_extractor
        self.vae_scale_factor = 8
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def _execution_device(self):
        return self.unet.device

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels_latents, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            latents = torch.randn(shape, generator=generator, dtype=dtype).to(device)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def prepare_extra_step_kwargs(self, generator, eta):
        kw = {}
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        if "eta" in params:
            kw["eta"] = eta
        if "generator" in params:
            kw["generator"] = generator
        return kw

    def maybe_convert_prompt(self, prompt, tokenizer):
        return prompt

    def to(self, *a, **k):
        return self


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the mount is read-only
    if "diffusers" in sys.modules and getattr(sys.modules["diffusers"], "_comat_shim", False):
        return

    class StableDiffusionPipeline(_PipelineBase):
        pass

    class StableDiffusionXLPipeline(_PipelineBase):
        def __init__(self, vae=None, text_encoder=None, text_encoder_2=None, tokenizer=None, tokenizer_2=None,
                     unet=None, scheduler=None, **kw):
            super().__init__(vae=vae, text_encoder=text_encoder, tokenizer=tokenizer, unet=unet, scheduler=scheduler, **kw)
            self.text_encoder_2, self.tokenizer_2 = text_encoder_2, tokenizer_2

        def encode_prompt(self, prompt=None, prompt_2=None, device=None, num_images_per_prompt=1, do_classifier_free_guidance=True,
                          negative_prompt=None, negative_prompt_2=None, prompt_embeds=None, negative_prompt_embeds=None,
                          pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None, lora_scale=None, clip_skip=None):
            """diffusers StableDiffusionXLPipeline.encode_prompt restricted to what the pins need: all four embeddings are
            passed in and only repeated ``num_images_per_prompt`` times (the text-encoding half lives in
            oracle/comat_ref.encode_prompt_sdxl)."""
            if prompt_embeds is None or pooled_prompt_embeds is None:
                raise NotImplementedError("shim: pass prompt_embeds and pooled_prompt_embeds")
            n = num_images_per_prompt
            rep = lambda t: None if t is None else t.repeat(1, n, *([1] * (t.dim() - 2))).view(t.shape[0] * n, *t.shape[1:])
            if not do_classifier_free_guidance:
                negative_prompt_embeds = negative_pooled_prompt_embeds = None
            return rep(prompt_embeds), rep(negative_prompt_embeds), rep(pooled_prompt_embeds), rep(negative_pooled_prompt_embeds)

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    class LoraLoaderMixin:
        pass

    class TextualInversionLoaderMixin:
        pass

    d = _mod("diffusers", _comat_shim=True, StableDiffusionPipeline=StableDiffusionPipeline,
             StableDiffusionXLPipeline=StableDiffusionXLPipeline, AutoencoderKL=sdm.AutoencoderKL,
             UNet2DConditionModel=sdm.UNet2DConditionModel, DDPMScheduler=sdm.DDPMScheduler,
             DPMSolverMultistepScheduler=_Dummy)
    d.__path__ = []
    for pkg in ("diffusers.pipelines", "diffusers.pipelines.stable_diffusion", "diffusers.pipelines.stable_diffusion_xl",
                "diffusers.models", "diffusers.schedulers", "diffusers.utils"):
        _mod(pkg).__path__ = []
    _mod("diffusers.pipelines.stable_diffusion.pipeline_stable_diffusion", rescale_noise_cfg=sdm.rescale_noise_cfg)
    _mod("diffusers.pipelines.stable_diffusion.safety_checker", StableDiffusionSafetyChecker=_Dummy)
    _mod("diffusers.pipelines.stable_diffusion_xl.pipeline_output", StableDiffusionXLPipelineOutput=_Dummy)
    _mod("diffusers.loaders", LoraLoaderMixin=LoraLoaderMixin, TextualInversionLoaderMixin=TextualInversionLoaderMixin,
         text_encoder_lora_state_dict=lambda *a, **k: {})
    sys.modules["diffusers.models"].__dict__.update(AutoencoderKL=sdm.AutoencoderKL,
                                                    UNet2DConditionModel=sdm.UNet2DConditionModel)
    _mod("diffusers.models.lora", LoRALinearLayer=sdm.LoRALinearLayer,
         adjust_lora_scale_text_encoder=lambda *a, **k: None)
    _mod("diffusers.models.attention_processor", AttnAddedKVProcessor=type("AttnAddedKVProcessor", (), {}),
         AttnAddedKVProcessor2_0=type("AttnAddedKVProcessor2_0", (), {}),
         SlicedAttnAddedKVProcessor=type("SlicedAttnAddedKVProcessor", (), {}))
    sys.modules["diffusers.schedulers"].__dict__.update(KarrasDiffusionSchedulers=_Dummy)
    sys.modules["diffusers.utils"].__dict__.update(USE_PEFT_BACKEND=False, scale_lora_layers=lambda *a, **k: None,
                                                   unscale_lora_layers=lambda *a, **k: None,
                                                   check_min_version=lambda *a, **k: None)
    _mod("diffusers.utils.import_utils", is_xformers_available=lambda: False)
    _mod("diffusers.optimization", get_scheduler=lambda *a, **k: None)

    _mod("spacy", load=lambda name: (lambda text: text))
    _mod("ultralytics", YOLO=_Dummy)
    for pkg in ("seg_model", "seg_model.gsam", "seg_model.gsam.EfficientSAM", "seg_model.gsam.EfficientSAM.FastSAM",
                "seg_model.gsam.GroundingDINO", "seg_model.gsam.GroundingDINO.groundingdino",
                "seg_model.gsam.GroundingDINO.groundingdino.util", "groundingdino", "groundingdino.datasets"):
        _mod(pkg).__path__ = []
    _mod("seg_model.gsam.EfficientSAM.FastSAM.tools", box_prompt=None)
    _mod("seg_model.gsam.GroundingDINO.groundingdino.util.inference", load_model=None, predict=None)
    _mod("groundingdino.datasets.transforms")


def import_reference(name: str):
    """import a reference module (e.g. 'attn_utils.tc_loss_utils') through the shim."""
    install()
    import importlib
    return importlib.import_module(name)
