"""Loading real checkpoints: ``PIPELINE.from_pretrained(model_path, ...)`` (training_utils/pipeline.py:19-39) without
diffusers - a diffusers-layout directory (what ``snapshot_download`` of runwayml/stable-diffusion-v1-5 or
stabilityai/stable-diffusion-xl-base-1.0 leaves on disk)::

    model_index.json
    unet/config.json            unet/diffusion_pytorch_model[.fp16].safetensors | .bin
    vae/config.json             vae/diffusion_pytorch_model[.fp16].safetensors  | .bin
    text_encoder[_2]/           (transformers CLIPTextModel[WithProjection].from_pretrained, local files only)
    tokenizer[_2]/              (transformers CLIPTokenizer.from_pretrained, local files only)

is read into the forward-less containers of ``comat_b200.containers`` (state-dict keys identical to diffusers', so the load is
``strict``), which the executors then pack.  Nothing is downloaded: ``model_path`` must exist locally.

Conventions of the diffusers configs handled here (un-vendored; published config schema):
* ``attention_head_dim`` is the number of heads when ``num_attention_heads`` is absent / null (a historical misnomer);
* ``down_block_types`` / ``up_block_types`` entries containing ``CrossAttn`` carry transformer blocks;
* VAE attention weights saved before the attention refactor are named ``query / key / value / proj_attn`` and are renamed to
  ``to_q / to_k / to_v / to_out.0``; only ``decoder.*`` and ``post_quant_conv.*`` are needed (the encoder is never run on this path).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch

from . import containers as Cn

_OLD_VAE_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


def _read_state(folder: str, stem: str = "diffusion_pytorch_model", variant: Optional[str] = None) -> Dict[str, torch.Tensor]:
    names = ([f"{stem}.{variant}.safetensors", f"{stem}.{variant}.bin"] if variant else []) + [f"{stem}.safetensors", f"{stem}.bin"]
    for n in names:
        path = os.path.join(folder, n)
        if os.path.exists(path):
            if n.endswith(".safetensors"):
                from safetensors.torch import load_file
                return load_file(path)
            return torch.load(path, map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"no {stem}[.variant].safetensors|.bin under {folder}")


def unet_kwargs_from_config(cfg: dict) -> dict:
    """diffusers ``unet/config.json`` -> keyword arguments of ``containers.UNet2DConditionModel``."""
    heads = cfg.get("num_attention_heads") or cfg["attention_head_dim"]
    nb = len(cfg["block_out_channels"])
    tl = cfg.get("transformer_layers_per_block", 1)
    unsupported = {k: cfg[k] for k in ("class_embed_type", "encoder_hid_dim", "time_cond_proj_dim", "conv_in_kernel") if cfg.get(k) not in (None, 3)}
    if unsupported or cfg.get("mid_block_type", "UNetMidBlock2DCrossAttn") != "UNetMidBlock2DCrossAttn" or cfg.get("dual_cross_attention") or \
            cfg.get("only_cross_attention") or cfg.get("upcast_attention"):
        raise NotImplementedError(f"UNet config outside the SD1.x / SD2.x / SDXL family: {unsupported or 'attention variant'}")
    return dict(
        in_channels=cfg["in_channels"], out_channels=cfg["out_channels"], block_out_channels=tuple(cfg["block_out_channels"]),
        layers_per_block=cfg["layers_per_block"], cross_down=tuple("CrossAttn" in t for t in cfg["down_block_types"]),
        cross_up=tuple("CrossAttn" in t for t in cfg["up_block_types"]), heads=heads if isinstance(heads, int) else tuple(heads),
        cross_attention_dim=cfg["cross_attention_dim"], transformer_layers=(tl,) * nb if isinstance(tl, int) else tuple(tl),
        use_linear_projection=bool(cfg.get("use_linear_projection", False)), addition_embed_type=cfg.get("addition_embed_type"),
        addition_time_embed_dim=cfg.get("addition_time_embed_dim"),
        projection_class_embeddings_input_dim=cfg.get("projection_class_embeddings_input_dim"))


def load_unet(folder: str, device="cpu", variant: Optional[str] = None) -> Cn.UNet2DConditionModel:
    cfg = json.load(open(os.path.join(folder, "config.json")))
    with torch.device("meta"):
        unet = Cn.UNet2DConditionModel(**unet_kwargs_from_config(cfg))
    unet.load_state_dict(_read_state(folder, variant=variant), strict=True, assign=True)
    return unet.to(device).requires_grad_(False)


def load_vae(folder: str, device="cpu", variant: Optional[str] = None) -> Cn.AutoencoderKL:
    cfg = json.load(open(os.path.join(folder, "config.json")))
    with torch.device("meta"):
        vae = Cn.AutoencoderKL(block_out_channels=tuple(cfg["block_out_channels"]), scaling_factor=cfg.get("scaling_factor", 0.18215),
                               groups=cfg.get("norm_num_groups", 32))
    state = {}
    for k, v in _read_state(folder, variant=variant).items():
        if not (k.startswith("decoder.") or k.startswith("post_quant_conv.")):
            continue                                                   # encoder / quant_conv: not on this path
        parts = k.split(".")
        if "attentions" in parts and parts[-2] in _OLD_VAE_ATTN:
            parts[-2] = _OLD_VAE_ATTN[parts[-2]]
            k = ".".join(parts)
            if v.dim() == 4:                                           # very old checkpoints store these as 1x1 convs
                v = v[:, :, 0, 0]
        state[k] = v
    vae.load_state_dict(state, strict=True, assign=True)
    return vae.to(device).requires_grad_(False)


def load_text_encoder(folder: str, device="cpu", with_projection: bool = False):
    from transformers import CLIPTextModel, CLIPTextModelWithProjection
    cls = CLIPTextModelWithProjection if with_projection else CLIPTextModel
    return cls.from_pretrained(folder, local_files_only=True).to(device).eval().requires_grad_(False)


def load_tokenizer(folder: str):
    from transformers import CLIPTokenizer
    return CLIPTokenizer.from_pretrained(folder, local_files_only=True)


def from_pretrained(cls, model_path: str, revision=None, torch_type=None, vae=None, unet=None, *, dtype=torch.float16, device="cuda",
                    lora_rank: Optional[int] = None, variant: Optional[str] = None, tokenizer=None, tokenizer_2=None, **_ignored):
    """``cls`` = one of the ``comat_b200.pipelines`` classes.  ``revision`` / ``torch_type`` are accepted for signature
    compatibility (training_utils/pipeline.py:22-37 passes them; ``torch_type`` is a typo kwarg diffusers ignores).  ``vae`` /
    ``unet``: pre-built parameter containers replacing the ones in the directory (the SDXL scripts pass a fp16-fix VAE and a
    fine-tuned UNet, pipeline.py:26-36).  ``lora_rank``: install LoRA on every attention projection
    (``set_pipeline_trainable_module``, :84-115) before the executors pack the weights."""
    from .modules import EngineUNet, EngineVAE
    from .text_encoder import EngineCLIPText
    if not os.path.isdir(model_path):
        raise FileNotFoundError(f"{model_path}: from_pretrained reads a local diffusers-layout directory (no Hub access)")
    sub = lambda name: os.path.join(model_path, name)
    unet = unet if unet is not None else load_unet(sub("unet"), device, variant)
    vae = vae if vae is not None else load_vae(sub("vae"), device, variant)
    if lora_rank:
        unet.install_lora(lora_rank)
    kw = {}
    if os.path.isdir(sub("text_encoder")):
        kw["text_encoder"] = EngineCLIPText(load_text_encoder(sub("text_encoder"), device), dtype)
    if tokenizer is not None or os.path.isdir(sub("tokenizer")):
        kw["tokenizer"] = tokenizer if tokenizer is not None else load_tokenizer(sub("tokenizer"))
    if cls.is_sdxl:
        if os.path.isdir(sub("text_encoder_2")):
            kw["text_encoder_2"] = EngineCLIPText(load_text_encoder(sub("text_encoder_2"), device, with_projection=True), dtype)
        if tokenizer_2 is not None or os.path.isdir(sub("tokenizer_2")):
            kw["tokenizer_2"] = tokenizer_2 if tokenizer_2 is not None else load_tokenizer(sub("tokenizer_2"))
        index = sub("model_index.json")
        if os.path.exists(index):
            kw["force_zeros_for_empty_prompt"] = bool(json.load(open(index)).get("force_zeros_for_empty_prompt", True))
    return cls(EngineVAE(vae, dtype), EngineUNet(unet, dtype), **kw)
