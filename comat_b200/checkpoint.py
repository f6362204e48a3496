"""Checkpoint wire format: what ``save_and_evaluate`` writes (training_script.py:382-426) and the resume block
reads back (:156-205), plus the state the reference does NOT save and an exact resume needs (optimiser moments, step counters,
RNG streams).

Layout of ``<output_dir>/checkpoint-<global_step>/``:

* ``pytorch_lora_weights.safetensors`` - the generator UNet's LoRA factors, fp32, metadata ``{"format": "pt"}``.  Tensor names
  start from the reference's own ``unet_lora_state_dict`` (training_script.py:50-66): ``unet.<module path>.lora.{down,up}.weight``
  with diffusers module paths (``down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q`` ...).  The reference then hands
  that dict to diffusers' ``LoraLoaderMixin.save_lora_weights`` (:397-401), which - in the pinned 0.22-0.25 range, un-vendored,
  restated from its published behaviour - prefixes every entry with ``unet.`` once more; its loader strips every ``unet.``
  occurrence, so both spellings load there.  ``diffusers_prefix=True`` (default) writes what that call chain writes
  (``unet.unet.<path>...``); the reader here accepts either.
* ``D_sd/pytorch_lora_weights.safetensors`` + ``D_sd/mlp.pt`` (``torch.save`` of the head's ``state_dict``, re-floated on load,
  :193-198) when the GAN loss is on (:410-426).
* ``trainer_state.pt`` - NOT in the reference: AdamW moments and step counts of both optimisers, ``global_step``, Python /
  torch (CPU + CUDA) RNG states.  Absent file => the reference's behaviour (fresh optimiser).
"""
from __future__ import annotations

import os
import random
from typing import Dict, Optional

import torch

LORA_FILE = "pytorch_lora_weights.safetensors"
STATE_FILE = "trainer_state.pt"


def _container(unet):
    """the parameter-owning diffusers-shaped module behind an ``EngineUNet`` (or the module itself)."""
    return getattr(unet, "ref", unet)


def unet_lora_state_dict(unet) -> Dict[str, torch.Tensor]:
    """training_script.py:50-66: ``unet.<module path>.lora.<down|up>.weight`` -> parameter, in module order."""
    out = {}
    for name, module in _container(unet).named_modules():
        if hasattr(module, "set_lora_layer"):
            lora = getattr(module, "lora_layer", None)
            if lora is not None:
                for matrix, p in lora.state_dict().items():
                    out[f"unet.{name}.lora.{matrix}"] = p
    return out


def text_encoder_lora_state_dict(text_encoder) -> Dict[str, torch.Tensor]:
    """diffusers' ``text_encoder_lora_state_dict`` (training_script.py:388-389; un-vendored, restated from the 0.22-0.25 sources):
    ``text_model.encoder.layers.<i>.self_attn.<q|k|v|out>_proj.lora_linear_layer.<down|up>.weight`` -> parameter."""
    model = getattr(text_encoder, "ref", text_encoder)
    out = {}
    for i, lyr in enumerate(model.text_model.encoder.layers):
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            lora = getattr(getattr(lyr.self_attn, proj), "lora_layer", None)
            if lora is not None:
                for matrix, p in lora.state_dict().items():
                    out[f"text_model.encoder.layers.{i}.self_attn.{proj}.lora_linear_layer.{matrix}"] = p
    return out


def save_lora_weights(save_directory: str, unet_lora_layers: Dict[str, torch.Tensor], diffusers_prefix: bool = True,
                      weight_name: str = LORA_FILE, text_encoder_lora_layers: Optional[Dict[str, torch.Tensor]] = None) -> str:
    """``LoraLoaderMixin.save_lora_weights(save_directory=..., unet_lora_layers=..., text_encoder_lora_layers=...)`` as called at
    training_script.py:397-401: one file, ``unet.`` / ``text_encoder.`` prefixed entries."""
    from safetensors.torch import save_file
    os.makedirs(save_directory, exist_ok=True)
    sd = {(f"unet.{k}" if diffusers_prefix else k): v.detach().to("cpu", torch.float32).contiguous()
          for k, v in unet_lora_layers.items()}
    for k, v in (text_encoder_lora_layers or {}).items():
        sd[f"text_encoder.{k}"] = v.detach().to("cpu", torch.float32).contiguous()
    path = os.path.join(save_directory, weight_name)
    save_file(sd, path, metadata={"format": "pt"})
    return path


def lora_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """``LoraLoaderMixin.lora_state_dict(file)`` (:180): tensors by bare module path, every leading ``unet.`` stripped."""
    from safetensors.torch import load_file
    if os.path.isdir(path):
        path = os.path.join(path, LORA_FILE)
    out = {}
    for k, v in load_file(path).items():
        if k.startswith("text_encoder"):
            continue                                    # read by text_lora_state_dict
        while k.startswith("unet."):
            k = k[len("unet."):]
        out[k] = v
    return out


def text_lora_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """the ``text_encoder.`` entries of a LoRA file, prefix stripped (``load_lora_into_text_encoder``'s input, :183-185)."""
    from safetensors.torch import load_file
    if os.path.isdir(path):
        path = os.path.join(path, LORA_FILE)
    return {k[len("text_encoder."):]: v for k, v in load_file(path).items() if k.startswith("text_encoder.")}


def load_lora_into_text_encoder(state: Dict[str, torch.Tensor], text_encoder) -> int:
    """copy stored text-encoder LoRA factors into the installed adapters (in place); every stored tensor must find its adapter."""
    mine = text_encoder_lora_state_dict(text_encoder)
    missing = set(state) - set(mine)
    if missing or set(mine) - set(state):
        raise KeyError(f"text-encoder LoRA mismatch: {sorted(missing)[:3]} not in the model / {sorted(set(mine) - set(state))[:3]} not in the file")
    with torch.no_grad():
        for k, t in state.items():
            if tuple(mine[k].shape) != tuple(t.shape):
                raise ValueError(f"{k}: checkpoint shape {tuple(t.shape)} vs model {tuple(mine[k].shape)}")
            mine[k].copy_(t.to(mine[k].device, mine[k].dtype))
    if hasattr(text_encoder, "refresh_lora"):
        text_encoder.refresh_lora()
    return len(state)


def load_lora_into_unet(state: Dict[str, torch.Tensor], unet) -> int:
    """``LoraLoaderMixin.load_lora_into_unet`` (:181): copy the factors into the (already installed) LoRA layers; a module that
    has none yet gets one at the rank the stored ``down`` matrix carries.  Returns the number of tensors loaded; unknown
    module paths, shape mismatches and half-loaded adapters raise."""
    from .containers import LoRALinearLayer
    root = _container(unet)
    modules = dict(root.named_modules())
    seen = set()
    with torch.no_grad():
        for key, t in state.items():
            if ".lora." not in key:
                raise KeyError(f"not a LoRA entry: {key}")
            mod_name, matrix = key.split(".lora.")
            matrix = matrix[:-len(".weight")] if matrix.endswith(".weight") else matrix
            m = modules.get(mod_name)
            if m is None or not hasattr(m, "set_lora_layer"):
                raise KeyError(f"checkpoint names a module this UNet does not have: {mod_name}")
            if getattr(m, "lora_layer", None) is None:
                rank = state[f"{mod_name}.lora.down.weight"].shape[0]
                m.set_lora_layer(LoRALinearLayer(m.in_features, m.out_features, rank).to(m.weight.device))
            p = getattr(m.lora_layer, matrix).weight
            if tuple(p.shape) != tuple(t.shape):
                raise ValueError(f"{key}: checkpoint shape {tuple(t.shape)} vs model {tuple(p.shape)}")
            p.copy_(t.to(p.device, p.dtype))           # in place: the optimiser's flat buffer keeps owning the storage
            seen.add((mod_name, matrix))
    for mod_name, _ in list(seen):
        if not {(mod_name, "down"), (mod_name, "up")} <= seen:
            raise KeyError(f"adapter {mod_name} is missing its down or up matrix")
    if hasattr(unet, "refresh_lora"):
        unet.refresh_lora()                            # 16-bit operand images / folded weights follow the fp32 masters
    return len(seen)


def checkpoint_dir(output_dir: str, global_step: int) -> str:
    return os.path.join(output_dir, f"checkpoint-{global_step}")                    # :497-499


def latest_checkpoint(output_dir: str) -> Optional[str]:
    """``--resume_from_checkpoint latest`` (:166-171): highest ``checkpoint-<n>`` directory, or None."""
    if not os.path.isdir(output_dir):
        return None
    dirs = [d for d in os.listdir(output_dir) if d.startswith("checkpoint")]
    dirs = sorted(dirs, key=lambda x: int(x.split("-")[1]))
    return os.path.join(output_dir, dirs[-1]) if dirs else None


def _optim_state(opt):
    return {"m": opt.m.detach().cpu(), "v": opt.v.detach().cpu(), "step_count": opt.step_count, "n": opt.n}


def _load_optim_state(opt, st):
    if st["n"] != opt.n:
        raise ValueError(f"optimiser state has {st['n']} elements, the model has {opt.n}")
    opt.m.copy_(st["m"])
    opt.v.copy_(st["v"])
    opt.step_count = int(st["step_count"])


def save_checkpoint(trainer, output_dir: str, global_step: Optional[int] = None, diffusers_prefix: bool = True,
                    with_trainer_state: bool = True) -> str:
    """what ``save_and_evaluate(save_path, global_step)`` persists (:382-426) for the LoRA configuration (+ trainer_state.pt)."""
    step = trainer.global_step if global_step is None else global_step
    path = checkpoint_dir(output_dir, step)
    if hasattr(trainer, "sync"):
        trainer.sync()                                 # outstanding side-stream optimiser tails land before the parameters are read
    text = text_encoder_lora_state_dict(trainer.pipeline.text_encoder) if getattr(trainer, "train_text", False) else None     # :386-389
    save_lora_weights(path, unet_lora_state_dict(trainer.pipeline.unet), diffusers_prefix, text_encoder_lora_layers=text)
    if trainer.D is not None:
        d_dir = os.path.join(path, "D_sd")
        save_lora_weights(d_dir, unet_lora_state_dict(trainer.D.unet), diffusers_prefix)
        torch.save({k: v.detach().cpu() for k, v in trainer.D.mlp.state_dict().items()}, os.path.join(d_dir, "mlp.pt"))
    if with_trainer_state:
        st = {"global_step": step, "G": _optim_state(trainer.optimizer), "python_rng": trainer.rng.getstate(),
              "python_global_rng": random.getstate(), "torch_rng": torch.get_rng_state()}
        if trainer.D is not None and trainer.D_optimizer is not None:
            st["D"] = _optim_state(trainer.D_optimizer)
        if torch.cuda.is_available():
            st["cuda_rng"] = torch.cuda.get_rng_state_all()
        torch.save(st, os.path.join(path, STATE_FILE))
    return path


def load_checkpoint(trainer, path_or_output_dir: str, resume: str = "latest") -> Optional[int]:
    """the resume block (:156-205).  ``resume='latest'`` picks the newest ``checkpoint-<n>`` under the directory and - like the
    reference (:191) - only then restores the discriminator; an explicit checkpoint path restores the generator only.
    Returns the restored ``global_step`` (None when there is nothing to resume from)."""
    if resume == "latest":
        path = latest_checkpoint(path_or_output_dir)
        if path is None:
            return None                                                             # :173-177 "Starting a new training run"
    else:
        path = path_or_output_dir
    if hasattr(trainer, "sync"):
        trainer.sync()
    load_lora_into_unet(lora_state_dict(os.path.join(path, LORA_FILE)), trainer.pipeline.unet)
    if getattr(trainer, "train_text", False):                                       # :182-185
        load_lora_into_text_encoder(text_lora_state_dict(os.path.join(path, LORA_FILE)), trainer.pipeline.text_encoder)
    with_d = trainer.D is not None and resume == "latest"
    if with_d:
        load_lora_into_unet(lora_state_dict(os.path.join(path, "D_sd", LORA_FILE)), trainer.D.unet)
        head = torch.load(os.path.join(path, "D_sd", "mlp.pt"), map_location="cpu")
        with torch.no_grad():
            for k, p in trainer.D.mlp.state_dict().items():
                p.copy_(head[k].float())                                            # :196-198 (in place; re-floated)
    step = int(os.path.basename(os.path.normpath(path)).split("-")[1])             # :203
    trainer.global_step = step
    sp = os.path.join(path, STATE_FILE)
    if os.path.exists(sp):
        st = torch.load(sp, map_location="cpu", weights_only=False)
        _load_optim_state(trainer.optimizer, st["G"])
        if with_d and "D" in st and trainer.D_optimizer is not None:
            _load_optim_state(trainer.D_optimizer, st["D"])
        trainer.rng.setstate(st["python_rng"])
        random.setstate(st["python_global_rng"])
        torch.set_rng_state(st["torch_rng"])
        if "cuda_rng" in st and torch.cuda.is_available() and len(st["cuda_rng"]) == torch.cuda.device_count():
            torch.cuda.set_rng_state_all(st["cuda_rng"])
    return step
