"""Concept-matching reward: BLIP teacher-forced caption NLL of the prompt given the generated image.

Mirrors ``Blip`` (concept_mat_utils/caption_blip.py:14-59) and ``CaptionModelWrapper`` (training_script.py:69-97).
Tokenisation is outside the hot path (no vocabulary on disk offline): token ids are passed in ``input_ids`` /
``attention_mask`` (the reference forwards ``**batch`` into ``score``).

The captioner network itself is a ``BlipEngine`` (comat_b200/blip_engine.py: native executor on the tcgen05 GEMM / attention /
LayerNorm kernels) or any object with its ``caption_loss(pixel_values, input_ids, attention_mask, labels)`` method; an HF module
is refused (no library path in the product - tests/hf_blip.py wraps one as the comparator).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import image_ops

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
IGNORE_INDEX = -100
LIBRARY_CALLS = 0


class Blip(torch.nn.Module):
    def __init__(self, model, device=None, prompt_length: int = 4, pad_token_id: int = 0, tokenizer=None):
        """``model``: a BlipEngine (frozen weights, caption_blip.py:20-21).
        ``prompt_length`` = len(tok("a photography of").input_ids) - 1 (caption_blip.py:38-39) = 4 for BERT wordpieces.
        ``tokenizer`` (optional, the BLIP processor's BertTokenizer protocol): lets ``score`` take prompt strings (:47-48)."""
        super().__init__()
        self.model = model
        self.prompt = "a photography of"
        self.prompt_length = prompt_length
        self.pad_token_id = pad_token_id
        self.tokenizer = tokenizer
        if tokenizer is not None:
            self.prompt_length = len(tokenizer(self.prompt).input_ids) - 1
            self.pad_token_id = tokenizer.pad_token_id
        if not hasattr(model, "caption_loss"):
            raise TypeError("Blip needs a comat_b200.blip_engine.BlipEngine (or an object with its caption_loss method); an HF "
                            "BlipForConditionalGeneration is not executed by the product - wrap it: BlipEngine(model, dtype)")

    def preprocess(self, images: torch.Tensor) -> torch.Tensor:
        """Resize((384,384), BICUBIC, antialias) + Normalize(CLIP mean/std), differentiable (caption_blip.py:33-36,45)."""
        return image_ops.resize_bicubic_aa_normalize(images, 384, CLIP_MEAN, CLIP_STD)

    def score(self, images, prompts=None, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None, **_):
        if input_ids is None:
            if self.tokenizer is None or prompts is None:
                raise NotImplementedError("pass input_ids/attention_mask (BERT tokenisation of 'a photography of ' + prompt.lower()) "
                                          "or build Blip with a tokenizer")
            t = self.tokenizer([self.prompt + " " + p.lower() for p in prompts], return_tensors="pt", padding="longest")   # :47-48
            input_ids, attention_mask = t.input_ids.to(images.device), t.attention_mask.to(images.device)
        pix = self.preprocess(images)
        labels = input_ids.masked_fill(input_ids == self.pad_token_id, IGNORE_INDEX)        # caption_blip.py:51-53
        labels[:, : self.prompt_length] = IGNORE_INDEX                                       # :54
        loss = self.model.caption_loss(pix, input_ids, attention_mask, labels)
        return -loss.float()                                                                 # :57-58 (one scalar for the batch)


class CaptionModelWrapper(torch.nn.Module):
    """training_script.py:69-97: weights * reward per caption model + 'total'."""

    def __init__(self, caption_model: Sequence[str], weights: Sequence[float], blip_model: Blip):
        super().__init__()
        self.model_name = list(caption_model)
        self.weights = dict(zip(caption_model, weights))
        self.blip_model = blip_model

    def forward(self, images, prompts, text_encoder=None, return_feature=False, step=-1, batch=None):
        rewards = {}
        if "Blip" in self.model_name:
            rewards["Blip"] = self.blip_model.score(images, prompts, **(batch or {})) * self.weights["Blip"]
        rewards["total"] = sum(rewards[k] for k in self.model_name)
        return rewards
