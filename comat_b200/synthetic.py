"""Random-init weights at the real geometries + synthetic batches — there is no Hub / dataset access.

Used by bench.py, smoke() and the GPU tests.  Everything is seeded; weights are created directly on the device.
"""
from __future__ import annotations

import random
from types import SimpleNamespace
from typing import Dict

import torch

from . import containers as Cn


def build_sd15(device, dtype=torch.float16, rank=128, seed=42, lora_up_std=0.0, tiny=False):
    """(unet, vae) parameter containers at SD1.5 geometry (859.5 M + 49.5 M params), random init, LoRA installed."""
    torch.manual_seed(seed)
    with torch.device(device):
        if tiny:
            unet = Cn.UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=64)
            vae = Cn.AutoencoderKL(block_out_channels=(64, 64, 128, 128))
        else:
            unet = Cn.UNet2DConditionModel()
            vae = Cn.AutoencoderKL()
    unet.requires_grad_(False)
    vae.requires_grad_(False)
    unet.install_lora(rank, up_std=lora_up_std)
    return unet, vae


def build_sdxl_unet(device, rank=128, seed=42, tiny=False):
    torch.manual_seed(seed)
    with torch.device(device):
        if tiny:
            unet = Cn.UNet2DConditionModel(**{**Cn.SDXL_UNET, "block_out_channels": (64, 128, 256), "heads": (2, 4, 8),
                                              "cross_attention_dim": 64, "transformer_layers": (1, 1, 2),
                                              "addition_time_embed_dim": 8, "projection_class_embeddings_input_dim": 64})
        else:
            unet = Cn.UNet2DConditionModel(**Cn.SDXL_UNET)
    unet.requires_grad_(False)
    unet.install_lora(rank)
    return unet


def build_blip(device, dtype=torch.float16, seed=0, large=True, label_smoothing=0.1):
    """Random-init HF BlipForConditionalGeneration at blip-image-captioning-large geometry: the parameter
    container for the captioner.  Non-degenerate vision init (default initializer_range 1e-10 kills image gradients).
    label_smoothing 0.1 follows the transformers==4.31.0 pin (requirements.txt:1)."""
    from transformers import BlipConfig, BlipForConditionalGeneration
    if large:
        vis = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16)
        txt = dict(hidden_size=768, encoder_hidden_size=1024, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12)
    else:
        vis = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2)
        txt = dict(hidden_size=128, encoder_hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2)
    vis.update(image_size=384, patch_size=16, initializer_range=0.02, attention_dropout=0.0)
    txt.update(vocab_size=30524, max_position_embeddings=512, label_smoothing=label_smoothing, hidden_dropout_prob=0.0,
               attention_probs_dropout_prob=0.0, bos_token_id=30522, pad_token_id=0, sep_token_id=102)
    cfg = BlipConfig(vision_config=vis, text_config=txt)
    cfg.label_smoothing = label_smoothing
    torch.manual_seed(seed)
    m = BlipForConditionalGeneration(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.startswith("vision_model") and p.ndim >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    m.eval().requires_grad_(False)
    return m.to(device=device, dtype=dtype)


def build_clip_text(device, dtype=torch.float16, seed=7, which="clip_l", tiny=False, **overrides):
    """Random-init HF CLIP text tower at the geometry the SD checkpoints ship: ``clip_l`` = CLIPTextModel 12 x 768,
    12 heads, quick-GELU (SD1.5 ``text_encoder`` / SDXL ``text_encoder``); ``bigg`` = CLIPTextModelWithProjection 32 x 1280,
    20 heads, GELU, projection 1280 (SDXL ``text_encoder_2``).  ``eos_token_id = 2`` is the legacy value those configs carry."""
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    if which == "clip_l":
        kw = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, hidden_act="quick_gelu",
                  projection_dim=768)
        cls = CLIPTextModel
    elif which == "bigg":
        kw = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20, hidden_act="gelu",
                  projection_dim=1280)
        cls = CLIPTextModelWithProjection
    else:
        raise NotImplementedError(which)
    if tiny:
        kw.update(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, projection_dim=64)
    kw.update(overrides)
    cfg = CLIPTextConfig(vocab_size=49408, max_position_embeddings=77, eos_token_id=2, bos_token_id=0, pad_token_id=1, **kw)
    torch.manual_seed(seed)
    with torch.device(device):
        m = cls(cfg)
    m.eval().requires_grad_(False)
    return m.to(dtype=dtype)


class SyntheticClipTokenizer:
    """Stand-in for ``CLIPTokenizer`` (its vocabulary / merges files are not on disk and there is no Hub access): the same call
    protocol and framing - BOS 49406, one id per whitespace word (a stable hash into the BPE id range), EOS 49407, padding with
    ``pad_token_id`` (49407 for SD1.5 / SDXL tokenizer 1, 0 for SDXL tokenizer 2) - so ``encode_prompt`` runs on prompt strings.
    Not a BPE: real deployments pass the real tokenizer, which has this interface."""

    bos_token_id, eos_token_id = 49406, 49407

    def __init__(self, model_max_length: int = 77, pad_token_id: int = 49407):
        self.model_max_length, self.pad_token_id = model_max_length, pad_token_id

    @staticmethod
    def _word_id(w: str) -> int:
        h = 2166136261
        for ch in w.encode("utf-8"):
            h = ((h ^ ch) * 16777619) & 0xFFFFFFFF            # FNV-1a: stable across processes (hash() is salted)
        return 1000 + h % 48000

    def __call__(self, text, padding="max_length", max_length=None, truncation=True, return_tensors="pt", **_):
        texts = [text] if isinstance(text, str) else list(text)
        rows = [[self.bos_token_id] + [self._word_id(w) for w in t.lower().split()] + [self.eos_token_id] for t in texts]
        limit = max_length or self.model_max_length
        if truncation:
            rows = [r if len(r) <= limit else r[:limit - 1] + [self.eos_token_id] for r in rows]
        T = limit if padding == "max_length" else max(len(r) for r in rows)
        ids = torch.full((len(rows), T), self.pad_token_id, dtype=torch.long)
        am = torch.zeros(len(rows), T, dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = torch.tensor(r)
            am[i, :len(r)] = 1
        return SimpleNamespace(input_ids=ids, attention_mask=am)


class SyntheticBertTokenizer:
    """Stand-in for the BLIP processor's ``BertTokenizer`` (vocabulary not on disk): [CLS] 101 + one id per whitespace word +
    [SEP] 102, right-padded with 0 to the longest row; 'a photography of' maps to the real wordpiece ids 1037 5855 1997 so
    ``prompt_length = len(tok('a photography of').input_ids) - 1 = 4`` as in caption_blip.py:38-39.  Other words hash (FNV-1a)
    into [1000, 30000) like SURVEY 8d's synthetic ids."""

    pad_token_id, cls_token_id, sep_token_id = 0, 101, 102
    _FIXED = {"a": 1037, "photography": 5855, "of": 1997}

    def _ids(self, text: str):
        return [self.cls_token_id] + [self._FIXED.get(w) or 1000 + SyntheticClipTokenizer._word_id(w) % 29000 for w in text.lower().split()] + [self.sep_token_id]

    def __call__(self, text, padding="longest", return_tensors=None, **_):
        if isinstance(text, str):
            ids = self._ids(text)
            if return_tensors is None:
                return SimpleNamespace(input_ids=ids, attention_mask=[1] * len(ids))
            text = [text]
        rows = [self._ids(t) for t in text]
        T = max(len(r) for r in rows)
        ids = torch.zeros(len(rows), T, dtype=torch.long)
        am = torch.zeros(len(rows), T, dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = torch.tensor(r)
            am[i, :len(r)] = 1
        return SimpleNamespace(input_ids=ids, attention_mask=am)


def random_mask(g, size=512, empty=False):
    m = torch.zeros(1, 1, size, size, dtype=torch.bool)
    if empty:
        return m
    for _ in range(int(torch.randint(1, 3, (1,), generator=g))):
        h = int(torch.randint(size // 5, size * 6 // 10, (1,), generator=g))
        w = int(torch.randint(size // 5, size * 6 // 10, (1,), generator=g))
        y = int(torch.randint(0, size - h, (1,), generator=g))
        x = int(torch.randint(0, size - w, (1,), generator=g))
        m[..., y:y + h, x:x + w] = True
    return m


def synthetic_batch(B: int, seed: int, ctx_dim: int = 768, res: int = 512, attrcon: bool = True, gan: bool = True,
                    pinned: bool = False, pooled_dim: int = 0, gan_ctx_dim: int = 0) -> Dict:
    """Host-side batch (what a dataloader would hand over): prompt / null embeddings, BLIP token ids, attribute token
    lists + per-word masks, 'real' latents for the discriminator.  ``pooled_dim`` > 0 adds SDXL's pooled text
    embeddings (B, pooled_dim); ``gan_ctx_dim``: context width of the discriminator's own null embedding when it differs from the
    generator's (SDXL generator + SD1.5 discriminator, scripts/sdxl.sh:15)."""
    g = torch.Generator().manual_seed(seed)
    rr = random.Random(seed)
    lat = res // 8
    b: Dict = {
        "prompt_embeds": torch.randn(B, 77, ctx_dim, generator=g),
        "null_embeds": torch.randn(1, 77, ctx_dim, generator=g).expand(B, -1, -1).contiguous(),
    }
    rows = []
    for _ in range(B):
        L = max(4, min(40, round(rr.gauss(14, 4))))
        rows.append([101, 1037, 5855, 1997] + [rr.randrange(1000, 30000) for _ in range(L)] + [102])
    T = max(len(r) for r in rows)
    ids = torch.zeros(B, T, dtype=torch.long)
    am = torch.zeros(B, T, dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, :len(r)] = torch.tensor(r)
        am[i, :len(r)] = 1
    b["blip"] = {"input_ids": ids, "attention_mask": am}
    if attrcon:
        words, masks = [], []
        for _ in range(B):
            nw = rr.randint(1, 3)
            pos = rr.sample(range(1, 40), 9)
            ws = []
            for _ in range(nw):
                k = rr.randint(1, 3)
                ws.append([pos.pop() for _ in range(k)])
            words.append(ws)
            masks.append(torch.cat([random_mask(g, res, empty=(rr.random() < 0.1)) for _ in range(nw)]))   # (nw,1,res,res)
        b["words"], b["masks_host"] = words, masks
    if pooled_dim:
        b["pooled_prompt_embeds"] = torch.randn(B, pooled_dim, generator=g)
        b["pooled_null_embeds"] = torch.randn(1, pooled_dim, generator=g).expand(B, -1).contiguous()
    if gan:
        b["gan_null_embeds"] = (b["null_embeds"].clone() if not gan_ctx_dim or gan_ctx_dim == ctx_dim else
                                torch.randn(1, 77, gan_ctx_dim, generator=g).expand(B, -1, -1).contiguous())
        b["real_latents"] = torch.randn(B, 4, lat, lat, generator=g)
    if pinned:
        for k, v in list(b.items()):
            if torch.is_tensor(v):
                b[k] = v.pin_memory()
        b["blip"] = {k: v.pin_memory() for k, v in b["blip"].items()}
        if attrcon:
            b["masks_host"] = [m.pin_memory() for m in b["masks_host"]]
    return b


def batch_to_device(b: Dict, device) -> (Dict, int):
    """H2D copy of one step's inputs (non_blocking from pinned memory); returns (device batch, bytes copied)."""
    out, nbytes = {}, 0
    for k, v in b.items():
        if torch.is_tensor(v):
            out[k] = v.to(device, non_blocking=True)
            nbytes += v.numel() * v.element_size()
        elif k == "blip":
            out[k] = {kk: vv.to(device, non_blocking=True) for kk, vv in v.items()}
            nbytes += sum(vv.numel() * vv.element_size() for vv in v.values())
        elif k == "masks_host":
            dm = [m.to(device, non_blocking=True) for m in v]
            nbytes += sum(m.numel() * m.element_size() for m in v)
            out["masks"] = [[m[i:i + 1] for i in range(m.shape[0])] for m in dm]
        else:
            out[k] = v
    return out, nbytes


def default_args(**over):
    from .arguments import parse_args
    a = parse_args([])
    for k, v in over.items():
        setattr(a, k, v)
    a.do_classifier_free_guidance = a.cfg_scale > 1.0
    return a
