"""Parameter containers with diffusers-compatible attribute names / state-dict keys (NO forward code).

The executors (comat_b200/engine.py) read weights from a diffusers-shaped module tree: a real
``diffusers.UNet2DConditionModel`` / ``AutoencoderKL`` when that library is installed, or these forward-less containers
(used for random-init synthetic weights — no Hub access offline — and for loading diffusers state dicts without
diffusers).  Geometry follows the published configs.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn

SD15_UNET = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 cross_down=(True, True, True, False), cross_up=(False, True, True, True), heads=8, cross_attention_dim=768,
                 transformer_layers=(1, 1, 1, 1), use_linear_projection=False, addition_embed_type=None,
                 addition_time_embed_dim=None, projection_class_embeddings_input_dim=None)
SDXL_UNET = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280), layers_per_block=2,
                 cross_down=(False, True, True), cross_up=(True, True, False), heads=(5, 10, 20), cross_attention_dim=2048,
                 transformer_layers=(1, 2, 10), use_linear_projection=True, addition_embed_type="text_time",
                 addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816)


class LoRALinearLayer(nn.Module):
    """diffusers.models.lora.LoRALinearLayer parameters: down ~ N(0, 1/rank), up = 0, fp32 (pipeline.py:94-115, 135-138)."""

    def __init__(self, in_features, out_features, rank):
        super().__init__()
        self.down = nn.Linear(in_features, rank, bias=False)
        self.up = nn.Linear(rank, out_features, bias=False)
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)


class _Lin(nn.Linear):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.lora_layer = None

    def set_lora_layer(self, l):
        self.lora_layer = l


class Attention(nn.Module):
    def __init__(self, dim, ctx_dim, heads, bias=False, groups=None, eps=1e-5):
        super().__init__()
        self.heads = heads
        self.to_q, self.to_k, self.to_v = _Lin(dim, dim, bias=bias), _Lin(ctx_dim or dim, dim, bias=bias), _Lin(ctx_dim or dim, dim, bias=bias)
        self.to_out = nn.ModuleList([_Lin(dim, dim, bias=True), nn.Identity()])
        self.group_norm = nn.GroupNorm(groups, dim, eps=eps) if groups else None


class _TBlock(nn.Module):
    def __init__(self, dim, heads, ctx):
        super().__init__()
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.attn1, self.attn2 = Attention(dim, None, heads), Attention(dim, ctx, heads)
        geglu = nn.Module()
        geglu.proj = nn.Linear(dim, 8 * dim)
        self.ff = nn.Module()
        self.ff.net = nn.ModuleList([geglu, nn.Identity(), nn.Linear(4 * dim, dim)])


class _Transformer(nn.Module):
    def __init__(self, ch, heads, ctx, layers, linear):
        super().__init__()
        self.use_linear_projection = linear
        self.norm = nn.GroupNorm(32, ch, eps=1e-6)
        self.proj_in = nn.Linear(ch, ch) if linear else nn.Conv2d(ch, ch, 1)
        self.transformer_blocks = nn.ModuleList([_TBlock(ch, heads, ctx) for _ in range(layers)])
        self.proj_out = nn.Linear(ch, ch) if linear else nn.Conv2d(ch, ch, 1)


class _Res(nn.Module):
    def __init__(self, cin, cout, temb, groups=32, eps=1e-5):
        super().__init__()
        self.norm1, self.conv1 = nn.GroupNorm(groups, cin, eps=eps), nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout) if temb else None
        self.norm2, self.conv2 = nn.GroupNorm(groups, cout, eps=eps), nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None


class _Sampler(nn.Module):
    def __init__(self, ch, stride):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=stride, padding=1)


class _Block(nn.Module):
    def __init__(self, resnets, attns, down=None, up=None):
        super().__init__()
        self.resnets = nn.ModuleList(resnets)
        if attns is not None:
            self.attentions = nn.ModuleList(attns)
        self.downsamplers = nn.ModuleList([down]) if down is not None else None
        self.upsamplers = nn.ModuleList([up]) if up is not None else None


class _TimeEmb(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1, self.linear_2 = nn.Linear(cin, dim), nn.Linear(dim, dim)


class UNet2DConditionModel(nn.Module):
    def __init__(self, **cfg):
        super().__init__()
        c = dict(SD15_UNET)
        c.update(cfg)
        self.config = SimpleNamespace(**c)
        boc = tuple(c["block_out_channels"])
        nb = len(boc)
        heads = (c["heads"],) * nb if isinstance(c["heads"], int) else tuple(c["heads"])
        tl, ctx, lin, lpb = tuple(c["transformer_layers"]), c["cross_attention_dim"], c["use_linear_projection"], c["layers_per_block"]
        temb = boc[0] * 4
        self.conv_in = nn.Conv2d(c["in_channels"], boc[0], 3, padding=1)
        self.time_embedding = _TimeEmb(boc[0], temb)
        if c["addition_embed_type"] == "text_time":
            self.add_embedding = _TimeEmb(c["projection_class_embeddings_input_dim"], temb)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i in range(nb):
            cin, out = out, boc[i]
            res = [_Res(cin if j == 0 else out, out, temb) for j in range(lpb)]
            att = [_Transformer(out, heads[i], ctx, tl[i], lin) for _ in range(lpb)] if c["cross_down"][i] else None
            self.down_blocks.append(_Block(res, att, down=_Sampler(out, 2) if i != nb - 1 else None))
        self.mid_block = _Block([_Res(boc[-1], boc[-1], temb), _Res(boc[-1], boc[-1], temb)],
                                [_Transformer(boc[-1], heads[-1], ctx, tl[-1], lin)])
        self.up_blocks = nn.ModuleList()
        rboc, rheads, rtl = boc[::-1], heads[::-1], tl[::-1]
        out = rboc[0]
        for i in range(nb):
            prev, out = out, rboc[i]
            cin = rboc[min(i + 1, nb - 1)]
            res = []
            for j in range(lpb + 1):
                skip = cin if j == lpb else out
                res.append(_Res((prev if j == 0 else out) + skip, out, temb))
            att = [_Transformer(out, rheads[i], ctx, rtl[i], lin) for _ in range(lpb + 1)] if c["cross_up"][i] else None
            self.up_blocks.append(_Block(res, att, up=_Sampler(out, 1) if i != nb - 1 else None))
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], c["out_channels"], 3, padding=1)

    @property
    def attn_processors(self):
        return {f"{n}.processor": None for n, m in self.named_modules() if isinstance(m, Attention) and n.split(".")[-1] in ("attn1", "attn2")}

    def install_lora(self, rank: int, up_std: float = 0.0):
        """training_utils/pipeline.py:84-115: LoRA(rank) on to_q/to_k/to_v/to_out[0] of every attention; returns the
        trainable parameters in the reference's order (:123-143)."""
        params = []
        for name in self.attn_processors:
            m = self.get_submodule(name.rsplit(".", 1)[0])
            for lin in (m.to_q, m.to_k, m.to_v, m.to_out[0]):
                l = LoRALinearLayer(lin.in_features, lin.out_features, rank).to(lin.weight.device, torch.float32)
                if up_std > 0:
                    nn.init.normal_(l.up.weight, std=up_std)
                lin.set_lora_layer(l)
                params.extend(l.parameters())
        return params


class AutoencoderKL(nn.Module):
    """decoder half of the SD VAE (49 490 199 parameters with post_quant_conv at the default geometry)."""

    def __init__(self, block_out_channels=(128, 256, 512, 512), scaling_factor=0.18215, groups=32):
        super().__init__()
        self.config = SimpleNamespace(scaling_factor=scaling_factor, block_out_channels=tuple(block_out_channels), force_upcast=False)
        rb = tuple(block_out_channels)[::-1]
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        d = nn.Module()
        d.conv_in = nn.Conv2d(4, rb[0], 3, padding=1)
        d.mid_block = nn.Module()
        d.mid_block.resnets = nn.ModuleList([_Res(rb[0], rb[0], None, groups, 1e-6), _Res(rb[0], rb[0], None, groups, 1e-6)])
        d.mid_block.attentions = nn.ModuleList([Attention(rb[0], None, 1, bias=True, groups=groups, eps=1e-6)])
        d.up_blocks = nn.ModuleList()
        out = rb[0]
        for i in range(len(rb)):
            prev, out = out, rb[i]
            d.up_blocks.append(_Block([_Res(prev if j == 0 else out, out, None, groups, 1e-6) for j in range(3)], None,
                                      up=_Sampler(out, 1) if i != len(rb) - 1 else None))
        d.conv_norm_out = nn.GroupNorm(groups, rb[-1], eps=1e-6)
        d.conv_out = nn.Conv2d(rb[-1], 3, 3, padding=1)
        self.decoder = d
