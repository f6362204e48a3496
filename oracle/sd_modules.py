"""Plain-PyTorch restatement of the diffusers (0.22-0.25) modules the CoMat hot path executes.

ORACLE / TEST INFRASTRUCTURE — not imported by the product (see oracle/__init__.py).
**Parity unpinned by upstream**: diffusers is not vendored in /root/reference and cannot be
installed here; this file restates its published architecture and is
anchored by parameter totals (859 520 964 / 2 567 463 684 / 49 490 199) and by being hookable by the
reference's own ``register_attention_control`` (attn_utils/tc_attn_utils.py:96-196).

Module *class names* and *attribute names* follow diffusers so that (a) diffusers state-dict keys
load unchanged and (b) the reference's hook, which matches ``__class__.__name__ == 'Attention'``
(tc_attn_utils.py:166) and touches ``to_q/to_k/to_v/to_out/head_to_batch_dim/...``
(tc_attn_utils.py:98-159), works on these modules verbatim.

Named numerical conventions (cannot be checked offline; recorded as explicit choices):
  GN eps 1e-5 in UNet resnets / 1e-6 in Transformer2DModel and the VAE; GEGLU = hidden * gelu(gate)
  with ``hidden, gate = proj(x).chunk(2, -1)``; sinusoid flip_sin_to_cos=True, freq_shift=0;
  attention scale d^-0.5; Upsample = nearest x2 then conv3x3; Downsample = conv3x3 stride 2 pad 1;
  skip concat order cat([hidden, skip], dim=1); SDXL add_embedding input cat([text_embeds, time_ids_emb]).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# LoRA  (diffusers.models.lora; call sites training_utils/pipeline.py:94-115)
# --------------------------------------------------------------------------------------------
class LoRALinearLayer(nn.Module):
    def __init__(self, in_features, out_features, rank=4, network_alpha=None, device=None, dtype=None):
        super().__init__()
        self.down = nn.Linear(in_features, rank, bias=False, device=device, dtype=dtype)
        self.up = nn.Linear(rank, out_features, bias=False, device=device, dtype=dtype)
        self.network_alpha = network_alpha
        self.rank = rank
        self.in_features = in_features
        self.out_features = out_features
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)

    def forward(self, hidden_states):
        orig_dtype = hidden_states.dtype
        dtype = self.down.weight.dtype
        down = self.down(hidden_states.to(dtype))
        up = self.up(down)
        if self.network_alpha is not None:
            up = up * (self.network_alpha / self.rank)
        return up.to(orig_dtype)


class LoRACompatibleLinear(nn.Linear):
    def __init__(self, *args, lora_layer: Optional[LoRALinearLayer] = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def set_lora_layer(self, lora_layer):
        self.lora_layer = lora_layer

    def forward(self, hidden_states, scale: float = 1.0):
        if self.lora_layer is None:
            return super().forward(hidden_states)
        return super().forward(hidden_states) + scale * self.lora_layer(hidden_states)


class LoRACompatibleConv(nn.Conv2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = None

    def set_lora_layer(self, lora_layer):
        self.lora_layer = lora_layer

    def forward(self, hidden_states, scale: float = 1.0):
        return super().forward(hidden_states)


# --------------------------------------------------------------------------------------------
# embeddings
# --------------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0,
                           scale=1.0, max_period=10000):
    half_dim = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half_dim, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half_dim - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, flip_sin_to_cos=self.flip_sin_to_cos,
                                      downscale_freq_shift=self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


# --------------------------------------------------------------------------------------------
# Attention (attribute surface = what tc_attn_utils.py:98-159 touches)
# --------------------------------------------------------------------------------------------
class AttnProcessor:
    """Default processor: same arithmetic as the reference's hooked forward (tc_attn_utils.py:104-161)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        residual = hidden_states
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            b, c, h, w = hidden_states.shape
            hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
        if attn.group_norm is not None:
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        query = attn.head_to_batch_dim(query)
        key = attn.head_to_batch_dim(key)
        value = attn.head_to_batch_dim(value)
        probs = attn.get_attention_scores(query, key, attention_mask)
        hidden_states = torch.bmm(probs, value)
        hidden_states = attn.batch_to_head_dim(hidden_states)
        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        if input_ndim == 4:
            hidden_states = hidden_states.transpose(-1, -2).reshape(b, c, h, w)
        if attn.residual_connection:
            hidden_states = hidden_states + residual
        return hidden_states / attn.rescale_output_factor


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False,
                 norm_num_groups=None, eps=1e-5, residual_connection=False, rescale_output_factor=1.0,
                 out_bias=True, upcast_softmax=False):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.upcast_softmax = upcast_softmax
        self.upcast_attention = False
        self.residual_connection = residual_connection
        self.rescale_output_factor = rescale_output_factor
        self.spatial_norm = None
        self.norm_cross = None
        self.group_norm = (nn.GroupNorm(num_channels=query_dim, num_groups=norm_num_groups, eps=eps, affine=True)
                           if norm_num_groups is not None else None)
        self.to_q = LoRACompatibleLinear(query_dim, self.inner_dim, bias=bias)
        self.to_k = LoRACompatibleLinear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = LoRACompatibleLinear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([LoRACompatibleLinear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(0.0)])
        self.processor = AttnProcessor()

    def head_to_batch_dim(self, tensor):
        b, s, d = tensor.shape
        h = self.heads
        return tensor.reshape(b, s, h, d // h).permute(0, 2, 1, 3).reshape(b * h, s, d // h)

    def batch_to_head_dim(self, tensor):
        bh, s, d = tensor.shape
        h = self.heads
        return tensor.reshape(bh // h, h, s, d).permute(0, 2, 1, 3).reshape(bh // h, s, d * h)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        return attention_mask  # always None on this path

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        baddbmm_input = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device)
        scores = torch.baddbmm(baddbmm_input, query, key.transpose(-1, -2), beta=0, alpha=self.scale)
        if self.upcast_softmax:
            scores = scores.float()
        probs = scores.softmax(dim=-1)
        return probs.to(dtype)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = LoRACompatibleLinear(dim_in, dim_out * 2)

    def forward(self, hidden_states, scale: float = 1.0):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(0.0), LoRACompatibleLinear(inner, dim)])

    def forward(self, hidden_states):
        for m in self.net:
            hidden_states = m(hidden_states)
        return hidden_states


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(query_dim=dim, heads=heads, dim_head=dim_head, bias=False)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(query_dim=dim, cross_attention_dim=cross_attention_dim, heads=heads,
                               dim_head=dim_head, bias=False)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, hidden_states, encoder_hidden_states=None):
        hidden_states = self.attn1(self.norm1(hidden_states)) + hidden_states
        hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states) + hidden_states
        hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, num_layers, cross_attention_dim, norm_num_groups=32,
                 use_linear_projection=False):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = LoRACompatibleLinear(in_channels, inner)
        else:
            self.proj_in = LoRACompatibleConv(in_channels, inner, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        if use_linear_projection:
            self.proj_out = LoRACompatibleLinear(inner, in_channels)
        else:
            self.proj_out = LoRACompatibleConv(inner, in_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, hidden_states, encoder_hidden_states=None):
        b, c, h, w = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        if not self.use_linear_projection:
            hidden_states = self.proj_in(hidden_states)
            inner = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(b, h * w, inner)
        else:
            inner = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(b, h * w, inner)
            hidden_states = self.proj_in(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, encoder_hidden_states=encoder_hidden_states)
        if not self.use_linear_projection:
            hidden_states = hidden_states.reshape(b, h, w, inner).permute(0, 3, 1, 2).contiguous()
            hidden_states = self.proj_out(hidden_states)
        else:
            hidden_states = self.proj_out(hidden_states)
            hidden_states = hidden_states.reshape(b, h, w, inner).permute(0, 3, 1, 2).contiguous()
        return hidden_states + residual


# --------------------------------------------------------------------------------------------
# ResNet / sampling blocks
# --------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels=1280, groups=32, eps=1e-5, output_scale_factor=1.0):
        super().__init__()
        self.norm1 = nn.GroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = LoRACompatibleConv(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = LoRACompatibleLinear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(num_groups=groups, num_channels=out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = LoRACompatibleConv(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = nn.SiLU()
        self.output_scale_factor = output_scale_factor
        self.conv_shortcut = (LoRACompatibleConv(in_channels, out_channels, kernel_size=1, stride=1, padding=0)
                              if in_channels != out_channels else None)

    def forward(self, input_tensor, temb=None):
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        if self.time_emb_proj is not None and temb is not None:
            h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = LoRACompatibleConv(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = LoRACompatibleConv(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample, **_):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels)
                                      for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None):
        outs = ()
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class CrossAttnDownBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample, heads, cross_attention_dim,
                 transformer_layers=1, use_linear_projection=False):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels)
                                      for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(heads, out_channels // heads, out_channels, transformer_layers, cross_attention_dim,
                               use_linear_projection=use_linear_projection) for _ in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None):
        outs = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, in_channels, temb_channels, heads, cross_attention_dim, transformer_layers=1,
                 use_linear_projection=False):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels),
                                      ResnetBlock2D(in_channels, in_channels, temb_channels)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(heads, in_channels // heads, in_channels, transformer_layers, cross_attention_dim,
                               use_linear_projection=use_linear_projection)])

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        hidden_states = self.attentions[0](hidden_states, encoder_hidden_states=encoder_hidden_states)
        return self.resnets[1](hidden_states, temb)


class UpBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers, add_upsample, **_):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(rin + skip, out_channels, temb_channels))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None):
        for resnet in self.resnets:
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class CrossAttnUpBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers, add_upsample, heads,
                 cross_attention_dim, transformer_layers=1, use_linear_projection=False):
        super().__init__()
        resnets, attns = [], []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(rin + skip, out_channels, temb_channels))
            attns.append(Transformer2DModel(heads, out_channels // heads, out_channels, transformer_layers,
                                            cross_attention_dim, use_linear_projection=use_linear_projection))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList(attns)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None):
        for resnet, attn in zip(self.resnets, self.attentions):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


# --------------------------------------------------------------------------------------------
# UNet2DConditionModel
# --------------------------------------------------------------------------------------------
SD15_UNET_CONFIG = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    attention_head_dim=8, num_attention_heads=None, cross_attention_dim=768, transformer_layers_per_block=1,
    use_linear_projection=False, addition_embed_type=None, addition_time_embed_dim=None,
    projection_class_embeddings_input_dim=None, norm_num_groups=32,
)
SDXL_UNET_CONFIG = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280), layers_per_block=2,
    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
    up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
    attention_head_dim=(5, 10, 20), num_attention_heads=None, cross_attention_dim=2048,
    transformer_layers_per_block=(1, 2, 10), use_linear_projection=True, addition_embed_type="text_time",
    addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816, norm_num_groups=32,
)


def tiny_unet_config(sdxl=False, width=32, cross_attention_dim=64):
    """Reduced geometry with the same topology (for CPU-speed tests). head dim stays a multiple of 8."""
    if not sdxl:
        cfg = dict(SD15_UNET_CONFIG)
        cfg.update(block_out_channels=(width, 2 * width, 4 * width, 4 * width), attention_head_dim=4,
                   cross_attention_dim=cross_attention_dim)
    else:
        cfg = dict(SDXL_UNET_CONFIG)
        cfg.update(block_out_channels=(width, 2 * width, 4 * width), attention_head_dim=(2, 4, 8),
                   cross_attention_dim=cross_attention_dim, transformer_layers_per_block=(1, 1, 2),
                   addition_time_embed_dim=8, projection_class_embeddings_input_dim=6 * 8 + 16)
    return cfg


class UNet2DConditionModel(nn.Module):
    def __init__(self, **config):
        super().__init__()
        cfg = dict(SD15_UNET_CONFIG)
        cfg.update(config)
        self.config = SimpleNamespace(**cfg)
        boc = tuple(cfg["block_out_channels"])
        n_blocks = len(boc)
        time_embed_dim = boc[0] * 4
        heads_cfg = cfg["num_attention_heads"] or cfg["attention_head_dim"]
        heads = (heads_cfg,) * n_blocks if isinstance(heads_cfg, int) else tuple(heads_cfg)
        tl = cfg["transformer_layers_per_block"]
        tl = (tl,) * n_blocks if isinstance(tl, int) else tuple(tl)
        cad = cfg["cross_attention_dim"]
        ulp = cfg["use_linear_projection"]
        lpb = cfg["layers_per_block"]

        self.conv_in = LoRACompatibleConv(cfg["in_channels"], boc[0], kernel_size=3, padding=1)
        self.time_proj = Timesteps(boc[0], True, 0)
        self.time_embedding = TimestepEmbedding(boc[0], time_embed_dim)
        if cfg["addition_embed_type"] == "text_time":
            self.add_time_proj = Timesteps(cfg["addition_time_embed_dim"], True, 0)
            self.add_embedding = TimestepEmbedding(cfg["projection_class_embeddings_input_dim"], time_embed_dim)

        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, typ in enumerate(cfg["down_block_types"]):
            in_ch, out_ch = out_ch, boc[i]
            final = i == n_blocks - 1
            if typ == "CrossAttnDownBlock2D":
                blk = CrossAttnDownBlock2D(in_ch, out_ch, time_embed_dim, lpb, not final, heads[i], cad, tl[i], ulp)
            else:
                blk = DownBlock2D(in_ch, out_ch, time_embed_dim, lpb, not final)
            self.down_blocks.append(blk)

        self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], time_embed_dim, heads[-1], cad, tl[-1], ulp)

        self.up_blocks = nn.ModuleList()
        rboc, rheads, rtl = boc[::-1], heads[::-1], tl[::-1]
        out_ch = rboc[0]
        for i, typ in enumerate(cfg["up_block_types"]):
            prev, out_ch = out_ch, rboc[i]
            in_ch = rboc[min(i + 1, n_blocks - 1)]
            final = i == n_blocks - 1
            if typ == "CrossAttnUpBlock2D":
                blk = CrossAttnUpBlock2D(in_ch, prev, out_ch, time_embed_dim, lpb + 1, not final, rheads[i], cad, rtl[i], ulp)
            else:
                blk = UpBlock2D(in_ch, prev, out_ch, time_embed_dim, lpb + 1, not final)
            self.up_blocks.append(blk)

        self.conv_norm_out = nn.GroupNorm(num_channels=boc[0], num_groups=cfg["norm_num_groups"], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = LoRACompatibleConv(boc[0], cfg["out_channels"], kernel_size=3, padding=1)
        self.gradient_checkpointing = False

    # --- diffusers surface used by training_utils/pipeline.py:84-187 ---
    @property
    def attn_processors(self):
        procs = {}
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                procs[f"{name}.processor"] = m.processor
        return procs

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def enable_gradient_checkpointing(self):
        self.gradient_checkpointing = True  # arithmetic-neutral; the oracle never recomputes

    def forward(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None,
                added_cond_kwargs=None, return_dict=True):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.int64, device=sample.device)
        elif timesteps.ndim == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        t_emb = self.time_proj(timesteps).to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb)
        if self.config.addition_embed_type == "text_time":
            text_embeds = added_cond_kwargs["text_embeds"]
            time_ids = added_cond_kwargs["time_ids"]
            time_embeds = self.add_time_proj(time_ids.flatten()).reshape(text_embeds.shape[0], -1)
            add_embeds = torch.cat([text_embeds, time_embeds], dim=-1).to(emb.dtype)
            emb = emb + self.add_embedding(add_embeds)

        sample = self.conv_in(sample)
        down_res = (sample,)
        for blk in self.down_blocks:
            sample, res = blk(sample, emb, encoder_hidden_states)
            down_res += res
        sample = self.mid_block(sample, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            res, down_res = down_res[-n:], down_res[:-n]
            sample = blk(sample, res, emb, encoder_hidden_states)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        if not return_dict:
            return (sample,)
        return SimpleNamespace(sample=sample)


# --------------------------------------------------------------------------------------------
# AutoencoderKL.decode   (call site TrainableSDPipeline.py:220)
# --------------------------------------------------------------------------------------------
class UNetMidBlock2D(nn.Module):
    def __init__(self, ch, groups=32, eps=1e-6):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, None, groups, eps), ResnetBlock2D(ch, ch, None, groups, eps)])
        self.attentions = nn.ModuleList([
            Attention(ch, heads=1, dim_head=ch, bias=True, norm_num_groups=groups, eps=eps, residual_connection=True,
                      rescale_output_factor=1.0, upcast_softmax=True)])

    def forward(self, x):
        x = self.resnets[0](x, None)
        x = self.attentions[0](x)
        return self.resnets[1](x, None)


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_ch, out_ch, num_layers, add_upsample, groups=32, eps=1e-6):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, None, groups, eps)
                                      for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x, None)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                x = u(x)
        return x


class Decoder(nn.Module):
    def __init__(self, in_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 groups=32):
        super().__init__()
        rboc = tuple(block_out_channels)[::-1]
        self.conv_in = nn.Conv2d(in_channels, rboc[0], 3, padding=1)
        self.mid_block = UNetMidBlock2D(rboc[0], groups)
        self.up_blocks = nn.ModuleList()
        out_ch = rboc[0]
        for i in range(len(rboc)):
            prev, out_ch = out_ch, rboc[i]
            self.up_blocks.append(UpDecoderBlock2D(prev, out_ch, layers_per_block + 1, i != len(rboc) - 1, groups))
        self.conv_norm_out = nn.GroupNorm(num_channels=rboc[-1], num_groups=groups, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(rboc[-1], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for blk in self.up_blocks:
            x = blk(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    """Decoder half only (the hot path never encodes)."""

    def __init__(self, block_out_channels=(128, 256, 512, 512), latent_channels=4, scaling_factor=0.18215,
                 groups=32, force_upcast=False):
        super().__init__()
        self.config = SimpleNamespace(scaling_factor=scaling_factor, force_upcast=force_upcast,
                                      block_out_channels=tuple(block_out_channels), latent_channels=latent_channels)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.decoder = Decoder(latent_channels, 3, block_out_channels, 2, groups)

    @property
    def dtype(self):
        return self.post_quant_conv.weight.dtype

    def decode(self, z, return_dict=True):
        x = self.decoder(self.post_quant_conv(z))
        if not return_dict:
            return (x,)
        return SimpleNamespace(sample=x)


# --------------------------------------------------------------------------------------------
# DDPMScheduler / rescale_noise_cfg   (call sites TrainableSDPipeline.py:95,136,161,166)
# --------------------------------------------------------------------------------------------
class DDPMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1,
                 variance_type="fixed_small"):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule="scaled_linear", steps_offset=steps_offset,
                                      timestep_spacing="leading", clip_sample=False, prediction_type="epsilon",
                                      variance_type=variance_type)
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    @classmethod
    def from_config(cls, config, **kw):
        return cls(**kw) if not isinstance(config, dict) else cls(**{**config, **kw})

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (torch.arange(0, num_inference_steps) * ratio).flip(0).to(torch.int64) + self.config.steps_offset
        self.timesteps = ts.to(device) if device is not None else ts

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coefficients(self, t: int):
        """(alpha_bar_t, alpha_bar_prev, beta_t_eff, sigma) as python floats for timestep t."""
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        var = torch.clamp((b_prev / b_t) * cur_beta, min=1e-20)
        return a_t, a_prev, b_t, b_prev, cur_alpha, cur_beta, var

    def step(self, model_output, timestep, sample, generator=None, return_dict=True, variance_noise=None):
        t = int(timestep)
        a_t, a_prev, b_t, b_prev, cur_alpha, cur_beta, var = self.coefficients(t)
        pred_x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        c_x0 = (a_prev ** 0.5 * cur_beta) / b_t
        c_xt = cur_alpha ** 0.5 * b_prev / b_t
        prev = c_x0 * pred_x0 + c_xt * sample
        if t > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                             dtype=model_output.dtype)
            prev = prev + (var ** 0.5) * variance_noise
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev, pred_original_sample=pred_x0)


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    std_text = noise_pred_text.std(dim=list(range(1, noise_pred_text.ndim)), keepdim=True)
    std_cfg = noise_cfg.std(dim=list(range(1, noise_cfg.ndim)), keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def install_lora(unet: UNet2DConditionModel, rank: int, up_std: float = 0.0, seed: Optional[int] = None):
    """training_utils/pipeline.py:84-115 + :123-143 — LoRA(rank) on to_q/to_k/to_v/to_out[0] of every
    attention, fp32 params; returns the trainable-parameter list in the reference's order.
    ``up_std`` > 0 gives a non-degenerate ``up`` for gradient checks (reference zero-inits it)."""
    g = torch.Generator().manual_seed(seed) if seed is not None else None
    params = []
    for name in unet.attn_processors:
        mod = unet
        for n in name.split(".")[:-1]:
            mod = getattr(mod, n)
        for lin in (mod.to_q, mod.to_k, mod.to_v, mod.to_out[0]):
            lora = LoRALinearLayer(lin.in_features, lin.out_features, rank=rank).to(lin.weight.device)
            if g is not None:
                lora.down.weight.data.copy_(torch.randn(lora.down.weight.shape, generator=g) / rank)
            if up_std > 0:
                lora.up.weight.data.copy_(torch.randn(lora.up.weight.shape, generator=g) * up_std)
            lora.to(torch.float32)
            lin.set_lora_layer(lora)
            params.extend(lora.parameters())
    return params


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
