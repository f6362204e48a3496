"""Explicit forward/backward executor for the UNet / VAE-decoder hot path.

No autograd tape inside the networks: every op launches hand-written kernels through the C ABI (comat_b200.ops) and,
when a ``Tape`` is active, records a closure that launches the matching backward kernels.  The executor is plain
Python issuing launches on the current stream, so a whole forward (and a whole backward) can be captured in a CUDA
graph and replayed (engine.GraphedUNet) — CUDA graphs instead of a tracing compiler.

Layout: activations are 16-bit NHWC ``(n, H, W, C)`` / token-major ``(n, L, C)``; statistics, softmax, losses and
the latent chain are fp32.  Base weights are frozen (training_utils/pipeline.py:66-71): only data gradients and the
LoRA A/B weight gradients (training_utils/pipeline.py:94-115, 123-143) are produced.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

import os

from . import attention as attn_ops
from . import ops
from . import unet_weights as UW

# GroupNorm statistics accumulated in the epilogue of the GEMM / conv that produces the normalised tensor (north star: ResBlock
# conv + GN fused).  COMAT_GN_EPILOGUE=0 restores the two-pass GroupNorm everywhere (A/B measurements).
FUSE_GN_STATS = os.environ.get("COMAT_GN_EPILOGUE", "1") != "0"


class Var:
    """value + accumulated gradient (+ the GroupNorm statistics of ``v`` when the GEMM that produced it accumulated them in its
    epilogue: ``gn`` = (groups, sums tensor), consumed by ``groupnorm``)"""
    __slots__ = ("v", "g", "needs_grad", "gn")

    def __init__(self, v, needs_grad=True):
        self.v, self.g, self.needs_grad, self.gn = v, None, needs_grad, None


class Tape:
    def __init__(self):
        self.ops = []

    def record(self, fn):
        self.ops.append(fn)

    def backward(self):
        for fn in reversed(self.ops):
            fn()
        self.ops = []


def _acc(var: Var, g: torch.Tensor):
    if not var.needs_grad:
        return
    var.g = g if var.g is None else ops.elementwise("add", var.g, g)


# ------------------------------------------------------------------------------------------------------------
# parameter containers
# ------------------------------------------------------------------------------------------------------------
class ConvW:
    def __init__(self, conv: torch.nn.Conv2d, dtype, cin_pad: int = 0):
        w = conv.weight.detach()
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.k = w.shape[2]
        self.stride = conv.stride[0]
        self.bias = conv.bias.detach().float().contiguous() if conv.bias is not None else None
        self.cin_pad = cin_pad or self.cin
        if self.k == 1:
            self.w_f, self.taps_f = UW.to16(w.reshape(self.cout, self.cin), dtype), None
            self.w_d = UW.to16(w.reshape(self.cout, self.cin).t(), dtype)
            self.taps_d = None
        elif self.stride == 1:
            self.w_f, self.taps_f = UW.to16(UW.pack_conv3x3(w, self.cin_pad), dtype), UW.TAPS_3x3
            wd = UW.pack_conv3x3_dgrad(w, UW.pad_channels(self.cout))
            if self.cin_pad != self.cin:                      # padded input channels get (zero) gradient rows
                wd = torch.cat([wd, wd.new_zeros(self.cin_pad - self.cin, wd.shape[1])], 0)
            self.w_d, self.taps_d = UW.to16(wd, dtype), UW.TAPS_3x3
        else:
            wf, tf = UW.pack_conv_stride2(w)
            wd, td = UW.pack_conv_stride2_dgrad(w)
            self.w_f, self.taps_f, self.w_d, self.taps_d = UW.to16(wf, dtype), tf, UW.to16(wd, dtype), td


class LinW:
    def __init__(self, lin: torch.nn.Linear, dtype):
        w = lin.weight.detach()
        self.n, self.k = w.shape
        self.w = UW.to16(w, dtype)               # [N, K]
        self.wt = UW.to16(w.t(), dtype)          # [K, N]  (dgrad operand)
        self.bias = lin.bias.detach().float().contiguous() if lin.bias is not None else None


class LoRAW:
    """fp32 master parameters (trainable) + per-step operand images.  y = W x + up(down(x)).
    The images are views into arenas owned by the UNet executor (``UNetEngine._build_lora_arenas``), refreshed for ALL
    adapters by a handful of flat conversions (one adapter at a time it was ~1500 tiny launches per optimiser step)."""

    def __init__(self, lora, dtype):
        self.down = lora.down.weight        # (r, K) fp32 nn.Parameter
        self.up = lora.up.weight            # (N, r) fp32 nn.Parameter
        self.rank = self.down.shape[0]
        self.dtype = dtype
        self.g_down = None
        self.g_up = None
        self.wgrad = True
        self.direct, self.inv_scale = False, 1.0     # set per backward pass by modules._UNetFn
        # views assigned by the executor: 16-bit operands (and transposes), bf16 hi / lo images for the gradient projection
        self.down16 = self.up16 = self.down16_t = self.up16_t = None
        self.down_bh = self.up_bh = self.down_bl = self.up_bl = None


class _LoRAArenas:
    """All LoRA factors of one executor as flat buffers: fp32 staging, engine-dtype image, bf16 hi / lo images (fp32 range,
    ~16 mantissa bits together: operands of the product-gradient projection, where the tensor core needs one format for
    both operands), per-adapter transposed images.  Stable addresses (captured CUDA graphs keep reading them)."""

    def __init__(self, loras: List["LoRAW"], dtype):
        self.loras = loras
        dev = loras[0].down.device
        self.params = [p for l in loras for p in (l.down, l.up)]
        n = sum(p.numel() for p in self.params)
        self.f32 = torch.empty(n, dtype=torch.float32, device=dev)
        self.tmp = torch.empty(n, dtype=torch.float32, device=dev)
        self.h16 = torch.empty(n, dtype=dtype, device=dev)
        self.bh = torch.empty(n, dtype=torch.bfloat16, device=dev)
        self.bl = torch.empty(n, dtype=torch.bfloat16, device=dev)
        self.f32_views, self.t_dst, self.t_src = [], [], []
        off = 0
        for l in loras:
            for name, p in (("down", l.down), ("up", l.up)):
                k = p.numel()
                sl = slice(off, off + k)
                self.f32_views.append(self.f32[sl].view_as(p))
                v16 = self.h16[sl].view_as(p)
                setattr(l, name + "16", v16)
                setattr(l, name + "_bh", self.bh[sl].view_as(p))
                setattr(l, name + "_bl", self.bl[sl].view_as(p))
                t = torch.empty(p.shape[1], p.shape[0], dtype=dtype, device=dev)
                setattr(l, name + "16_t", t)
                self.t_dst.append(t)
                self.t_src.append(v16.t())
                off += k

    def refresh(self):
        torch._foreach_copy_(self.f32_views, [p.detach() for p in self.params])
        self.h16.copy_(self.f32)
        self.bh.copy_(self.f32)
        torch.sub(self.f32, self.bh, out=self.tmp)          # fp32 - bf16 promotes to fp32
        self.bl.copy_(self.tmp)
        torch._foreach_copy_(self.t_dst, self.t_src)


class NormW:
    def __init__(self, m, groups=None):
        self.gamma = m.weight.detach().float().contiguous()
        self.beta = m.bias.detach().float().contiguous()
        self.eps = m.eps
        self.groups = groups


# ------------------------------------------------------------------------------------------------------------
# ops (forward + recorded backward)
# ------------------------------------------------------------------------------------------------------------
def conv(tape: Optional[Tape], xs: Sequence[Var], cw: ConvW, rowvec: Optional[torch.Tensor] = None,
         residual: Optional[Var] = None, out_gn: Optional[int] = None) -> Var:
    """conv3x3 (stride 1 or 2) / conv1x1 over one or two channel segments (torch.cat fused away), fused
    + bias + per-sample time-embedding row vector + residual.  NHWC in/out.
    ``out_gn``: group count of the GroupNorm that consumes the result - its statistics ride in this GEMM's epilogue."""
    n, H, W = xs[0].v.shape[:3]
    chans = [x.v.shape[-1] for x in xs]
    koffs = [0, chans[0]] if len(xs) == 2 else [0]
    res = residual.v if residual is not None else None
    gn = None if (out_gn is None or not FUSE_GN_STATS) else (out_gn, H * W)
    if cw.k == 1:
        y = ops.gemm([x.v.reshape(n * H * W, c) for x, c in zip(xs, chans)], [cw.w_f] * len(xs), b_koff=koffs + [0],
                     bias=cw.bias, rowvec=rowvec, rows_per_group=H * W, residual=res, gn=gn)
        y, sums = y if gn is not None else (y, None)
        y = y.reshape(n, H, W, cw.cout)
    elif cw.stride == 1:
        y = ops.gemm([x.v for x in xs], [cw.w_f] * len(xs), b_koff=koffs + [0], bias=cw.bias, rowvec=rowvec,
                     rows_per_group=H * W, residual=res, conv_taps=cw.taps_f, c_total=sum(chans) if len(xs) == 2 else cw.cin_pad, gn=gn)
        y, sums = y if gn is not None else (y, None)
    else:
        xs2d = ops.spatial(xs[0].v, "s2d")
        y = ops.gemm([xs2d], [cw.w_f], bias=cw.bias, conv_taps=cw.taps_f, gn=gn)
        y, sums = y if gn is not None else (y, None)
    out = Var(y)
    if sums is not None:
        out.gn = (out_gn, sums)
    if tape is not None:
        def bwd():
            dy = out.g
            if dy is None:
                return
            if residual is not None:
                _acc(residual, dy)
            if cw.k == 1:
                dy2 = dy.reshape(-1, cw.cout)
                off = 0
                for x, c in zip(xs, chans):
                    if x.needs_grad:
                        g = ops.gemm([dy2], [cw.w_d[off:off + c]], residual=x.g.reshape(-1, c) if x.g is not None else None)
                        x.g = g.reshape(x.v.shape)
                    off += c
            elif cw.stride == 1:
                cop = UW.pad_channels(cw.cout)
                dyp = dy if cop == cw.cout else torch.nn.functional.pad(dy, (0, cop - cw.cout))   # conv_out only (4 -> 64)
                off = 0
                for x, c in zip(xs, chans):
                    if x.needs_grad:
                        x.g = ops.gemm([dyp], [cw.w_d[off:off + c]], conv_taps=cw.taps_d, c_total=cop, residual=x.g)
                    off += c
            else:
                if xs[0].needs_grad:
                    g = ops.spatial(ops.gemm([dy], [cw.w_d], conv_taps=cw.taps_d), "d2s")
                    _acc(xs[0], g)
        tape.record(bwd)
    return out


def linear(tape, x: Var, lw: LinW, lora: Optional[LoRAW] = None, residual: Optional[Var] = None,
           out_gn: Optional[int] = None) -> Var:
    """y = x W^T (+ (x down^T) up^T) + b (+ residual); token-major (.., K) -> (.., N).
    ``out_gn`` (x is (n, L, K)): group count of the GroupNorm that consumes the result (statistics in the epilogue)."""
    xv = x.v
    x2 = xv.reshape(-1, lw.k)
    res = residual.v.reshape(-1, lw.n) if residual is not None else None
    gn = None if (out_gn is None or not FUSE_GN_STATS or xv.dim() != 3) else (out_gn, xv.shape[1])
    if lora is not None:
        t = ops.gemm([x2], [lora.down16])
        y = ops.gemm([x2, t], [lw.w, lora.up16], bias=lw.bias, residual=res, gn=gn)
    else:
        t = None
        y = ops.gemm([x2], [lw.w], bias=lw.bias, residual=res, gn=gn)
    y, sums = y if gn is not None else (y, None)
    out = Var(y.reshape(*xv.shape[:-1], lw.n))
    if sums is not None:
        out.gn = (out_gn, sums)
    if tape is not None:
        def bwd():
            dy = out.g
            if dy is None:
                return
            dy2 = dy.reshape(-1, lw.n)
            if residual is not None:
                _acc(residual, dy)
            if lora is not None:
                u = ops.gemm([dy2], [lora.up16_t])                                   # dy . up      (M, r)
                if lora.wgrad:
                    if lora.direct:
                        # straight into the optimiser's flat gradient buffer (param.grad views), loss scale undone by alpha:
                        # no per-parameter AccumulateGrad add, no unscale pass (training_script.py:659 accumulates the same sums)
                        ops.gemm_tn(dy2, t, accumulate_into=lora.up.grad, alpha=lora.inv_scale)      # d up   += dy^T t  (N, r)
                        ops.gemm_tn(u, x2, accumulate_into=lora.down.grad, alpha=lora.inv_scale)    # d down += u^T x   (r, K)
                    else:
                        lora.g_up = ops.gemm_tn(dy2, t, accumulate_into=lora.g_up)
                        lora.g_down = ops.gemm_tn(u, x2, accumulate_into=lora.g_down)
                if x.needs_grad:
                    g = ops.gemm([dy2, u], [lw.wt, lora.down16_t], residual=x.g.reshape(-1, lw.k) if x.g is not None else None)
                    x.g = g.reshape(xv.shape)
            elif x.needs_grad:
                g = ops.gemm([dy2], [lw.wt], residual=x.g.reshape(-1, lw.k) if x.g is not None else None)
                x.g = g.reshape(xv.shape)
        tape.record(bwd)
    return out


def groupnorm(tape, x: Var, nw: NormW, silu: bool) -> Var:
    if x.gn is not None and x.gn[0] == nw.groups:
        # statistics came with x from the epilogue of the GEMM that produced it: one pass (read x, write y)
        y, mr = ops.groupnorm_fwd_from_sums(x.v, x.gn[1], nw.gamma, nw.beta, nw.groups, nw.eps, silu)
    else:
        y, mr = ops.groupnorm_fwd(x.v, nw.gamma, nw.beta, nw.groups, nw.eps, silu)
    out = Var(y)
    if tape is not None:
        def bwd():
            if out.g is not None and x.needs_grad:
                _acc(x, ops.groupnorm_bwd(x.v, out.g, nw.gamma, nw.beta, mr, nw.groups, silu))
        tape.record(bwd)
    return out


def layernorm(tape, x: Var, nw: NormW) -> Var:
    y, mr = ops.layernorm_fwd(x.v, nw.gamma, nw.beta, nw.eps)
    out = Var(y)
    if tape is not None:
        def bwd():
            if out.g is not None and x.needs_grad:
                _acc(x, ops.layernorm_bwd(x.v, out.g, nw.gamma, mr))
        tape.record(bwd)
    return out


def geglu(tape, hg: Var) -> Var:
    out = Var(ops.geglu_fwd(hg.v))
    if tape is not None:
        def bwd():
            if out.g is not None:
                _acc(hg, ops.geglu_bwd(hg.v, out.g))
        tape.record(bwd)
    return out


def upsample2x(tape, x: Var) -> Var:
    out = Var(ops.spatial(x.v, "up2"))
    if tape is not None:
        def bwd():
            if out.g is not None and x.needs_grad:
                _acc(x, ops.spatial(out.g, "up2_bwd"))
        tape.record(bwd)
    return out


def concat(tape, a: Var, b: Var) -> Var:
    ca, cb = a.v.shape[-1], b.v.shape[-1]
    out = Var(ops.concat_channels(a.v, b.v))
    if tape is not None:
        def bwd():
            if out.g is None:
                return
            g = out.g
            if a.needs_grad:
                _acc(a, g[..., :ca].contiguous())
            if b.needs_grad:
                _acc(b, g[..., ca:].contiguous())
        tape.record(bwd)
    return out


def attention(tape, q: Var, k: Var, v: Var, heads: int, export_probs: bool = False, export_from: int = 0):
    """softmax(scale q k^T) v per head; optionally exports the fp32 probabilities (n*heads, Lq, Lk) as a Var whose
    gradient (from the attention-map loss) is added into the softmax backward (SURVEY 'hard parts')."""
    o, probs, saved = attn_ops.attention_fwd(q.v, k.v, v.v, heads, export_probs, need_bwd=tape is not None, export_from=export_from)
    out = Var(o)
    pvar = Var(probs) if export_probs else None
    if tape is not None:
        def bwd():
            if out.g is None and (pvar is None or pvar.g is None):
                return
            dq, dk, dv = attn_ops.attention_bwd(saved, out.g, pvar.g if pvar is not None else None)
            _acc(q, dq)
            _acc(k, dk)
            _acc(v, dv)
        tape.record(bwd)
    return out, pvar


# ------------------------------------------------------------------------------------------------------------
# network executors built from diffusers-shaped modules (state-dict compatible with oracle/sd_modules.py or diffusers)
# ------------------------------------------------------------------------------------------------------------
class _MergedLin:
    """LinW-shaped view of a projection with its LoRA folded in: w = W + up.down (N, K), wt = w^T (K, N) for dgrad."""
    __slots__ = ("w", "wt", "bias", "n", "k")

    def __init__(self, w, bias):
        self.w, self.wt, self.bias = w, None, bias
        self.n, self.k = w.shape


class _Attn:
    def __init__(self, m, dtype):
        self.heads = m.heads
        self.q, self.k, self.v, self.o = (LinW(l, dtype) for l in (m.to_q, m.to_k, m.to_v, m.to_out[0]))
        self.loras = [LoRAW(l.lora_layer, dtype) if getattr(l, "lora_layer", None) is not None else None
                      for l in (m.to_q, m.to_k, m.to_v, m.to_out[0])]
        self.gn = NormW(m.group_norm, m.group_norm.num_groups) if m.group_norm is not None else None
        self.mq = self.mk = self.mv = self.mo = None      # merged projections, built by build_merged()
        self.G = [None, None, None, None]                 # fp32 product-gradient accumulators dy^T x of q, k, v, o (UNetEngine arena)

    # ---- LoRA folded into the projection weights (passes that need no LoRA weight gradient)
    # y = W x + up(down(x)) (training_utils/pipeline.py:94-115) == (W + up.down) x: the no-grad rollout forwards and the
    # frozen-LoRA discriminator pass (gan_sdxl.py:52-89) run ONE plain GEMM per projection; q|k|v (self-attention) and k|v
    # (text context) live in one row-concatenated buffer so they are one GEMM each.
    def build_merged(self):
        dev, dt = self.q.w.device, self.q.w.dtype
        C = self.q.n
        if self.k.n != C or self.v.n != C:
            raise ValueError("attention projections must share the inner width")
        if self.k.k == self.q.k:                          # self-attention geometry: one (3C, K) buffer
            self.w_qkv = torch.empty(3 * C, self.q.k, dtype=dt, device=dev)
            wq, self.w_kv = self.w_qkv[:C], self.w_qkv[C:]
        else:
            self.w_qkv = None
            wq = torch.empty(C, self.q.k, dtype=dt, device=dev)
            self.w_kv = torch.empty(2 * C, self.k.k, dtype=dt, device=dev)
        wo = torch.empty(self.o.n, self.o.k, dtype=dt, device=dev)
        self.mq, self.mk, self.mv, self.mo = (_MergedLin(w, l.bias) for w, l in
                                              ((wq, self.q), (self.w_kv[:C], self.k), (self.w_kv[C:], self.v), (wo, self.o)))
        bs = [self.q.bias, self.k.bias, self.v.bias]
        cat = lambda xs, ls: None if all(b is None for b in xs) else torch.cat(
            [b if b is not None else torch.zeros(l.n, dtype=torch.float32, device=dev) for b, l in zip(xs, ls)])
        self.b_qkv = cat(bs, (self.q, self.k, self.v))
        self.b_kv = cat(bs[1:], (self.k, self.v))

    def refresh_merged(self, transposed: bool):
        if self.mq is None:
            self.build_merged()
        for ml, lw, lora in zip((self.mq, self.mk, self.mv, self.mo), (self.q, self.k, self.v, self.o), self.loras):
            if lora is None:
                ml.w.copy_(lw.w)
            else:
                ops.gemm([lora.up16], [lora.down16_t], residual=lw.w, out=ml.w)          # W + up.down, rounded once
            if transposed:
                if ml.wt is None:
                    ml.wt = torch.empty(lw.k, lw.n, dtype=lw.w.dtype, device=lw.w.device)
                if lora is None:
                    ml.wt.copy_(lw.wt)
                else:
                    ops.gemm([lora.down16_t], [lora.up16], residual=lw.wt, out=ml.wt)    # (W + up.down)^T


class _TBlock:
    def __init__(self, m, dtype):
        self.n1, self.n2, self.n3 = NormW(m.norm1), NormW(m.norm2), NormW(m.norm3)
        self.a1, self.a2 = _Attn(m.attn1, dtype), _Attn(m.attn2, dtype)
        self.ff1, self.ff2 = LinW(m.ff.net[0].proj, dtype), LinW(m.ff.net[2], dtype)
        # no-grad passes: GEGLU fused into the projection's epilogue - weight rows interleaved (2j = hidden_j, 2j+1 = gate_j) so an
        # accumulator column pair is one output of ``hidden * gelu(gate)`` (diffusers GEGLU.forward: proj(x).chunk(2, -1))
        w, b = m.ff.net[0].proj.weight.detach(), m.ff.net[0].proj.bias.detach()
        inner = w.shape[0] // 2
        self.ff1_il_w = UW.to16(torch.stack([w[:inner], w[inner:]], 1).reshape(2 * inner, -1), dtype)
        self.ff1_il_b = torch.stack([b[:inner], b[inner:]], 1).reshape(-1).float().contiguous()


class _Transformer:
    def __init__(self, m, dtype):
        self.gn = NormW(m.norm, m.norm.num_groups)
        self.linear_proj = m.use_linear_projection
        if self.linear_proj:
            self.pin, self.pout = LinW(m.proj_in, dtype), LinW(m.proj_out, dtype)
        else:
            self.pin, self.pout = ConvW(m.proj_in, dtype), ConvW(m.proj_out, dtype)
        self.blocks = [_TBlock(b, dtype) for b in m.transformer_blocks]


class _Res:
    def __init__(self, m, dtype):
        self.n1, self.n2 = NormW(m.norm1, m.norm1.num_groups), NormW(m.norm2, m.norm2.num_groups)
        self.c1, self.c2 = ConvW(m.conv1, dtype), ConvW(m.conv2, dtype)
        self.temb = LinW(m.time_emb_proj, dtype) if m.time_emb_proj is not None else None
        self.short = ConvW(m.conv_shortcut, dtype) if m.conv_shortcut is not None else None


def _resnet(tape, r: _Res, x: Var, temb_act16, skip: Optional[Var] = None, out_gn: Optional[int] = None) -> Var:
    """ResnetBlock2D.  ``skip``: the UNet skip tensor — GroupNorm spans the concatenation so it is
    materialised once; the 1x1 shortcut reads the two segments directly.
    GroupNorm statistics: norm2's come out of conv1's epilogue, norm1's out of whatever GEMM produced ``x`` (``x.gn``; not
    across a concatenation), and ``out_gn`` asks conv2 for the statistics of the GroupNorm that consumes this block's output."""
    xin = concat(tape, x, skip) if skip is not None else x
    h = groupnorm(tape, xin, r.n1, True)
    rowvec = None
    if r.temb is not None and temb_act16 is not None:
        if isinstance(temb_act16, dict):                                                 # all projections done by one GEMM
            rowvec = temb_act16[id(r)]
        else:
            rowvec = ops.gemm([temb_act16], [r.temb.w], bias=r.temb.bias, out_fp32=True)  # (n, Cout) fp32, constant wrt params
    h = conv(tape, [h], r.c1, rowvec=rowvec, out_gn=r.n2.groups)
    h = groupnorm(tape, h, r.n2, True)
    sc = conv(tape, [xin], r.short) if r.short is not None else xin
    return conv(tape, [h], r.c2, residual=sc, out_gn=out_gn)


def _attn_layer(tape, a: _Attn, x: Var, ctx: Optional[Var], residual: Var, capture=None, place=None, mode="train",
                cross_kv=None):
    """mode 'train': LoRA branch explicit (weight gradients wanted).  'frozen': taped, LoRA folded into the weights (data
    gradients only).  'merged': no tape - folded weights, fused q|k|v / k|v GEMMs, ``cross_kv`` = the text context's
    pre-projected (n, T, 2C) k|v (identical for every UNet call of a rollout, so it is computed once per optimiser step)."""
    if mode == "product" and tape is not None:
        return _attn_layer_product(tape, a, x, ctx, residual, capture, place, cross_kv)
    export = capture is not None and ctx is not None and capture.wants(place)
    if mode == "merged" and tape is None:
        C = a.q.n
        x2 = x.v.reshape(-1, a.q.k)
        lead = x.v.shape[:-1]
        if ctx is None and a.w_qkv is not None:
            qkv = ops.gemm([x2], [a.w_qkv], bias=a.b_qkv).reshape(*lead, 3 * C)
            q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        else:
            q = ops.gemm([x2], [a.mq.w], bias=a.mq.bias).reshape(*lead, C)
            if cross_kv is not None and ctx is not None:
                kv = cross_kv
            else:
                src = x.v if ctx is None else ctx.v
                kv = ops.gemm([src.reshape(-1, a.k.k)], [a.w_kv], bias=a.b_kv).reshape(*src.shape[:-1], 2 * C)
            k, v = kv[..., :C], kv[..., C:]
        o, probs, _ = attn_ops.attention_fwd(q, k, v, a.heads, export, need_bwd=False, export_from=capture.sample_from if export else 0)
        if capture is not None:
            capture.push(Var(probs) if export else None, ctx is not None, place)
        y = ops.gemm([o.reshape(-1, C)], [a.mo.w], bias=a.mo.bias, residual=residual.v.reshape(-1, a.mo.n))
        return Var(y.reshape(*lead, a.mo.n))
    if mode == "train":
        lq, lk, lv, lo, loras = a.q, a.k, a.v, a.o, a.loras
    else:
        lq, lk, lv, lo, loras = a.mq, a.mk, a.mv, a.mo, (None, None, None, None)
    q = linear(tape, x, lq, loras[0])
    src = x if ctx is None else ctx
    k = linear(tape, src, lk, loras[1])
    v = linear(tape, src, lv, loras[2])
    o, p = attention(tape, q, k, v, a.heads, export_probs=export, export_from=capture.sample_from if export else 0)
    if capture is not None:
        capture.push(p, ctx is not None, place)
    return linear(tape, o, lo, loras[3], residual=residual)


def _attn_layer_product(tape, a: _Attn, x: Var, ctx: Optional[Var], residual: Var, capture=None, place=None, cross_kv=None):
    """Taped attention layer with LoRA weight gradients, 'product' form.  Forward runs on the folded weights exactly like the
    no-grad path (one q|k|v GEMM, one k|v GEMM, strided attention operands).  Backward: data gradients through the folded
    weights, and for every LoRA'd projection ONE accumulation  G += dy^T x  (fp32, full weight shape) instead of the
    explicit branch's  t = x down^T,  u = dy up,  d up = dy^T t,  d down = u^T x : since
    d up = (dy^T x) down^T  and  d down = up^T (dy^T x)  are linear in G and the factors are constant between optimiser steps,
    G is accumulated over all back-propagated sampler steps and projected once per optimiser step
    (UNetEngine.finalize_lora_grads).  Same gradients as training_utils/pipeline.py:94-115's autograd."""
    C = a.q.n
    xv = x.v
    lead = xv.shape[:-1]
    x2 = xv.reshape(-1, a.q.k)
    export = capture is not None and ctx is not None and capture.wants(place)
    if ctx is None and a.w_qkv is not None:
        qkv = ops.gemm([x2], [a.w_qkv], bias=a.b_qkv).reshape(*lead, 3 * C)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        src2 = x2
    else:
        q = ops.gemm([x2], [a.mq.w], bias=a.mq.bias).reshape(*lead, C)
        src = xv if ctx is None else ctx.v
        src2 = src.reshape(-1, a.k.k)
        if cross_kv is not None and ctx is not None:
            kv = cross_kv
        else:
            kv = ops.gemm([src2], [a.w_kv], bias=a.b_kv).reshape(*src.shape[:-1], 2 * C)
        k, v = kv[..., :C], kv[..., C:]
    o, probs, saved = attn_ops.attention_fwd(q, k, v, a.heads, export, need_bwd=True, export_from=capture.sample_from if export else 0)
    pvar = Var(probs) if export else None
    if capture is not None:
        capture.push(pvar, ctx is not None, place)
    o2 = o.reshape(-1, C)
    y = ops.gemm([o2], [a.mo.w], bias=a.mo.bias, residual=residual.v.reshape(-1, a.mo.n))
    out = Var(y.reshape(*lead, a.mo.n))

    def wants(i):
        return a.loras[i] is not None and a.loras[i].wgrad

    def bwd():
        dy = out.g
        if dy is None and (pvar is None or pvar.g is None):
            return
        do = None
        if dy is not None:
            dy2 = dy.reshape(-1, a.mo.n)
            _acc(residual, dy)
            if wants(3):
                ops.gemm_tn(dy2, o2, accumulate_into=a.G[3])
            do = ops.gemm([dy2], [a.mo.wt]).reshape(o.shape)
        dq, dk, dv = attn_ops.attention_bwd(saved, do, pvar.g if pvar is not None else None)
        dq2, dk2, dv2 = dq.reshape(-1, C), dk.reshape(-1, C), dv.reshape(-1, C)
        if wants(0):
            ops.gemm_tn(dq2, x2, accumulate_into=a.G[0])
        if wants(1):
            ops.gemm_tn(dk2, src2, accumulate_into=a.G[1])
        if wants(2):
            ops.gemm_tn(dv2, src2, accumulate_into=a.G[2])
        if x.needs_grad:
            g = ops.gemm([dq2], [a.mq.wt], residual=x.g.reshape(-1, a.q.k) if x.g is not None else None)
            if ctx is None:
                g = ops.gemm([dk2], [a.mk.wt], residual=g)
                g = ops.gemm([dv2], [a.mv.wt], residual=g)
            x.g = g.reshape(xv.shape)
        if ctx is not None and ctx.needs_grad:            # d(text context) through the folded k / v projections (off by default)
            g = ops.gemm([dk2], [a.mk.wt], residual=ctx.g.reshape(-1, a.k.k) if ctx.g is not None else None)
            g = ops.gemm([dv2], [a.mv.wt], residual=g)
            ctx.g = g.reshape(ctx.v.shape)
    tape.record(bwd)
    return out


def _transformer(tape, t: _Transformer, x: Var, ctx: Var, capture=None, place=None, mode="train", cross_kv=None,
                 out_gn: Optional[int] = None) -> Var:
    n, H, W, C = x.v.shape
    h = groupnorm(tape, x, t.gn, False)
    if t.linear_proj:
        h = linear(tape, _reshape(tape, h, (n, H * W, C)), t.pin)
    else:
        h = _reshape(tape, conv(tape, [h], t.pin), (n, H * W, t.pin.cout))
    for b in t.blocks:
        h = _attn_layer(tape, b.a1, layernorm(tape, h, b.n1), None, h, capture, place, mode)
        h = _attn_layer(tape, b.a2, layernorm(tape, h, b.n2), ctx, h, capture, place, mode,
                        None if cross_kv is None else cross_kv.get(id(b.a2)))
        if tape is None:
            y3 = layernorm(None, h, b.n3).v
            ff = Var(ops.gemm([y3.reshape(-1, y3.shape[-1])], [b.ff1_il_w], bias=b.ff1_il_b, act="geglu").reshape(*y3.shape[:-1], -1), False)
        else:
            ff = geglu(tape, linear(tape, layernorm(tape, h, b.n3), b.ff1))
        h = linear(tape, ff, b.ff2, residual=h)
    if t.linear_proj:
        return _reshape(tape, linear(tape, h, t.pout, residual=_reshape(tape, x, (n, H * W, C)), out_gn=out_gn), (n, H, W, C))
    return conv(tape, [_reshape(tape, h, (n, H, W, h.v.shape[-1]))], t.pout, residual=x, out_gn=out_gn)


def _reshape(tape, x: Var, shape) -> Var:
    out = Var(x.v.reshape(shape), x.needs_grad)
    out.gn = x.gn                                   # per-(image, group) statistics do not depend on the view
    if tape is not None:
        def bwd():
            if out.g is not None:
                _acc(x, out.g.reshape(x.v.shape))
        tape.record(bwd)
    return out


class AttnCapture:
    """Product-side AttentionStore (attn_utils/tc_attn_utils.py:53-94): keeps the exported cross-attention
    probability Vars of the requested places for one UNet pass."""

    def __init__(self, train_layer_ls):
        self.places = sorted({s.split("_")[0] for s in train_layer_ls})
        # first sample of the batch whose maps are captured: the attrcon step stores the CONDITIONAL half of the CFG batch
        # (AttrConcenTrainableSDPipeline.py:239-279), so one full-batch UNet call with sample_from = n/2 replaces the two half calls
        self.sample_from = 0
        self.reset()

    def reset(self):
        self.store = {"down": [], "mid": [], "up": []}
        self.count = 0

    def wants(self, place):
        return place in self.places

    def push(self, pvar, is_cross, place):
        self.count += 1
        if pvar is not None and is_cross and place in self.places:
            self.store[place].append(pvar)

    def attn_dict(self, reses=(64, 32, 16, 8), poses=("down", "mid", "up")):
        """get_cross_attn_map_from_unet (tc_attn_utils.py:198-216) -> ({'up_16': [tensor (B*h,res,res,T)]}, same of Vars)"""
        out, vars_ = {}, {}
        for pos in poses:
            for res in reses:
                sel = [p for p in self.store[pos] if p.v.shape[1] == res * res]
                if sel:
                    out[f"{pos}_{res}"] = [p.v.reshape(-1, res, res, p.v.shape[-1]) for p in sel]
                    vars_[f"{pos}_{res}"] = sel
        return out, vars_


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers Timesteps(flip_sin_to_cos=True, freq_shift=0): [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    arg = t.float()[:, None] * freqs[None]
    return torch.cat([arg.cos(), arg.sin()], -1)


class UNetEngine:
    """UNet2DConditionModel executor (SD1.5 and SDXL geometries)."""

    def __init__(self, unet: torch.nn.Module, dtype=torch.float16):
        self.dtype = dtype
        self.cfg = unet.config
        self.te1, self.te2 = LinW(unet.time_embedding.linear_1, dtype), LinW(unet.time_embedding.linear_2, dtype)
        self.tdim = unet.time_embedding.linear_1.in_features
        self.sdxl = getattr(unet.config, "addition_embed_type", None) == "text_time"
        if self.sdxl:
            self.ae1, self.ae2 = LinW(unet.add_embedding.linear_1, dtype), LinW(unet.add_embedding.linear_2, dtype)
            self.add_dim = unet.config.addition_time_embed_dim
        self.conv_in = ConvW(unet.conv_in, dtype, cin_pad=64)
        self.down = []
        for blk in unet.down_blocks:
            attns = [_Transformer(a, dtype) for a in blk.attentions] if hasattr(blk, "attentions") else None
            ds = ConvW(blk.downsamplers[0].conv, dtype) if blk.downsamplers is not None else None
            self.down.append(([_Res(r, dtype) for r in blk.resnets], attns, ds))
        mb = unet.mid_block
        self.mid = ([_Res(r, dtype) for r in mb.resnets], [_Transformer(a, dtype) for a in mb.attentions])
        self.up = []
        for blk in unet.up_blocks:
            attns = [_Transformer(a, dtype) for a in blk.attentions] if hasattr(blk, "attentions") else None
            us = ConvW(blk.upsamplers[0].conv, dtype) if blk.upsamplers is not None else None
            self.up.append(([_Res(r, dtype) for r in blk.resnets], attns, us))
        self.norm_out = NormW(unet.conv_norm_out, unet.conv_norm_out.num_groups)
        self.conv_out = ConvW(unet.conv_out, dtype)
        # every ResBlock's time_emb_proj applied by ONE GEMM per UNet call (22 tiny launches -> 1): concatenated weights
        self._res_all = [r for rs, _, _ in self.down for r in rs] + list(self.mid[0]) + [r for rs, _, _ in self.up for r in rs]
        self._temb_w = torch.cat([r.temb.w for r in self._res_all], 0).contiguous()
        self._temb_b = torch.cat([r.temb.bias for r in self._res_all], 0).contiguous()
        self.loras: List[LoRAW] = []
        self._attns: List[_Attn] = []
        self._cross_attns: List[_Attn] = []
        for blocks in [self.down, [(self.mid[0], self.mid[1], None)], self.up]:
            for _, attns, _ in blocks:
                for t in attns or []:
                    for b in t.blocks:
                        self._cross_attns.append(b.a2)
                        for a in (b.a1, b.a2):
                            self._attns.append(a)
                            self.loras.extend(l for l in a.loras if l is not None)
        self.lora_version, self._merged_version, self._merged_t = 0, -1, False
        self._arenas = None
        self.refresh_lora()
        # 'product' (default): LoRA weight gradients through the accumulated full-size product G = dy^T x, projected onto the
        # factors once per optimiser step; 'explicit': the literal LoRA branch (t = x down^T ...) in forward and backward
        self.lora_train_impl = "product"
        # Folding (W' = round16(W16 + up16.down16)) quantises the adapter's contribution to the ulp of W: harmless in fp16 (ulp 2^-11 |W|)
        # once the adapter has grown for a few steps, but in bf16 (ulp 2^-8 |W| ~ 1.2e-4 at |W| ~ 0.03) a young adapter is rounded away
        # entirely and the loss does not see it.  bf16 engines therefore run the literal LoRA branch y = W x + up(down(x)) in every
        # pass (the reference's arithmetic, training_utils/pipeline.py:94-115) - slower, exact in the adapter.
        self.fold_lora = dtype != torch.bfloat16
        if not self.fold_lora:
            self.lora_train_impl = "explicit"
        self._G = self._G_hi = self._G_lo = None
        self._gproj = []
        self.G_dirty = False

    # ---- LoRA plumbing
    def refresh_lora(self):
        if self.loras:
            if self._arenas is None:
                self._arenas = _LoRAArenas(self.loras, self.dtype)
            self._arenas.refresh()
        self.lora_version += 1                 # folded weights / cached context projections are stale now

    def ensure_merged(self, transposed: bool = False):
        """(re)build the LoRA-folded projection weights if the LoRA parameters changed since the last build.  Runs eagerly
        (never inside a CUDA-graph capture: callers invoke it before capturing / replaying)."""
        if self._merged_version != self.lora_version:
            for a in self._attns:
                a.refresh_merged(transposed or self._merged_t)
            self._merged_version, self._merged_t = self.lora_version, transposed or self._merged_t
        elif transposed and not self._merged_t:
            for a in self._attns:
                a.refresh_merged(True)
            self._merged_t = True

    def _ensure_G(self):
        """fp32 arena of the product-gradient accumulators (same shapes as the LoRA'd projection weights) + its bf16 hi/lo images."""
        if self._G is not None:
            return
        shapes = []
        for a in self._attns:
            for i, lw in enumerate((a.q, a.k, a.v, a.o)):
                if a.loras[i] is not None:
                    shapes.append((a, i, lw.n, lw.k))
        total = sum(n * k for _, _, n, k in shapes)
        dev = self.te1.w.device
        self._G = torch.zeros(max(total, 4), dtype=torch.float32, device=dev)
        self._G_hi = torch.empty(max(total, 4), dtype=torch.bfloat16, device=dev)
        self._G_lo = torch.empty(max(total, 4), dtype=torch.bfloat16, device=dev)
        off = 0
        for a, i, n, k in shapes:
            a.G[i] = self._G[off:off + n * k].view(n, k)
            self._gproj.append((off, n, k, a.loras[i]))
            off += n * k

    def finalize_lora_grads(self, inv_scale: float = 1.0, into_param_grads: bool = True):
        """project the accumulated products onto the LoRA factors:  d up = G down^T (N, r),  d down = up^T G (r, K), G fed to
        the tensor cores as bf16 hi + lo (two K segments / two accumulating passes: ~16 mantissa bits at fp32 range).
        ``into_param_grads``: accumulate into ``param.grad`` (the optimiser's flat buffer) and return None; otherwise return
        the gradients in ``lora_params()`` order.  Clears the accumulators."""
        if self._G is None or not self.G_dirty:
            return None if into_param_grads else [None] * (2 * len(self.loras))
        ops.split_f32_bf16x2(self._G, self._G_hi, self._G_lo, inv_scale)
        res = {}
        for off, n, k, lora in self._gproj:
            hi, lo = self._G_hi[off:off + n * k].view(n, k), self._G_lo[off:off + n * k].view(n, k)
            # (gh + gl)(dh + dl)^T without the gl.dl term (2^-16 relative), likewise for up^T G
            if into_param_grads:
                g_up, g_down = lora.up.grad, lora.down.grad
            else:
                g_up = torch.zeros(n, lora.rank, dtype=torch.float32, device=hi.device)
                g_down = torch.zeros(lora.rank, k, dtype=torch.float32, device=hi.device)
                res[id(lora)] = (g_down, g_up)
            ops.gemm([hi, lo], [lora.down_bh, lora.down_bh], out=g_up, accumulate=True)
            ops.gemm([hi], [lora.down_bl], out=g_up, accumulate=True)
            ops.gemm_tn(lora.up_bh, hi, accumulate_into=g_down)
            ops.gemm_tn(lora.up_bh, lo, accumulate_into=g_down)
            ops.gemm_tn(lora.up_bl, hi, accumulate_into=g_down)
        self._G.zero_()
        self.G_dirty = False
        if into_param_grads:
            return None
        return [g for l in self.loras for g in res.get(id(l), (None, None))]

    def cross_kv(self, ctx16: torch.Tensor, out=None):
        """k|v projections of the text context for every cross-attention layer: {id(layer): (n, T, 2C)}.  The context and
        the (folded) weights are the same for all UNet calls between two optimiser steps, so callers cache the result."""
        self.ensure_merged()
        n, T, D = ctx16.shape
        c2 = ctx16.reshape(n * T, D)
        res = {}
        for a in self._cross_attns:
            o = None if out is None else out[id(a)].reshape(n * T, 2 * a.q.n)
            res[id(a)] = ops.gemm([c2], [a.w_kv], bias=a.b_kv, out=o).reshape(n, T, 2 * a.q.n)
        return res

    def set_lora_wgrad(self, flag: bool):
        for l in self.loras:
            l.wgrad = flag

    def zero_lora_grads(self):
        for l in self.loras:
            l.g_down = l.g_up = None

    def lora_params(self):
        return [p for l in self.loras for p in (l.down, l.up)]

    def lora_grads(self):
        return [g for l in self.loras for g in (l.g_down, l.g_up)]

    # ---- forward
    def temb(self, t: torch.Tensor, n: int, added_cond=None):
        """SiLU(time embedding) as 16-bit (n, 1280) — constant w.r.t. all trainable tensors."""
        te = timestep_embedding(t.reshape(-1).expand(n).to(self.te1.w.device), self.tdim).to(self.dtype)
        e = ops.gemm([ops.gemm([te], [self.te1.w], bias=self.te1.bias, act="silu")], [self.te2.w], bias=self.te2.bias, out_fp32=True)
        if self.sdxl:
            tid = added_cond["time_ids"].reshape(-1)
            te2 = timestep_embedding(tid, self.add_dim).reshape(n, -1)
            add = torch.cat([added_cond["text_embeds"].float(), te2], -1).to(self.dtype)
            e = e + ops.gemm([ops.gemm([add], [self.ae1.w], bias=self.ae1.bias, act="silu")], [self.ae2.w], bias=self.ae2.bias, out_fp32=True)
        return ops.elementwise("silu", e.to(self.dtype))

    def forward(self, tape: Optional[Tape], x: Var, t: torch.Tensor, ctx: torch.Tensor, capture: Optional[AttnCapture] = None,
                added_cond=None, lora_mode: Optional[str] = None, cross_kv=None) -> Var:
        """x: Var of NHWC 16-bit latents zero-padded to 64 channels (n, h, w, 64); ctx: (n, 77, D) 16-bit.
        ``lora_mode``: 'train' (explicit LoRA branch, weight gradients), 'frozen' (taped, LoRA folded into the weights) or
        'merged' (no tape, folded + fused projections; default without a tape).  ``cross_kv``: cached ``self.cross_kv(ctx)``.
        returns Var (n, h, w, 4)."""
        n = x.v.shape[0]
        mode = lora_mode or ("train" if tape is not None else "merged")
        if mode == "merged" and tape is not None:
            mode = "frozen"
        if not self.fold_lora and self.loras and self.lora_train_impl != "product":
            mode, cross_kv = "train", None         # bf16: explicit LoRA branch in every pass (wgrad switched per pass by set_lora_wgrad)
        if mode == "train" and self.lora_train_impl == "product":
            mode = "product"
            self._ensure_G()
            tape.record(lambda: setattr(self, "G_dirty", True))       # runs last in the backward: accumulators hold this pass
        if mode != "train":
            self.ensure_merged(transposed=tape is not None)
        temb16 = self.temb(t, n, added_cond)
        allp = ops.gemm([temb16], [self._temb_w], bias=self._temb_b, out_fp32=True)       # (n, sum Cout) fp32
        temb, off = {}, 0
        for r in self._res_all:
            temb[id(r)] = allp[:, off:off + r.temb.n]
            off += r.temb.n
        # the text context is a constant of the CoMat step (frozen text encoders); a caller that wants d(ctx) passes a Var
        cvar = ctx if isinstance(ctx, Var) else Var(ctx, needs_grad=False)
        # GroupNorm statistics ride in the epilogue of the GEMM that produces the normalised tensor: every producer is told the
        # group count of the GroupNorm that reads its output (None: the consumer normalises a concatenation or is not a GroupNorm)
        ops.gn_arena_reset(x.v.device)
        mid_gn = self.mid[0][0].n1.groups
        h = conv(tape, [x], self.conv_in, out_gn=self.down[0][0][0].n1.groups)
        skips = [h]
        for bi, (resnets, attns, ds) in enumerate(self.down):
            for i, r in enumerate(resnets):
                nxt = resnets[i + 1].n1.groups if i + 1 < len(resnets) else (None if ds is not None else mid_gn)
                h = _resnet(tape, r, h, temb, out_gn=attns[i].gn.groups if attns is not None else nxt)
                if attns is not None:
                    h = _transformer(tape, attns[i], h, cvar, capture, "down", mode, cross_kv, out_gn=nxt)
                skips.append(h)
            if ds is not None:
                h = conv(tape, [h], ds, out_gn=self.down[bi + 1][0][0].n1.groups if bi + 1 < len(self.down) else mid_gn)
                skips.append(h)
        h = _resnet(tape, self.mid[0][0], h, temb, out_gn=self.mid[1][0].gn.groups)
        h = _transformer(tape, self.mid[1][0], h, cvar, capture, "mid", mode, cross_kv, out_gn=self.mid[0][1].n1.groups)
        h = _resnet(tape, self.mid[0][1], h, temb)                    # consumed through a concatenation with the skip tensor
        n_up = len(self.up)
        for bi, (resnets, attns, us) in enumerate(self.up):
            for i, r in enumerate(resnets):
                # only the very last block output feeds a GroupNorm directly (conv_norm_out); the others are concatenated first
                last = bi == n_up - 1 and i == len(resnets) - 1 and us is None
                fin = self.norm_out.groups if last else None
                h = _resnet(tape, r, h, temb, skip=skips.pop(), out_gn=(attns[i].gn.groups if attns is not None else fin))
                if attns is not None:
                    h = _transformer(tape, attns[i], h, cvar, capture, "up", mode, cross_kv, out_gn=fin)
            if us is not None:
                h = conv(tape, [upsample2x(tape, h)], us)
        h = groupnorm(tape, h, self.norm_out, True)
        return conv(tape, [h], self.conv_out)


class VAEDecoderEngine:
    """AutoencoderKL.decode executor (TrainableSDPipeline.py:220)."""

    def __init__(self, vae: torch.nn.Module, dtype=torch.float16):
        self.dtype = dtype
        self.scaling_factor = vae.config.scaling_factor
        self.pq = ConvW(vae.post_quant_conv, dtype, cin_pad=64)
        d = vae.decoder
        self.conv_in = ConvW(d.conv_in, dtype, cin_pad=64)
        self.mid_res = [_Res(r, dtype) for r in d.mid_block.resnets]
        self.mid_attn = _Attn(d.mid_block.attentions[0], dtype)
        self.ups = []
        for blk in d.up_blocks:
            us = ConvW(blk.upsamplers[0].conv, dtype) if blk.upsamplers is not None else None
            self.ups.append(([_Res(r, dtype) for r in blk.resnets], us))
        self.norm_out = NormW(d.conv_norm_out, d.conv_norm_out.num_groups)
        self.conv_out = ConvW(d.conv_out, dtype)

    def forward(self, tape: Optional[Tape], z: Var) -> Var:
        """z: Var NHWC 16-bit (n,h,w,64) = latents / scaling_factor zero-padded; returns (n, 8h, 8w, 3)."""
        n, H, W, _ = z.v.shape
        # post_quant_conv is 1x1 4->4: run on the padded tensor, then re-pad its 4 outputs to 64 for conv_in
        ops.gn_arena_reset(z.v.device)
        pq = conv(tape, [z], self.pq_as_padded())
        h = conv(tape, [pq], self.conv_in, out_gn=self.mid_res[0].n1.groups)
        a = self.mid_attn
        h = _resnet(tape, self.mid_res[0], h, None, out_gn=a.gn.groups)
        hn = _reshape(tape, groupnorm(tape, h, a.gn, False), (n, H * W, h.v.shape[-1]))
        hr = _reshape(tape, h, (n, H * W, h.v.shape[-1]))
        h = _reshape(tape, _attn_layer(tape, a, hn, None, hr), (n, H, W, h.v.shape[-1]))
        # every later GroupNorm reads the output of the conv right before it (no concatenations in the decoder)
        chain = [r for resnets, _ in self.ups for r in resnets]
        nxt_of = {id(r): (chain[i + 1].n1.groups if i + 1 < len(chain) else self.norm_out.groups) for i, r in enumerate(chain)}
        h = _resnet(tape, self.mid_res[1], h, None, out_gn=chain[0].n1.groups)
        for resnets, us in self.ups:
            for j, r in enumerate(resnets):
                h = _resnet(tape, r, h, None, out_gn=None if (us is not None and j == len(resnets) - 1) else nxt_of[id(r)])
            if us is not None:
                h = conv(tape, [upsample2x(tape, h)], us, out_gn=nxt_of[id(resnets[-1])])
        h = groupnorm(tape, h, self.norm_out, True)
        return conv(tape, [h], self.conv_out)

    def pq_as_padded(self) -> ConvW:
        """post_quant_conv re-expressed as a 64->64 1x1 conv (zero rows/cols) so its output feeds conv_in directly."""
        if not hasattr(self, "_pq64"):
            cw = object.__new__(ConvW)
            w = torch.zeros(64, 64, dtype=self.dtype, device=self.pq.w_f.device)
            w[: self.pq.cout, : self.pq.cin] = self.pq.w_f[:, : self.pq.cin]
            cw.cout, cw.cin, cw.k, cw.stride, cw.cin_pad = 64, 64, 1, 1, 64
            cw.bias = torch.zeros(64, dtype=torch.float32, device=w.device)
            cw.bias[: self.pq.cout] = self.pq.bias
            cw.w_f, cw.taps_f, cw.w_d, cw.taps_d = w.contiguous(), None, w.t().contiguous(), None
            self._pq64 = cw
        return self._pq64
