"""Native executor for the BLIP captioner on the concept-matching path (reached from
concept_mat_utils/caption_blip.py:56-58 -> HF BlipForConditionalGeneration.forward).

ViT-L/16 vision tower (patch-embed as a GEMM over non-overlapping patches, 24 pre-LN blocks, post-LN) and the BERT-style text
decoder (causal self-attention, cross-attention over the 577 image tokens, post-LN FFN, LM head, label-smoothed CE) run on the
same tcgen05 GEMM / attention kernels and HBM-bound LayerNorm / GELU kernels as the UNet.  All weights are frozen
(caption_blip.py:20-21): only the data gradient w.r.t. the pixel values is produced (explicit tape, no autograd inside).

The parameter owner is an HF ``BlipForConditionalGeneration`` (real checkpoint or random init); weights are packed once.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import attention as attn_ops
from . import engine as E
from . import ops

_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
_lib.register_signature("comat_ce_label_smooth_fwd", [_vp, _vp, _vp, _vp, _i, _i, _ll, _f, _ll, _vp])
_lib.register_signature("comat_ce_label_smooth_bwd", [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll, _f, _ll, _i, _vp])


def ce_fwd(logits, labels, V, eps):
    """label-smoothed CE over fp32 logits (R, ld): returns (row_stats (R,2), out2 = {mean loss, #valid})."""
    R = logits.shape[0]
    stats = torch.empty(R, 2, dtype=torch.float32, device=logits.device)
    out2 = torch.empty(2, dtype=torch.float32, device=logits.device)
    _lib.check(_lib.lib().comat_ce_label_smooth_fwd(logits.data_ptr(), labels.data_ptr(), stats.data_ptr(), out2.data_ptr(), R, V,
                                                    logits.stride(0), eps, -100, _lib.stream_ptr()), "ce_fwd")
    _lib.count_launch(2)
    return stats, out2


def ce_bwd(logits, labels, stats, out2, gout, V, Vpad, eps, dtype):
    R = logits.shape[0]
    d = torch.empty(R, Vpad, dtype=dtype, device=logits.device)
    _lib.check(_lib.lib().comat_ce_label_smooth_bwd(logits.data_ptr(), labels.data_ptr(), stats.data_ptr(), out2.data_ptr(),
                                                    gout.float().contiguous().data_ptr(), d.data_ptr(), R, V, Vpad, logits.stride(0), eps,
                                                    -100, ops.DT[dtype], _lib.stream_ptr()), "ce_bwd")
    _lib.count_launch()
    return d


def _gelu(tape, x: E.Var) -> E.Var:
    out = E.Var(ops.elementwise("gelu", x.v))
    if tape is not None:
        def bwd():
            if out.g is not None:
                E._acc(x, ops.elementwise("gelu_bwd", x.v, out.g))
        tape.record(bwd)
    return out


def _add_const(tape, x: E.Var, c: torch.Tensor) -> E.Var:
    out = E.Var(ops.elementwise("add", x.v, c))
    if tape is not None:
        def bwd():
            if out.g is not None:
                E._acc(x, out.g)
        tape.record(bwd)
    return out


def _mha(tape, q: E.Var, k: E.Var, v: E.Var, heads, kv_lens=None, causal=False) -> E.Var:
    o, _, lse = attn_ops.attention_fwd_native(q.v, k.v, v.v, heads, need_lse=tape is not None, kv_lens=kv_lens, causal=causal)
    out = E.Var(o)
    if tape is not None:
        qv, kv, vv = q.v.contiguous(), k.v.contiguous(), v.v.contiguous()

        def bwd():
            if out.g is None:
                return
            dq, dk, dv = attn_ops.attention_bwd_native(qv, kv, vv, o, lse, None, heads, out.g, None, kv_lens=kv_lens, causal=causal)
            E._acc(q, dq); E._acc(k, dk); E._acc(v, dv)
        tape.record(bwd)
    return out


class _Lin(E.LinW):
    pass


def _split_qkv(lin: torch.nn.Linear, dtype):
    """HF BlipAttention fuses q,k,v in one Linear(dim, 3*dim); split so each projection is a contiguous (n, L, H*d) GEMM output."""
    w, b = lin.weight.detach(), lin.bias.detach() if lin.bias is not None else None
    dim = w.shape[1]
    outs = []
    for i in range(3):
        l = torch.nn.Linear(dim, dim, bias=b is not None, device=w.device, dtype=w.dtype)
        l.weight.data.copy_(w[i * dim:(i + 1) * dim])
        if b is not None:
            l.bias.data.copy_(b[i * dim:(i + 1) * dim])
        outs.append(E.LinW(l, dtype))
    return outs


class BlipEngine:
    def __init__(self, model, dtype=torch.float16):
        self.dtype = dtype
        vm = model.vision_model
        cfg_v, cfg_t = model.config.vision_config, model.config.text_config
        self.heads_v, self.heads_t = cfg_v.num_attention_heads, cfg_t.num_attention_heads
        self.patch = cfg_v.patch_size
        emb = vm.embeddings
        pw = emb.patch_embedding.weight.detach()                         # (dim, 3, p, p)
        lin = torch.nn.Linear(pw[0].numel(), pw.shape[0], device=pw.device, dtype=pw.dtype)
        lin.weight.data.copy_(pw.reshape(pw.shape[0], -1))
        lin.bias.data.copy_(emb.patch_embedding.bias.detach())
        self.patch_lin = E.LinW(lin, dtype)
        self.cls = emb.class_embedding.detach().to(dtype)                # (1,1,dim)
        self.pos = emb.position_embedding.detach().to(dtype)             # (1,577,dim)
        self.vlayers = []
        for lyr in vm.encoder.layers:
            q, k, v = _split_qkv(lyr.self_attn.qkv, dtype)
            self.vlayers.append(dict(n1=E.NormW(lyr.layer_norm1), q=q, k=k, v=v, proj=E.LinW(lyr.self_attn.projection, dtype),
                                     n2=E.NormW(lyr.layer_norm2), fc1=E.LinW(lyr.mlp.fc1, dtype), fc2=E.LinW(lyr.mlp.fc2, dtype)))
        self.post_ln = E.NormW(vm.post_layernorm)
        bert = model.text_decoder.bert
        self.word_emb = bert.embeddings.word_embeddings.weight.detach()
        self.pos_emb = bert.embeddings.position_embeddings.weight.detach()
        self.emb_ln = E.NormW(bert.embeddings.LayerNorm)
        self.tlayers = []
        for lyr in bert.encoder.layer:
            sa, ca = lyr.attention, lyr.crossattention
            self.tlayers.append(dict(
                sq=E.LinW(sa.self.query, dtype), sk=E.LinW(sa.self.key, dtype), sv=E.LinW(sa.self.value, dtype),
                so=E.LinW(sa.output.dense, dtype), sln=E.NormW(sa.output.LayerNorm),
                cq=E.LinW(ca.self.query, dtype), ck=E.LinW(ca.self.key, dtype), cv=E.LinW(ca.self.value, dtype),
                co=E.LinW(ca.output.dense, dtype), cln=E.NormW(ca.output.LayerNorm),
                fi=E.LinW(lyr.intermediate.dense, dtype), fo=E.LinW(lyr.output.dense, dtype), fln=E.NormW(lyr.output.LayerNorm)))
        head = model.text_decoder.cls.predictions
        self.tr = E.LinW(head.transform.dense, dtype)
        self.tr_ln = E.NormW(head.transform.LayerNorm)
        dec_w = head.decoder.weight.detach()
        self.V = dec_w.shape[0]
        self.Vpad = (self.V + 63) // 64 * 64
        self.dec_w = dec_w.to(dtype).contiguous()                                        # (V, hid)
        wt = torch.zeros(dec_w.shape[1], self.Vpad, dtype=dtype, device=dec_w.device)    # dgrad operand (hid, Vpad), zero padded
        wt[:, :self.V] = dec_w.t().to(dtype)
        self.dec_wt = wt
        self.dec_b = (head.decoder.bias if head.decoder.bias is not None else head.bias).detach().float().contiguous()
        self.eps_ls = float(getattr(model.text_decoder, "label_smoothing", getattr(cfg_t, "label_smoothing", 0.0)))
        # loss scale of the fp16 backward (activation gradients of the frozen tower underflow otherwise); halved by the trainer after an overflow
        self.grad_scale = 16384.0 if dtype == torch.float16 else 1.0

    # ------------------------------------------------------------------------------------------------
    def vision(self, tape, pix: E.Var) -> E.Var:
        """pix: Var (B, 3, S, S) 16-bit -> image embeds Var (B, 577, dim)."""
        B, Cc, S, _ = pix.v.shape
        p, g = self.patch, S // self.patch
        patches = pix.v.reshape(B, Cc, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, Cc * p * p).contiguous()
        pv = E.Var(patches)
        if tape is not None:
            def bwd_patch():
                if pv.g is not None:
                    E._acc(pix, pv.g.reshape(B, g, g, Cc, p, p).permute(0, 3, 1, 4, 2, 5).reshape(B, Cc, S, S).contiguous())
            tape.record(bwd_patch)
        x = E.linear(tape, pv, self.patch_lin)                           # (B, 576, dim)
        xc = E.Var(torch.cat([self.cls.expand(B, 1, -1), x.v], 1).contiguous())
        if tape is not None:
            def bwd_cat():
                if xc.g is not None:
                    E._acc(x, xc.g[:, 1:].contiguous())
            tape.record(bwd_cat)
        h = _add_const(tape, xc, self.pos[:, : xc.v.shape[1]].expand(B, -1, -1).contiguous())
        for L in self.vlayers:
            y = E.layernorm(tape, h, L["n1"])
            a = _mha(tape, E.linear(tape, y, L["q"]), E.linear(tape, y, L["k"]), E.linear(tape, y, L["v"]), self.heads_v)
            h = E.linear(tape, a, L["proj"], residual=h)
            y = E.layernorm(tape, h, L["n2"])
            h = E.linear(tape, _gelu(tape, E.linear(tape, y, L["fc1"])), L["fc2"], residual=h)
        return E.layernorm(tape, h, self.post_ln)

    def decoder(self, tape, img: E.Var, input_ids, attention_mask):
        """returns hidden Var (B, T, hid) of the text decoder (causal LM over the prompt tokens, cross-attending the image)."""
        B, T = input_ids.shape
        emb = self.word_emb[input_ids] + self.pos_emb[:T][None]
        h = E.Var(ops.layernorm_fwd(emb.to(self.dtype).contiguous(), self.emb_ln.gamma, self.emb_ln.beta, self.emb_ln.eps)[0], needs_grad=False)
        lens = attention_mask.sum(1).to(torch.int32).contiguous()
        for L in self.tlayers:
            a = _mha(tape, E.linear(tape, h, L["sq"]), E.linear(tape, h, L["sk"]), E.linear(tape, h, L["sv"]), self.heads_t,
                     kv_lens=lens, causal=True)
            h = E.layernorm(tape, E.linear(tape, a, L["so"], residual=h), L["sln"])
            c = _mha(tape, E.linear(tape, h, L["cq"]), E.linear(tape, img, L["ck"]), E.linear(tape, img, L["cv"]), self.heads_t)
            h = E.layernorm(tape, E.linear(tape, c, L["co"], residual=h), L["cln"])
            f = _gelu(tape, E.linear(tape, h, L["fi"]))
            h = E.layernorm(tape, E.linear(tape, f, L["fo"], residual=h), L["fln"])
        return h

    def caption_loss_tape(self, tape, pix: E.Var, input_ids, attention_mask, labels):
        img = self.vision(tape, pix)
        h = self.decoder(tape, img, input_ids, attention_mask)
        B, T, hid = h.v.shape
        # next-token prediction: logits[:, :-1] vs labels[:, 1:]  (modeling_blip_text.py:768-772)
        hs = E.Var(h.v[:, :-1].contiguous().reshape(B * (T - 1), hid))
        if tape is not None:
            def bwd_slice():
                if hs.g is not None:
                    g = torch.zeros_like(h.v)
                    g[:, :-1] = hs.g.reshape(B, T - 1, hid)
                    E._acc(h, g)
            tape.record(bwd_slice)
        t = E.layernorm(tape, _gelu(tape, E.linear(tape, hs, self.tr)), self.tr_ln)
        R = B * (T - 1)
        logits = ops.gemm([t.v], [self.dec_w], bias=self.dec_b, out_fp32=True)              # (R, V) fp32
        lab = labels[:, 1:].contiguous().reshape(-1).to(torch.int64)
        stats, out2 = ce_fwd(logits, lab, self.V, self.eps_ls)
        loss = E.Var(out2[:1])
        if tape is not None:
            def bwd_ce():
                gout = loss.g if loss.g is not None else torch.ones(1, device=logits.device)
                d = ce_bwd(logits, lab, stats, out2, gout, self.V, self.Vpad, self.eps_ls, self.dtype)
                E._acc(t, ops.gemm([d], [self.dec_wt]))                                       # (R, hid)
            tape.record(bwd_ce)
        return loss

    # ---- torch-autograd facing entry (one node): pixel_values (B,3,384,384) fp32 -> scalar loss
    def caption_loss(self, pixel_values, input_ids, attention_mask, labels):
        return _BlipLossFn.apply(self, pixel_values, input_ids, attention_mask, labels)


class _BlipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng: BlipEngine, pix, ids, mask, labels):
        need = ctx.needs_input_grad[1]
        tape = E.Tape() if need else None
        pv = E.Var(pix.to(eng.dtype).contiguous())
        loss = eng.caption_loss_tape(tape, pv, ids, mask, labels)
        ctx.tape, ctx.pv, ctx.loss, ctx.dt, ctx.eng = tape, pv, loss, pix.dtype, eng
        return loss.v.reshape(()).clone()

    @staticmethod
    def backward(ctx, g):
        S = ctx.eng.grad_scale
        ctx.loss.g = (g.reshape(1).float() * S).contiguous()
        ctx.tape.backward()
        gp = (ctx.pv.g.float() / S).to(ctx.dt) if ctx.pv.g is not None else None
        ctx.tape = ctx.pv = ctx.loss = None
        return None, gp, None, None, None
