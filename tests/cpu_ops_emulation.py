"""TEST INFRASTRUCTURE: a torch (CPU, fp32/fp64) emulation of the *semantics* of the C-ABI ops in comat_b200/ops.py.

It exists so the host-side executor logic (tape, backward formulas, weight packing, segment/tap bookkeeping) can be
verified against the oracle in the CPU-only build container.  It is never imported by the product; tests install it
with ``monkeypatch`` and the product still has no CPU path.
"""
import torch
import torch.nn.functional as F


def gemm(a_segs, b_segs, *, b_koff=(0, 0), bias=None, rowvec=None, rows_per_group=1, act=None, residual=None, alpha=1.0,
         out=None, out_fp32=False, conv_taps=None, c_total=0, force_bn=0, split_k=0, accumulate=False, kernel=None, gn=None):
    a0 = a_segs[0]
    N = b_segs[0].shape[0]
    if conv_taps is not None:
        n, H, W, _ = a0.shape
        M = n * H * W
        ct = c_total if c_total else sum(a.shape[-1] for a in a_segs)
        acc = torch.zeros(n, H, W, N, dtype=torch.float64)
        for s, (a, b) in enumerate(zip(a_segs, b_segs)):
            C = a.shape[-1]
            ap = F.pad(a.double(), (0, 0, 2, 2, 2, 2))
            for t, (dh, dw) in enumerate(conv_taps):
                k0 = t * ct + b_koff[s]
                acc += ap[:, 2 + dh:2 + dh + H, 2 + dw:2 + dw + W, :] @ b[:, k0:k0 + C].double().t()
        acc = acc.reshape(M, N)
    else:
        M = a0.numel() // a0.shape[-1]
        acc = torch.zeros(M, N, dtype=torch.float64)
        for s, (a, b) in enumerate(zip(a_segs, b_segs)):
            K = a.shape[-1]
            acc += a.reshape(M, K).double() @ b[:, b_koff[s]:b_koff[s] + K].double().t()
    y = acc * alpha
    if bias is not None:
        y = y + bias.double()
    if rowvec is not None:
        y = y + rowvec.double().repeat_interleave(rows_per_group, 0)[:, :N]
    if act == "silu":
        y = F.silu(y)
    elif act == "gelu":
        y = F.gelu(y)
    elif act == "geglu":                       # interleaved columns (hidden_j, gate_j) -> hidden_j * gelu(gate_j)
        y = y[:, 0::2] * F.gelu(y[:, 1::2])
        N = N // 2
    if residual is not None:
        y = y + residual.reshape(M, N).double()
    sums = None
    if gn is not None:                         # epilogue statistics: (sum, sum of squares) per (image, group) of the fp32 results
        G, rpi = gn[0], (H * W if conv_taps is not None else gn[1])
        if M % rpi == 0 and split_k <= 1 and rpi % 32 == 0 and not out_fp32:
            yg = y.reshape(M // rpi, rpi, G, N // G)
            sums = torch.stack([yg.sum((1, 3)), (yg * yg).sum((1, 3))], -1).float().reshape(-1)
    y = y.to(torch.float32 if (out_fp32 or (out is not None and out.dtype == torch.float32)) else a0.dtype)
    if out is not None:
        if accumulate:
            out.reshape(M, N).add_(y.to(out.dtype))
        else:
            out.reshape(M, N).copy_(y)      # in-place destination (row-slice views of fused weight buffers, static graph inputs)
        y = out
    if conv_taps is not None:
        y = y.reshape(n, H, W, N)
    return (y, sums) if gn is not None else y


def gn_arena_reset(device):
    pass


def groupnorm_fwd_from_sums(x, sums, gamma, beta, G, eps, silu):
    n, C = x.shape[0], x.shape[-1]
    cnt = x.numel() // (n * G)
    s = sums.double().reshape(n, G, 2)
    mean = s[..., 0] / cnt
    var = (s[..., 1] / cnt - mean * mean).clamp_min(0)
    rstd = (var + eps).rsqrt()
    xf = x.double().reshape(n, -1, G, C // G)
    y = ((xf - mean[:, None, :, None]) * rstd[:, None, :, None]).reshape(n, -1, C) * gamma.double() + beta.double()
    if silu:
        y = F.silu(y)
    return y.reshape(x.shape).to(x.dtype), (mean[:, None, :, None], rstd[:, None, :, None], eps)


def groupnorm_fwd(x, gamma, beta, G, eps, silu):
    n, C = x.shape[0], x.shape[-1]
    xf = x.double().reshape(n, -1, G, C // G)
    mean = xf.mean((1, 3), keepdim=True)
    var = xf.var((1, 3), unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    y = ((xf - mean) * rstd).reshape(n, -1, C) * gamma.double() + beta.double()
    if silu:
        y = F.silu(y)
    return y.reshape(x.shape).to(x.dtype), (mean, rstd, eps)


def groupnorm_bwd(x, dy, gamma, beta, mr, G, silu):
    with torch.enable_grad():
        xr = x.detach().double().requires_grad_(True)
        y, _ = groupnorm_fwd(xr, gamma, beta, G, mr[2], silu)
        return torch.autograd.grad(y, xr, dy.double())[0].to(x.dtype)


def layernorm_fwd(x, gamma, beta, eps):
    return F.layer_norm(x.double(), (x.shape[-1],), gamma.double(), beta.double(), eps).to(x.dtype), eps


def layernorm_bwd(x, dy, gamma, mr):
    with torch.enable_grad():
        xr = x.detach().double().requires_grad_(True)
        y = F.layer_norm(xr, (x.shape[-1],), gamma.double(), torch.zeros_like(gamma).double(), mr)
        return torch.autograd.grad(y, xr, dy.double())[0].to(x.dtype)


def geglu_fwd(hg):
    h, g = hg.chunk(2, -1)
    return (h.double() * F.gelu(g.double())).to(hg.dtype)


def geglu_bwd(hg, dy):
    with torch.enable_grad():
        r = hg.detach().double().requires_grad_(True)
        h, g = r.chunk(2, -1)
        return torch.autograd.grad(h * F.gelu(g), r, dy.double())[0].to(hg.dtype)


def elementwise(op, x, y=None, alpha=1.0, beta=1.0):
    xd = x.double()
    yd = y.double() if y is not None else None
    if op == "silu":
        r = F.silu(xd)
    elif op == "silu_bwd":
        with torch.enable_grad():
            xr = xd.detach().requires_grad_(True)
            r = torch.autograd.grad(F.silu(xr), xr, yd)[0]
    elif op == "gelu":
        r = F.gelu(xd)
    elif op == "gelu_bwd":
        with torch.enable_grad():
            xr = xd.detach().requires_grad_(True)
            r = torch.autograd.grad(F.gelu(xr), xr, yd)[0]
    elif op == "add":
        r = xd + yd
    elif op == "scale":
        r = xd * alpha
    elif op == "axpby":
        r = alpha * xd + beta * yd
    else:
        raise NotImplementedError(op)
    return r.to(x.dtype)


def spatial(x, mode):
    n, H, W, C = x.shape
    if mode == "up2":
        return x.repeat_interleave(2, 1).repeat_interleave(2, 2)
    if mode == "up2_bwd":
        return x.reshape(n, H // 2, 2, W // 2, 2, C).sum((2, 4))
    if mode == "s2d":
        return x.reshape(n, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(n, H // 2, W // 2, 4 * C)
    return x.reshape(n, H, W, 2, 2, C // 4).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * H, 2 * W, C // 4)


def transpose16(x, pad_to=1):
    R, Cc = x.shape
    Rp = (R + pad_to - 1) // pad_to * pad_to
    out = x.new_zeros(Cc, Rp)
    out[:, :R] = x.t()
    return out


def concat_channels(a, b):
    return torch.cat([a, b], -1)


def latent_to_nhwc(x, dtype, cpad=64, scale=1.0):
    n, C, H, W = x.shape
    out = torch.zeros(n, H, W, cpad, dtype=dtype)
    out[..., :C] = (x * scale).permute(0, 2, 3, 1).to(dtype)
    return out


def nhwc_to_nchw_f32(x, cout, scale=1.0):
    return (x[..., :cout].float() * scale).permute(0, 3, 1, 2).contiguous()


def gemm_tn(a_km, b_kn, *, out_fp32=True, split_k=0, accumulate_into=None, alpha=1.0):
    y = alpha * (a_km.double().t() @ b_kn.double())
    y = y.float() if out_fp32 else y.to(a_km.dtype)
    if accumulate_into is not None:
        accumulate_into += y
        return accumulate_into
    return y


def split_f32_bf16x2(src, hi, lo, alpha=1.0):
    v = src.float() * alpha
    h = v.to(hi.dtype)
    hi.copy_(h)
    lo.copy_((v - h.float()).to(lo.dtype))


def gan_head_bce(eps, weight, bias, n_zero):
    pred = F.linear(eps.permute(0, 2, 3, 1).float(), weight.float(), bias.float())
    target = torch.ones_like(pred)
    target[:n_zero] = 0
    return F.binary_cross_entropy_with_logits(pred, target)


def install(monkeypatch):
    from comat_b200 import attention, ops
    monkeypatch.setattr(ops, "gan_head_bce", gan_head_bce)
    # CPU logic tests run the executors' attention through the torch comparator below (the product has no such path)
    monkeypatch.setattr(attention, "attention_fwd", attention_fwd)
    monkeypatch.setattr(attention, "attention_bwd", attention_bwd)
    for name in ("gemm", "gemm_tn", "split_f32_bf16x2", "groupnorm_fwd", "groupnorm_fwd_from_sums", "gn_arena_reset", "groupnorm_bwd", "layernorm_fwd", "layernorm_bwd", "geglu_fwd", "geglu_bwd",
                 "elementwise", "spatial", "transpose16", "concat_channels", "latent_to_nhwc", "nhwc_to_nchw_f32"):
        monkeypatch.setattr(ops, name, globals()[name])


# ---- attention / CE / resize emulations (BLIP executor logic tests)
def _ref_attn(q, k, v, heads, kv_lens=None, causal=False, dprobs=None):
    n, Lq, C = q.shape
    Lk = k.shape[1]
    d = C // heads
    sp = lambda x: x.double().reshape(n, x.shape[1], heads, d).permute(0, 2, 1, 3)
    s = sp(q) @ sp(k).transpose(-1, -2) * d ** -0.5
    ki = torch.arange(Lk)
    mask = torch.ones(n, 1, Lq, Lk, dtype=torch.bool)
    if kv_lens is not None:
        mask = mask & (ki[None, None, None, :] < kv_lens[:, None, None, None])
    if causal:
        mask = mask & (ki[None, None, None, :] <= torch.arange(Lq)[None, None, :, None])
    s = s.masked_fill(~mask, float("-inf"))
    p = s.softmax(-1)
    o = (p @ sp(v)).permute(0, 2, 1, 3).reshape(n, Lq, C)
    return o, p.reshape(n * heads, Lq, Lk), torch.logsumexp(s, -1).reshape(n * heads, Lq)


def attention_fwd(q, k, v, heads, export_probs=False, need_bwd=False, export_from=0):
    """comat_b200.attention.attention_fwd protocol: (o, probs fp32 of samples export_from.. | None, saved)"""
    o, p, _ = _ref_attn(q, k, v, heads)
    return o.to(q.dtype), (p[export_from * heads:].float() if export_probs else None), ((q, k, v, heads, export_from) if need_bwd else None)


def attention_bwd(saved, do, dprobs):
    q, k, v, heads, export_from = saved
    if do is None:
        do = torch.zeros_like(q)
    if dprobs is not None and export_from:
        dprobs = torch.cat([dprobs.new_zeros(export_from * heads, *dprobs.shape[1:]), dprobs])
    return attention_bwd_native(q, k, v, None, None, None, heads, do, dprobs)


def attention_fwd_native(q, k, v, heads, export_probs=False, need_lse=False, kv_lens=None, causal=False):
    o, p, lse = _ref_attn(q, k, v, heads, kv_lens, causal)
    return o.to(q.dtype), (p.float() if export_probs else None), (lse.float() if need_lse else None)


def attention_bwd_native(q, k, v, o, lse, probs, heads, do, dprobs, kv_lens=None, causal=False):
    with torch.enable_grad():
        qr, kr, vr = (t.detach().double().requires_grad_(True) for t in (q, k, v))
        o_, p_, _ = _ref_attn(qr, kr, vr, heads, kv_lens, causal)
        outs, gs = [o_], [do.double()]
        if dprobs is not None:
            outs.append(p_); gs.append(dprobs.double())
        gq, gk, gv = torch.autograd.grad(outs, (qr, kr, vr), gs)
    return gq.to(q.dtype), gk.to(q.dtype), gv.to(q.dtype)


def ce_fwd(logits, labels, V, eps):
    x = logits[:, :V].double()
    lse = torch.logsumexp(x, -1)
    valid = labels != -100
    y = labels.clamp_min(0)
    nll = lse - x.gather(1, y[:, None])[:, 0]
    smooth = lse - x.mean(-1)
    row = torch.where(valid, (1 - eps) * nll + eps * smooth, torch.zeros_like(nll))
    cnt = valid.sum().double()
    return torch.stack([lse, row], 1).float(), torch.stack([row.sum() / cnt, cnt]).float()


def ce_bwd(logits, labels, stats, out2, gout, V, Vpad, eps, dtype):
    x = logits[:, :V].double()
    p = (x - stats[:, :1].double()).exp()
    valid = (labels != -100)[:, None]
    oh = torch.zeros_like(p).scatter_(1, labels.clamp_min(0)[:, None], 1.0)
    g = float(gout.reshape(-1)[0]) / float(out2[1])
    d = torch.where(valid, g * (p - (1 - eps) * oh - eps / V), torch.zeros_like(p))
    out = torch.zeros(logits.shape[0], Vpad, dtype=dtype)
    out[:, :V] = d.to(dtype)
    return out


def install_blip(monkeypatch):
    install(monkeypatch)
    from comat_b200 import attention, blip_engine, image_ops
    monkeypatch.setattr(attention, "attention_fwd_native", attention_fwd_native)
    monkeypatch.setattr(attention, "attention_bwd_native", attention_bwd_native)
    monkeypatch.setattr(blip_engine, "ce_fwd", ce_fwd)
    monkeypatch.setattr(blip_engine, "ce_bwd", ce_bwd)

    def resize_norm(images, size, mean, std):
        x = F.interpolate(images.float(), size=(size, size), mode="bicubic", antialias=True, align_corners=False)
        return (x - torch.tensor(mean).view(1, -1, 1, 1)) / torch.tensor(std).view(1, -1, 1, 1)
    monkeypatch.setattr(image_ops, "resize_bicubic_aa_normalize", resize_norm)
