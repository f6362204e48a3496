"""Multi-head attention for the UNet / VAE / BLIP executors: softmax(scale * q k^T) v, optional fp32 probability
export (the hooked forward of attn_utils/tc_attn_utils.py:104-161 hands P to the AttentionStore) and a backward that
accepts an extra dP from the attention-map loss.

IMPL:
  "native" — hand-written tcgen05 kernels (csrc/attention.cu) through the C ABI.
  "torch"  — library path (aten SDPA / bmm+softmax), kept ONLY as the bring-up comparator for shapes the native
             kernel does not cover yet; every use is counted in LIBRARY_CALLS and reported by bench.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

IMPL = "torch"
LIBRARY_CALLS = 0


def _split(x, heads):
    n, L, C = x.shape
    return x.reshape(n, L, heads, C // heads).permute(0, 2, 1, 3)        # (n, h, L, d)


def attention_fwd(q, k, v, heads, export_probs=False, need_bwd=False):
    """q: (n, Lq, C), k/v: (n, Lk, C) 16-bit.  returns (o (n, Lq, C), probs fp32 (n*heads, Lq, Lk) | None, saved)"""
    global LIBRARY_CALLS
    LIBRARY_CALLS += 1
    n, Lq, C = q.shape
    d = C // heads
    qh, kh, vh = _split(q, heads), _split(k, heads), _split(v, heads)
    probs = None
    if export_probs:
        s = torch.matmul(qh.float(), kh.float().transpose(-1, -2)) * d ** -0.5
        p = s.softmax(-1)
        probs = p.reshape(n * heads, Lq, -1)
        o = torch.matmul(p.to(q.dtype), vh)
    else:
        o = F.scaled_dot_product_attention(qh, kh, vh)
    o = o.permute(0, 2, 1, 3).reshape(n, Lq, C).contiguous()
    saved = (q, k, v, heads, export_probs) if need_bwd else None
    return o, probs, saved


def attention_bwd(saved, do, dprobs):
    global LIBRARY_CALLS
    LIBRARY_CALLS += 1
    q, k, v, heads, export = saved
    n, Lq, C = q.shape
    d = C // heads
    with torch.enable_grad():
        q_, k_, v_ = (t.detach().requires_grad_(True) for t in (q, k, v))
        qh, kh, vh = _split(q_, heads), _split(k_, heads), _split(v_, heads)
        if export:
            s = torch.matmul(qh.float(), kh.float().transpose(-1, -2)) * d ** -0.5
            p = s.softmax(-1)
            o = torch.matmul(p.to(q.dtype), vh).permute(0, 2, 1, 3).reshape(n, Lq, C)
            outs, grads = [], []
            if do is not None:
                outs.append(o); grads.append(do)
            if dprobs is not None:
                outs.append(p.reshape(n * heads, Lq, -1)); grads.append(dprobs)
            dq, dk, dv = torch.autograd.grad(outs, (q_, k_, v_), grads)
        else:
            o = F.scaled_dot_product_attention(qh, kh, vh).permute(0, 2, 1, 3).reshape(n, Lq, C)
            dq, dk, dv = torch.autograd.grad(o, (q_, k_, v_), do)
    return dq.contiguous(), dk.contiguous(), dv.contiguous()
