"""Multi-head attention for the UNet / VAE / BLIP executors: softmax(scale * q k^T) v, optional fp32 probability
export (the hooked forward of attn_utils/tc_attn_utils.py:104-161 hands P to the AttentionStore) and a backward that
accepts an extra dP from the attention-map loss.

Every call runs the hand-written tcgen05 kernels (csrc/attention.cu, csrc/attention_bwd.cu) through the C ABI; a shape
they do not cover raises.  There is no library (aten) path in the product: the torch comparator the CPU logic tests run
the executors on lives in tests/cpu_ops_emulation.py.
"""
from __future__ import annotations

import torch

import ctypes as C

from . import _lib

LIBRARY_CALLS = 0        # always 0: kept so bench.py's per-module sum stays meaningful if a library call is ever introduced
NATIVE_HEAD_DIMS = (16, 32, 40, 64, 80, 128, 160)
_vp, _i, _f = C.c_void_p, C.c_int, C.c_float
_lib.register_signature("comat_attention_fwd", [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _i, _vp])
_lib.register_signature("comat_attention_fwd_strided", [_vp] * 7 + [_i] * 5 + [C.c_longlong] * 3 + [_f, _i, _vp, _i, _i, _vp])
_lib.register_signature("comat_attention_bwd", [_vp] * 12 + [_i, _i, _i, _i, _i, _f, _i, _vp, _i, _vp])
_lib.register_signature("comat_attention_bwd_strided", [_vp] * 12 + [_i] * 5 + [C.c_longlong] * 3 + [_f, _i, _vp, _i, _i, _vp])
_DT = {torch.float16: 1, torch.bfloat16: 2}


def native_supported(q, k, heads, export_probs):
    d = q.shape[-1] // heads
    return (q.is_cuda and q.dtype in _DT and d in NATIVE_HEAD_DIMS and (not export_probs or k.shape[1] <= 128))


def _rows(x):
    """(n, L, C) operand as the kernel reads it: rows of C contiguous elements at a uniform pitch.  Column slices of a wider
    matrix (fused q|k|v / k|v projection outputs) are passed in place; anything else is made contiguous."""
    n, L, Cc = x.shape
    if (x.stride(2) == 1 and x.stride(1) % 8 == 0 and x.stride(1) >= Cc and (n == 1 or x.stride(0) == L * x.stride(1))
            and x.data_ptr() % 16 == 0):
        return x, x.stride(1)
    return x.contiguous(), Cc


def attention_fwd_native(q, k, v, heads, export_probs=False, need_lse=False, kv_lens=None, causal=False, export_from=0):
    """tcgen05 fused attention forward (csrc/attention.cu).  returns (o, probs | None, lse | None).  ``export_from`` = first sample
    whose probabilities are exported (probs covers samples export_from .. n-1)."""
    (q, q_ld), (k, k_ld), (v, v_ld) = _rows(q), _rows(k), _rows(v)
    n, Lq, Cq = q.shape
    Lk = k.shape[1]
    d = Cq // heads
    L = _lib.lib()
    L.comat_attention_workspace_bytes.restype = C.c_size_t
    L.comat_attention_workspace_bytes.argtypes = [_i, _i, _i, _i]
    ws = torch.empty(int(L.comat_attention_workspace_bytes(n, Lk, heads, d)), dtype=torch.uint8, device=q.device)
    o = torch.empty(n, Lq, Cq, dtype=q.dtype, device=q.device)
    probs = torch.empty((n - export_from) * heads, Lq, Lk, dtype=torch.float32, device=q.device) if export_probs else None
    lse = torch.empty(n * heads, Lq, dtype=torch.float32, device=q.device) if need_lse else None
    _lib.check(L.comat_attention_fwd_strided(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(),
                                             None if probs is None else probs.data_ptr(), None if lse is None else lse.data_ptr(),
                                             ws.data_ptr(), n, Lq, Lk, heads, d, q_ld, k_ld, v_ld, float(d) ** -0.5, _DT[q.dtype],
                                             None if kv_lens is None else kv_lens.data_ptr(), int(causal), int(export_from) if export_probs else 0,
                                             _lib.stream_ptr()),
               "attention_fwd")
    _lib.count_launch(1)
    return o, probs, lse


def _split(x, heads):
    n, L, Cc = x.shape
    return x.reshape(n, L, heads, Cc // heads).permute(0, 2, 1, 3)        # (n, h, L, d)


def attention_unfused_fwd(q, k, v):
    """single-head attention with a head dim beyond the fused kernel's TMEM budget (VAE mid-block: d = 512, 4096 tokens):
    S = scale Q K^T and O = P V on the tcgen05 GEMM, row softmax in between.  returns (o, [P per sample])."""
    from . import ops
    n, L, d = q.shape
    outs, ps = [], []
    for b in range(n):
        s = ops.gemm([q[b]], [k[b]], alpha=float(d) ** -0.5)                 # (Lq, Lk)
        p = ops.softmax_rows(s)
        outs.append(ops.gemm([p], [ops.transpose16(v[b], 8)]))               # P V : B operand = V^T (d, Lk)
        ps.append(p)
    return torch.stack(outs), ps


def attention_unfused_bwd(q, k, v, ps, do):
    from . import ops
    n, L, d = q.shape
    scale = float(d) ** -0.5
    dq, dk, dv = [], [], []
    for b in range(n):
        p, g = ps[b], do[b].contiguous()
        dv.append(ops.gemm([ops.transpose16(p, 8)], [ops.transpose16(g, 8)]))      # P^T dO
        dp = ops.gemm([g], [v[b]])                                                  # dO V^T
        ds = ops.softmax_rows(p, dp)
        dq.append(ops.gemm([ds], [ops.transpose16(k[b], 8)], alpha=scale))          # dS K
        dk.append(ops.gemm([ops.transpose16(ds, 8)], [ops.transpose16(q[b], 8)], alpha=scale))   # dS^T Q
    return torch.stack(dq), torch.stack(dk), torch.stack(dv)


def attention_fwd(q, k, v, heads, export_probs=False, need_bwd=False, export_from=0):
    """q: (n, Lq, C), k/v: (n, Lk, C) 16-bit.  returns (o (n, Lq, C), probs fp32 ((n - export_from)*heads, Lq, Lk) | None, saved)"""
    if (heads == 1 and not export_probs and q.is_cuda and q.dtype in _DT and q.shape[-1] > 160
            and q.shape[-1] % 64 == 0 and k.shape[1] % 8 == 0 and k.shape[1] <= 8192):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        o, ps = attention_unfused_fwd(q, k, v)
        return o, None, (("unfused", q, k, v, ps) if need_bwd else None)
    if not native_supported(q, k, heads, export_probs):
        raise _lib.ComatError(f"attention: no native kernel for this call (device {q.device.type}, dtype {q.dtype}, head dim "
                              f"{q.shape[-1] // heads}, keys {k.shape[1]}, export={export_probs}); comat_b200 has no fallback path")
    o, probs, lse = attention_fwd_native(q, k, v, heads, export_probs, need_lse=need_bwd, export_from=export_from)
    if not need_bwd:
        return o, probs, None
    return o, probs, ("native", q, k, v, o, lse, probs, heads, export_from)   # strided views are read in place by the backward too


def attention_bwd_native(q, k, v, o, lse, probs, heads, do, dprobs, kv_lens=None, causal=False, dp_from=0):
    """tcgen05 fused attention backward (csrc/attention_bwd.cu): (dq, dk, dv).  ``dp_from``: first sample ``probs`` / ``dprobs`` cover."""
    (q, q_ld), (k, k_ld), (v, v_ld) = _rows(q), _rows(k), _rows(v)
    n, Lq, Cq = q.shape
    Lk = k.shape[1]
    d = Cq // heads
    L = _lib.lib()
    L.comat_attention_bwd_workspace_bytes.restype = C.c_size_t
    L.comat_attention_bwd_workspace_bytes.argtypes = [_i, _i, _i, _i, _i]
    ws = torch.empty(int(L.comat_attention_bwd_workspace_bytes(n, Lq, Lk, heads, d)), dtype=torch.uint8, device=q.device)
    do = torch.zeros_like(o) if do is None else do.contiguous()
    dq = torch.empty(n, Lq, Cq, dtype=q.dtype, device=q.device)
    dk = torch.empty(n, Lk, Cq, dtype=q.dtype, device=q.device)
    dv = torch.empty(n, Lk, Cq, dtype=q.dtype, device=q.device)
    if dprobs is not None:
        dprobs = dprobs.contiguous().float()
    _lib.check(L.comat_attention_bwd_strided(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                             None if dprobs is None else probs.data_ptr(), None if dprobs is None else dprobs.data_ptr(),
                                             dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), n, Lq, Lk, heads, d,
                                             q_ld, k_ld, v_ld, float(d) ** -0.5, _DT[q.dtype],
                                             None if kv_lens is None else kv_lens.data_ptr(), int(causal), int(dp_from) if dprobs is not None else 0,
                                             _lib.stream_ptr()),
               "attention_bwd")
    _lib.count_launch(3)          # row-statistics prep + dQ kernel + dK/dV kernel
    return dq, dk, dv


def attention_bwd(saved, do, dprobs):
    if saved[0] == "unfused":
        _, q, k, v, ps = saved
        return attention_unfused_bwd(q, k, v, ps, do)
    _, q, k, v, o, lse, probs, heads, export_from = saved
    return attention_bwd_native(q, k, v, o, lse, probs, heads, do, dprobs, dp_from=export_from)
