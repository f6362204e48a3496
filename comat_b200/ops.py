"""Thin Python bindings of the C-ABI compute entry points (include/comat_b200.h).

Tensors are torch CUDA tensors used purely as device memory + stream plumbing; every function here ends in a
``comat_*`` C call on the current CUDA stream.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib

DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


class GemmParams(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("dtype", C.c_int32), ("n_seg", C.c_int32),
                ("a", C.c_void_p * 2), ("a_ld", C.c_int64 * 2), ("a_k", C.c_int32 * 2),
                ("b", C.c_void_p * 2), ("b_ld", C.c_int64 * 2), ("b_koff", C.c_int32 * 2),
                ("conv", C.c_int32), ("n_img", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("n_taps", C.c_int32),
                ("c_total", C.c_int32), ("tap_dh", C.c_int32 * 9), ("tap_dw", C.c_int32 * 9),
                ("alpha", C.c_float), ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("rows_per_group", C.c_int32),
                ("act", C.c_int32), ("residual", C.c_void_p), ("res_ld", C.c_int64), ("out16", C.c_void_p),
                ("out_ld", C.c_int64), ("out32", C.c_void_p), ("out32_ld", C.c_int64), ("force_bn", C.c_int32),
                ("split_k", C.c_int32), ("accumulate", C.c_int32), ("splitk_ws", C.c_void_p), ("rowvec_ld", C.c_int64),
                ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32), ("b_dtype", C.c_int32), ("force_kernel", C.c_int32),
                ("gn_sums", C.c_void_p), ("gn_groups", C.c_int32), ("gn_rows_per_image", C.c_int32)]


_lib.register_signature("comat_gemm", [C.POINTER(GemmParams), C.c_void_p])
_lib.register_signature("comat_gemm_gn_supported", [C.POINTER(GemmParams)])

ACT = {None: 0, "none": 0, "silu": 1, "gelu": 2, "geglu": 3}
KERNEL = {None: 0, "auto": 0, "tile": 1, "persist": 2, "pair": 3}


def _load_tuning():
    """measured (tile width, kernel, split-K) choices per GEMM shape of the SD1.5 / SDXL step, written by tools/tune_gemm.py
    on a B200 (comat_b200/gemm_tuning.json).  Shapes that are not in the table use the C side's heuristics."""
    import json, os
    path = os.environ.get("COMAT_GEMM_TUNING_FILE") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_tuning.json")
    if os.environ.get("COMAT_GEMM_TUNING", "1") == "0" or not os.path.exists(path):
        return {}
    with open(path) as fh:
        raw = json.load(fh)
    return {tuple(tuple(x) if isinstance(x, list) else x for x in e["key"]): (e["bn"], e["kernel"], e["split_k"]) for e in raw["entries"]}


TUNING = _load_tuning()
PROFILE = None      # bench.py sets {"flops": 0.0, "events": []} for one instrumented step (per-launch CUDA events)
TAPS_3x3 = [(dh, dw) for dh in (-1, 0, 1) for dw in (-1, 0, 1)]


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


# GroupNorm statistics accumulated by a producing GEMM's epilogue (comat_gemm_params.gn_sums) must start from zero: the
# executors zero ONE arena per network call (gn_arena_reset) and every fused producer takes its (image, group, 2) slice from it,
# instead of one memset launch per GroupNorm.  Slices are consumed by the GroupNorm that follows on the same stream, so the arena
# is reused by the next call.
_GN_ARENA_FLOATS = 1 << 18
_gn_arena = {}


def gn_arena_reset(device):
    a = _gn_arena.get(device)
    if a is None:
        a = _gn_arena[device] = {"buf": torch.zeros(_GN_ARENA_FLOATS, dtype=torch.float32, device=device), "off": 0}
    else:
        a["buf"].zero_()
        a["off"] = 0


def _gn_take(floats: int, device) -> torch.Tensor:
    a = _gn_arena.get(device)
    if a is None or a["off"] + floats > _GN_ARENA_FLOATS:
        return torch.zeros(floats, dtype=torch.float32, device=device)
    o = a["off"]
    a["off"] = o + (floats + 3) // 4 * 4
    return a["buf"][o:o + floats]


def gemm(a_segs: Sequence[torch.Tensor], b_segs: Sequence[torch.Tensor], *, b_koff: Sequence[int] = (0, 0),
         bias: Optional[torch.Tensor] = None, rowvec: Optional[torch.Tensor] = None, rows_per_group: int = 1,
         act=None, residual: Optional[torch.Tensor] = None, alpha: float = 1.0, out: Optional[torch.Tensor] = None,
         out_fp32: bool = False, conv_taps=None, c_total: int = 0, force_bn: int = 0, split_k: int = 0,
         accumulate: bool = False, kernel=None, gn: Optional[tuple] = None):
    """out[m,n] = act(alpha * sum_s A_s[m,:] . B_s[n,:] + bias[n] + rowvec[m // rows_per_group, n]) + residual[m,n]

    ``gn = (groups, rows_per_image)``: also accumulate the GroupNorm statistics of ``out`` in the epilogue; returns
    ``(out, sums)`` with sums (images * groups * 2) fp32 = (sum, sum of squares) per (image, group), or ``(out, None)`` when the
    problem cannot take the fused statistics (split-K, fp32 output, images smaller than 32 rows): the caller then runs the
    two-pass GroupNorm.

    plain mode : A_s is (M, K_s) (last dim contiguous); B_s is (N, >=K_s) K-major.
    conv mode  : ``conv_taps`` = [(dh, dw), ...]; A_s is NHWC (n, H, W, C_s) contiguous; B_s is (N, taps*c_total)
                 with k = tap*c_total + b_koff[s] + c; returns (n, H, W, N).
    """
    a0 = a_segs[0]
    _lib.require_cuda(a0)
    dt = a0.dtype
    if dt not in (torch.float16, torch.bfloat16):
        raise _lib.ComatError("gemm: A must be fp16 or bf16")
    p = GemmParams()
    conv = conv_taps is not None
    if conv:
        n_img, H, W, _ = a0.shape
        M = n_img * H * W
        p.conv, p.n_img, p.H, p.W, p.n_taps = 1, n_img, H, W, len(conv_taps)
        p.c_total = c_total if c_total else sum(a.shape[-1] for a in a_segs)
        for i, (dh, dw) in enumerate(conv_taps):
            p.tap_dh[i], p.tap_dw[i] = dh, dw
    else:
        M = a0.numel() // a0.shape[-1]
    N = b_segs[0].shape[0]
    p.M, p.N, p.dtype, p.n_seg = M, N, DT[dt], len(a_segs)
    p.b_dtype = DT[b_segs[0].dtype] if b_segs[0].dtype != dt else 0       # fp16 x bf16 operands: formats are per operand
    keep = []
    koff_auto = 0
    for s, (a, b) in enumerate(zip(a_segs, b_segs)):
        if a.dtype != dt or b.dtype not in DT or b.dtype == torch.float32 or b.dtype != b_segs[0].dtype:
            raise _lib.ComatError("gemm: A segments share one 16-bit dtype, B segments share one (possibly the other) 16-bit dtype")
        if conv:
            a = a.contiguous()
            p.a_ld[s] = a.shape[-1]
        else:
            a = a.reshape(-1, a.shape[-1])
            if a.stride(-1) != 1:
                a = a.contiguous()
            p.a_ld[s] = a.stride(0) if a.shape[0] > 1 else a.shape[-1]
        if b.stride(-1) != 1:
            b = b.contiguous()
        keep += [a, b]
        p.a[s], p.a_k[s] = a.data_ptr(), a.shape[-1]
        p.b[s], p.b_ld[s] = b.data_ptr(), b.stride(0)
        p.b_koff[s] = b_koff[s]
    p.alpha = alpha
    if bias is not None:
        bias = bias.float().contiguous()
    if rowvec is not None:
        if rowvec.dtype != torch.float32 or rowvec.stride(-1) != 1:
            rowvec = rowvec.float().contiguous()
        p.rowvec_ld = rowvec.stride(0) if rowvec.dim() == 2 and rowvec.shape[0] > 1 else rowvec.shape[-1]
    p.bias, p.rowvec, p.rows_per_group, p.act = _p(bias), _p(rowvec), rows_per_group, ACT[act]
    if residual is not None:
        residual = residual.reshape(M, N) if residual.is_contiguous() else residual.contiguous().reshape(M, N)
        if residual.dtype != dt:
            raise _lib.ComatError("gemm: residual dtype mismatch")
        p.residual, p.res_ld = residual.data_ptr(), residual.stride(0)
    n_out = N // 2 if act == "geglu" else N      # fused GEGLU: interleaved (hidden_j, gate_j) weight rows -> hidden_j * gelu(gate_j)
    if out is None:
        out = torch.empty(M, n_out, dtype=torch.float32 if out_fp32 else dt, device=a0.device)
    o2 = out.reshape(M, n_out) if out.dim() != 2 else out
    if o2.stride(-1) != 1:
        raise _lib.ComatError("gemm: out must have unit inner stride")
    if o2.dtype == torch.float32:
        p.out32, p.out32_ld = o2.data_ptr(), o2.stride(0)
    else:
        p.out16, p.out_ld = o2.data_ptr(), o2.stride(0)
    if force_bn == 0 and split_k == 0 and kernel is None and TUNING and not accumulate:
        t = TUNING.get((M, N, tuple(a.shape[-1] for a in a_segs), len(conv_taps) if conv else 0))
        if t is not None:
            force_bn, kernel, split_k = t
    p.force_bn = force_bn
    p.force_kernel = KERNEL[kernel]
    if act == "geglu":
        split_k = 1
    if split_k == 0:                      # auto: fill the machine when the output has few tiles and K is long
        kb = sum((a.shape[-1] + 63) // 64 for a in a_segs) * (len(conv_taps) if conv else 1)
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        if tiles <= 74 and kb >= 32:
            split_k = max(1, min(kb // 8, 148 // tiles))
    kb_total = sum((a.shape[-1] + 63) // 64 for a in a_segs) * (len(conv_taps) if conv else 1)
    split_k = max(1, min(split_k, kb_total))
    if accumulate:                          # out (fp32) += result: atomics from every K-split CTA, no workspace / second pass
        if o2.dtype != torch.float32:
            raise _lib.ComatError("gemm: accumulate needs an fp32 out")
        p.split_k, p.accumulate = split_k, 1
    elif split_k > 1:
        ws = torch.empty(split_k * M * N, dtype=torch.float32, device=a0.device)
        p.split_k, p.splitk_ws = split_k, ws.data_ptr()
        keep.append(ws)
    sums = None
    if gn is not None:
        p.gn_groups = gn[0]
        p.gn_rows_per_image = (H * W) if conv else gn[1]
        if _lib.lib().comat_gemm_gn_supported(C.byref(p)) == 1:
            sums = _gn_take((M // p.gn_rows_per_image) * gn[0] * 2, a0.device)
            p.gn_sums = sums.data_ptr()
    if PROFILE is not None and "keys_only" not in PROFILE:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(_lib.lib().comat_gemm(C.byref(p), _lib.stream_ptr()), "gemm")
    _lib.count_launch(2 if (split_k > 1 and not accumulate) else 1)          # split-K adds the fixed-order reduction kernel
    if PROFILE is not None:
        if "keys_only" in PROFILE:
            e0 = e1 = None
        else:
            e1.record()
        fl = 2.0 * M * N * sum(a.shape[-1] for a in a_segs) * (len(conv_taps) if conv else 1)
        PROFILE["events"].append((e0, e1, (M, N, tuple(a.shape[-1] for a in a_segs), len(conv_taps) if conv else 0, int(p.split_k),
                                           bool(out_fp32)), fl))
        PROFILE["flops"] += fl
    if conv and out.dim() == 2:
        out = out.reshape(n_img, H, W, n_out)
    return (out, sums) if gn is not None else out


def gemm_tn(a_km: torch.Tensor, b_kn: torch.Tensor, *, out_fp32: bool = True, split_k: int = 0,
            accumulate_into: Optional[torch.Tensor] = None, alpha: float = 1.0) -> torch.Tensor:
    """out[m, n] = sum_k a_km[k, m] * b_kn[k, n]   (= a_km^T @ b_kn), both operands read as they lie in memory.

    The weight-gradient shape: K is the token dimension (tens of thousands), M and N are feature dimensions.  Both operands
    go through TMA as MN-major panels and the tcgen05 descriptors carry the MN-major flag, so no transposed copies are made.
    K-split partial sums are reduced in a fixed order (deterministic).  ``accumulate_into`` (fp32 (M, N), contiguous): every
    K-split CTA ADDS its partial sum to it with vector fp32 atomics (gradient accumulation over K-splits and over the K
    back-propagated sampler steps in one kernel: no partial workspace, no reduction pass, no elementwise add)."""
    _lib.require_cuda(a_km)
    dt = a_km.dtype
    if dt not in (torch.float16, torch.bfloat16) or b_kn.dtype not in (torch.float16, torch.bfloat16):
        raise _lib.ComatError("gemm_tn: operands must be 16-bit")
    if a_km.dim() != 2 or b_kn.dim() != 2 or a_km.shape[0] != b_kn.shape[0]:
        raise _lib.ComatError("gemm_tn: expected (K, M) and (K, N)")
    if a_km.stride(1) != 1:
        a_km = a_km.contiguous()
    if b_kn.stride(1) != 1:
        b_kn = b_kn.contiguous()
    K, M = a_km.shape
    N = b_kn.shape[1]
    p = GemmParams()
    p.M, p.N, p.dtype, p.n_seg = M, N, DT[dt], 1
    p.a[0], p.a_ld[0], p.a_k[0] = a_km.data_ptr(), a_km.stride(0), K
    p.b[0], p.b_ld[0] = b_kn.data_ptr(), b_kn.stride(0)
    p.a_mn_major, p.b_mn_major = 1, 1
    p.b_dtype = DT[b_kn.dtype] if b_kn.dtype != dt else 0
    p.alpha = alpha
    p.rows_per_group = 1
    kb = (K + 63) // 64
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    if split_k == 0 and tiles <= 74 and kb >= 32:
        split_k = max(1, min(kb // 8, 148 // tiles))
    acc = accumulate_into is not None and out_fp32 and accumulate_into.is_contiguous() and accumulate_into.dtype == torch.float32
    if acc:
        out = accumulate_into                  # fp32 atomics from every K-split CTA (no workspace, no reduction pass)
        p.accumulate = 1
    else:
        out = torch.empty(M, N, dtype=torch.float32 if out_fp32 else dt, device=a_km.device)
    if out_fp32:
        p.out32, p.out32_ld = out.data_ptr(), N
    else:
        p.out16, p.out_ld = out.data_ptr(), N
    ws = None
    split_k = max(1, min(split_k, kb))
    if acc:
        p.split_k = split_k
    elif split_k > 1:
        ws = torch.empty(split_k * M * N, dtype=torch.float32, device=a_km.device)
        p.split_k, p.splitk_ws = split_k, ws.data_ptr()
    timed = PROFILE is not None and "keys_only" not in PROFILE
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(_lib.lib().comat_gemm(C.byref(p), _lib.stream_ptr()), "gemm_tn")
    _lib.count_launch(2 if (split_k > 1 and not acc) else 1)
    if PROFILE is not None:
        if timed:
            e1.record()
        fl = 2.0 * M * N * K
        PROFILE["events"].append((e0 if timed else None, e1 if timed else None, (M, N, (K,), -1, int(p.split_k), bool(out_fp32)), fl))
        PROFILE["flops"] += fl
    if accumulate_into is not None and not acc:
        accumulate_into += out                 # K too short for the split-K epilogue: plain in-place add
        return accumulate_into
    return out


# ------------------------------------------------------------------------------------------------------------
# normalisation / activation / rearrangement bindings
# ------------------------------------------------------------------------------------------------------------
_vp, _i, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t
_lib.register_signature("comat_groupnorm_fwd", [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp])
_lib.register_signature("comat_groupnorm_fwd_from_sums", [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp])
_lib.register_signature("comat_groupnorm_bwd", [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp])
_lib.register_signature("comat_layernorm_fwd", [_vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _i, _vp])
_lib.register_signature("comat_layernorm_bwd", [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp])
_lib.register_signature("comat_geglu_fwd", [_vp, _vp, _ll, _i, _i, _vp])
_lib.register_signature("comat_geglu_bwd", [_vp, _vp, _vp, _ll, _i, _i, _vp])
_lib.register_signature("comat_elementwise", [_vp, _vp, _vp, _ll, _i, _f, _f, _i, _vp])
_lib.register_signature("comat_spatial", [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp])
_lib.register_signature("comat_transpose16", [_vp, _vp, _i, _i, _i, _vp])
_lib.register_signature("comat_copy2d16", [_vp, _vp, _ll, _i, _ll, _ll, _vp])
_lib.register_signature("comat_latent_to_nhwc", [_vp, _vp, _i, _i, _i, _i, _f, _i, _vp])
_lib.register_signature("comat_nhwc_to_nchw_f32", [_vp, _vp, _i, _i, _i, _i, _f, _i, _vp])


def _gn_ws(n, HW, G, dev):
    L = _lib.lib()
    L.comat_groupnorm_workspace_floats.restype = _sz
    L.comat_groupnorm_workspace_floats.argtypes = [_i, _i, _i]
    return torch.empty(int(L.comat_groupnorm_workspace_floats(n, HW, G)), dtype=torch.float32, device=dev)


def _call(name, *args, launches=1):
    _lib.check(getattr(_lib.lib(), name)(*args), name)
    _lib.count_launch(launches)


def groupnorm_fwd(x, gamma, beta, G, eps, silu):
    """x: (n, HW, C) or (n, H, W, C) 16-bit contiguous -> (y, mean_rstd)"""
    _lib.require_cuda(x)
    x = x.contiguous()
    n, C_ = x.shape[0], x.shape[-1]
    HW = x.numel() // (n * C_)
    y = torch.empty_like(x)
    mr = torch.empty(n * G * 2, dtype=torch.float32, device=x.device)
    _call("comat_groupnorm_fwd", x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mr.data_ptr(),
          _gn_ws(n, HW, G, x.device).data_ptr(), n, HW, C_, G, eps, int(silu), DT[x.dtype], _lib.stream_ptr(), launches=2)
    return y, mr


def groupnorm_fwd_from_sums(x, sums, gamma, beta, G, eps, silu):
    """GroupNorm whose (sum, sum of squares) per (image, group) were accumulated by the GEMM that produced ``x``
    (``gemm(..., gn=...)``): one pass over x.  Same return values as ``groupnorm_fwd``."""
    _lib.require_cuda(x)
    x = x.contiguous()
    n, C_ = x.shape[0], x.shape[-1]
    HW = x.numel() // (n * C_)
    y = torch.empty_like(x)
    mr = torch.empty(n * G * 2, dtype=torch.float32, device=x.device)
    _call("comat_groupnorm_fwd_from_sums", x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mr.data_ptr(),
          sums.data_ptr(), n, HW, C_, G, eps, int(silu), DT[x.dtype], _lib.stream_ptr())
    return y, mr


def groupnorm_bwd(x, dy, gamma, beta, mr, G, silu):
    x, dy = x.contiguous(), dy.contiguous()
    n, C_ = x.shape[0], x.shape[-1]
    HW = x.numel() // (n * C_)
    dx = torch.empty_like(x)
    _call("comat_groupnorm_bwd", x.data_ptr(), dy.data_ptr(), dx.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mr.data_ptr(),
          _gn_ws(n, HW, G, x.device).data_ptr(), n, HW, C_, G, int(silu), DT[x.dtype], _lib.stream_ptr(), launches=2)
    return dx


def layernorm_fwd(x, gamma, beta, eps):
    _lib.require_cuda(x)
    x = x.contiguous()
    C_ = x.shape[-1]
    rows = x.numel() // C_
    y = torch.empty_like(x)
    mr = torch.empty(rows * 2, dtype=torch.float32, device=x.device)
    _call("comat_layernorm_fwd", x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mr.data_ptr(), rows, C_, eps,
          DT[x.dtype], _lib.stream_ptr())
    return y, mr


def layernorm_bwd(x, dy, gamma, mr):
    x, dy = x.contiguous(), dy.contiguous()
    C_ = x.shape[-1]
    dx = torch.empty_like(x)
    _call("comat_layernorm_bwd", x.data_ptr(), dy.data_ptr(), dx.data_ptr(), gamma.data_ptr(), mr.data_ptr(), x.numel() // C_, C_,
          DT[x.dtype], _lib.stream_ptr())
    return dx


def geglu_fwd(hg):
    _lib.require_cuda(hg)
    hg = hg.contiguous()
    Ch = hg.shape[-1] // 2
    out = torch.empty(*hg.shape[:-1], Ch, dtype=hg.dtype, device=hg.device)
    _call("comat_geglu_fwd", hg.data_ptr(), out.data_ptr(), hg.numel() // (2 * Ch), Ch, DT[hg.dtype], _lib.stream_ptr())
    return out


def geglu_bwd(hg, dy):
    hg, dy = hg.contiguous(), dy.contiguous()
    Ch = hg.shape[-1] // 2
    d = torch.empty_like(hg)
    _call("comat_geglu_bwd", hg.data_ptr(), dy.data_ptr(), d.data_ptr(), hg.numel() // (2 * Ch), Ch, DT[hg.dtype], _lib.stream_ptr())
    return d


EW = {"silu": 0, "silu_bwd": 1, "gelu": 2, "gelu_bwd": 3, "add": 4, "scale": 5, "axpby": 6}


def elementwise(op, x, y=None, alpha=1.0, beta=1.0):
    _lib.require_cuda(x)
    x = x.contiguous()
    y = y.contiguous() if y is not None else None
    out = torch.empty_like(x)
    _call("comat_elementwise", x.data_ptr(), _p(y), out.data_ptr(), x.numel(), EW[op], alpha, beta, DT[x.dtype], _lib.stream_ptr())
    return out


def spatial(x, mode):
    """mode: 'up2' (n,H,W,C)->(n,2H,2W,C) | 'up2_bwd' | 's2d' (n,H,W,C)->(n,H/2,W/2,4C) | 'd2s'"""
    _lib.require_cuda(x)
    x = x.contiguous()
    n, H, W, C_ = x.shape
    if mode == "up2":
        out, args = torch.empty(n, 2 * H, 2 * W, C_, dtype=x.dtype, device=x.device), (n, H, W, C_, 0)
    elif mode == "up2_bwd":
        out, args = torch.empty(n, H // 2, W // 2, C_, dtype=x.dtype, device=x.device), (n, H // 2, W // 2, C_, 1)
    elif mode == "s2d":
        out, args = torch.empty(n, H // 2, W // 2, 4 * C_, dtype=x.dtype, device=x.device), (n, H, W, C_, 2)
    else:
        out, args = torch.empty(n, 2 * H, 2 * W, C_ // 4, dtype=x.dtype, device=x.device), (n, 2 * H, 2 * W, C_ // 4, 3)
    _call("comat_spatial", x.data_ptr(), out.data_ptr(), *args, DT[x.dtype], _lib.stream_ptr())
    return out


def transpose16(x, pad_to=1):
    """(R, Cc) -> (Cc, R'), R' = R rounded up to ``pad_to`` (extra columns zero) so the result can be a K-major GEMM operand."""
    _lib.require_cuda(x)
    x = x.contiguous()
    R, Cc = x.shape
    Rp = (R + pad_to - 1) // pad_to * pad_to
    out = (torch.zeros if Rp != R else torch.empty)(Cc, Rp, dtype=x.dtype, device=x.device)
    _call("comat_transpose16", x.data_ptr(), out.data_ptr(), R, Cc, Rp, _lib.stream_ptr())
    return out


def concat_channels(a, b):
    """torch.cat([a, b], dim=-1) for NHWC 16-bit tensors (two strided row copies)."""
    _lib.require_cuda(a, b)
    a, b = a.contiguous(), b.contiguous()
    Ca, Cb = a.shape[-1], b.shape[-1]
    rows = a.numel() // Ca
    out = torch.empty(*a.shape[:-1], Ca + Cb, dtype=a.dtype, device=a.device)
    _call("comat_copy2d16", a.data_ptr(), out.data_ptr(), rows, Ca, Ca, Ca + Cb, _lib.stream_ptr())
    _call("comat_copy2d16", b.data_ptr(), out.data_ptr() + 2 * Ca, rows, Cb, Cb, Ca + Cb, _lib.stream_ptr())
    return out


def latent_to_nhwc(x_nchw_f32, dtype, cpad=64, scale=1.0):
    _lib.require_cuda(x_nchw_f32)
    x = x_nchw_f32.float().contiguous()
    n, Cin, H, W = x.shape
    out = torch.empty(n, H, W, cpad, dtype=dtype, device=x.device)
    _call("comat_latent_to_nhwc", x.data_ptr(), out.data_ptr(), n, Cin, H * W, cpad, scale, DT[dtype], _lib.stream_ptr())
    return out


def nhwc_to_nchw_f32(x, cout, scale=1.0):
    _lib.require_cuda(x)
    x = x.contiguous()
    n, H, W, ld = x.shape
    out = torch.empty(n, cout, H, W, dtype=torch.float32, device=x.device)
    _call("comat_nhwc_to_nchw_f32", x.data_ptr(), out.data_ptr(), n, cout, H * W, ld, scale, DT[x.dtype], _lib.stream_ptr())
    return out


_lib.register_signature("comat_softmax_rows", [_vp, _vp, _vp, _ll, _i, _i, _i, _vp])


def softmax_rows(x, dp=None):
    """x (R, C) 16-bit -> softmax over C; with ``dp``: x are probabilities and the result is p * (dp - sum(p*dp))."""
    _lib.require_cuda(x)
    x = x.contiguous()
    out = torch.empty_like(x)
    _call("comat_softmax_rows", x.data_ptr(), _p(dp.contiguous() if dp is not None else None), out.data_ptr(), x.shape[0], x.shape[1],
          0 if dp is None else 1, DT[x.dtype], _lib.stream_ptr())
    return out


_lib.register_signature("comat_split_f32_bf16x2", [_vp, _vp, _vp, _ll, _f, _vp])


def split_f32_bf16x2(src: torch.Tensor, hi: torch.Tensor, lo: torch.Tensor, alpha: float = 1.0):
    """alpha * src (fp32, flat) ~= hi + lo (two bf16 tensors of the same numel): ~16 mantissa bits at fp32 range."""
    _lib.require_cuda(src, hi, lo)
    if src.dtype != torch.float32 or hi.dtype != torch.bfloat16 or lo.dtype != torch.bfloat16 or not (
            src.is_contiguous() and hi.is_contiguous() and lo.is_contiguous()) or hi.numel() != src.numel() or lo.numel() != src.numel():
        raise _lib.ComatError("split_f32_bf16x2: contiguous fp32 source and two bf16 destinations of the same size")
    _call("comat_split_f32_bf16x2", src.data_ptr(), hi.data_ptr(), lo.data_ptr(), src.numel(), float(alpha), _lib.stream_ptr())


_lib.register_signature("comat_gan_head_bce_fwd", [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp])
_lib.register_signature("comat_gan_head_bce_bwd", [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp])


class _GanHeadBCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eps, w, b, n_zero):
        _lib.require_cuda(eps, w, b)
        eps = eps.float().contiguous()
        w32, b32 = w.detach().float().contiguous().reshape(-1), b.detach().float().contiguous().reshape(-1)
        n, C_, H, W = eps.shape
        loss = torch.zeros(1, dtype=torch.float32, device=eps.device)
        _call("comat_gan_head_bce_fwd", eps.data_ptr(), w32.data_ptr(), b32.data_ptr(), loss.data_ptr(), n, C_, H * W, int(n_zero), _lib.stream_ptr())
        ctx.save_for_backward(eps, w32, b32)
        ctx.n_zero, ctx.w_shape, ctx.b_shape = int(n_zero), w.shape, b.shape
        return (loss / float(n * H * W)).reshape(())

    @staticmethod
    def backward(ctx, g):
        eps, w32, b32 = ctx.saved_tensors
        n, C_, H, W = eps.shape
        d_eps = torch.empty_like(eps) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(w32) if ctx.needs_input_grad[1] else None
        db = torch.zeros_like(b32) if ctx.needs_input_grad[2] else None
        gout = g.detach().float().reshape(1).contiguous()
        _call("comat_gan_head_bce_bwd", eps.data_ptr(), w32.data_ptr(), b32.data_ptr(), gout.data_ptr(), _p(d_eps), _p(dw), _p(db),
              n, C_, H * W, ctx.n_zero, _lib.stream_ptr())
        return d_eps, None if dw is None else dw.reshape(ctx.w_shape), None if db is None else db.reshape(ctx.b_shape), None


def gan_head_bce(eps, weight, bias, n_zero: int):
    """mean BCEWithLogits(Linear(C, 1)(eps.permute(0, 2, 3, 1)), target) with target 0 for the first ``n_zero`` samples and 1 for the
    rest (gan_sdxl.py:84-89, :118-132), one fused kernel forward and one backward.  eps (n, C, H, W) fp32; weight (1, C), bias (1)."""
    return _GanHeadBCE.apply(eps, weight, bias, n_zero)
