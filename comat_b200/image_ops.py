"""Image-space glue of the reward path: bicubic-antialias resize to 384 + CLIP normalise (concept_mat_utils/caption_blip.py:33-36,45),
forward and backward on one table-driven CUDA kernel (csrc/resize.cu).

The tap tables restate aten's ``_compute_indices_weights_aa`` (UpSampleKernel / upsample_bicubic2d_aa, align_corners=False,
cubic a = -0.5): scale = in/out, support = 2*max(scale,1), taps [int(c - support + .5), int(c + support + .5)) around
c = scale*(i + .5), weights cubic((j - c + .5)/max(scale,1)) normalised to sum 1.  tests/test_image_ops_cpu.py checks the dense
operator built from these tables against F.interpolate(mode='bicubic', antialias=True)."""
from __future__ import annotations

import ctypes as C
from functools import lru_cache

import torch

from . import _lib

_vp, _i = C.c_void_p, C.c_int
LIBRARY_CALLS = 0          # this module is fully native (kept so bench.py can sum the per-module library-call counters)
_lib.register_signature("comat_resample2d", [_vp] * 10 + [_i] * 8 + [_vp])


def _cubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


@lru_cache(maxsize=16)
def aa_bicubic_taps(in_size: int, out_size: int):
    """(start[out], count[out], weights[out][K]) for the forward map, K = max tap count."""
    scale = in_size / out_size
    support = 2.0 * scale if scale >= 1.0 else 2.0
    inv = 1.0 / scale if scale >= 1.0 else 1.0
    starts, counts, ws = [], [], []
    for i in range(out_size):
        center = scale * (i + 0.5)
        xmin = max(int(center - support + 0.5), 0)
        xsize = min(int(center + support + 0.5), in_size) - xmin
        w = [_cubic((j + xmin - center + 0.5) * inv) for j in range(xsize)]
        tot = sum(w)
        w = [v / tot for v in w]
        starts.append(xmin); counts.append(xsize); ws.append(w)
    K = max(counts)
    wt = torch.zeros(out_size, K, dtype=torch.float32)
    for i, w in enumerate(ws):
        wt[i, :len(w)] = torch.tensor(w, dtype=torch.float32)
    return torch.tensor(starts, dtype=torch.int32), torch.tensor(counts, dtype=torch.int32), wt


@lru_cache(maxsize=16)
def aa_bicubic_taps_transposed(in_size: int, out_size: int):
    """tables of the adjoint: for each INPUT index the contiguous run of output indices that read it, with their weights."""
    st, ct, wt = aa_bicubic_taps(in_size, out_size)
    lo = [None] * in_size
    hi = [None] * in_size
    for o in range(out_size):
        for k in range(int(ct[o])):
            j = int(st[o]) + k
            lo[j] = o if lo[j] is None else min(lo[j], o)
            hi[j] = o if hi[j] is None else max(hi[j], o)
    starts = [0 if l is None else l for l in lo]
    counts = [0 if l is None else h - l + 1 for l, h in zip(lo, hi)]
    K = max(counts)
    w = torch.zeros(in_size, K, dtype=torch.float32)
    for j in range(in_size):
        for t in range(counts[j]):
            o = starts[j] + t
            k = j - int(st[o])
            if 0 <= k < int(ct[o]):
                w[j, t] = wt[o, k]
    return torch.tensor(starts, dtype=torch.int32), torch.tensor(counts, dtype=torch.int32), w


def dense_operator(in_size: int, out_size: int) -> torch.Tensor:
    """(out, in) matrix of the 1-D resampling operator (tests / documentation)."""
    st, ct, wt = aa_bicubic_taps(in_size, out_size)
    M = torch.zeros(out_size, in_size)
    for o in range(out_size):
        M[o, int(st[o]):int(st[o]) + int(ct[o])] = wt[o, :int(ct[o])]
    return M


_dev_tables = {}


def _tables(kind, a, b, device):
    key = (kind, a, b, str(device))
    if key not in _dev_tables:
        t = aa_bicubic_taps(a, b) if kind == "f" else aa_bicubic_taps_transposed(a, b)
        _dev_tables[key] = tuple(x.to(device) for x in t)
    return _dev_tables[key]


def _resample(x, ty, tx, OH, OW, scale, shift):
    _lib.require_cuda(x)
    x = x.float().contiguous()
    B, Cc, IH, IW = x.shape
    out = torch.empty(B, Cc, OH, OW, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().comat_resample2d(x.data_ptr(), out.data_ptr(), ty[0].data_ptr(), ty[1].data_ptr(), ty[2].data_ptr(),
                                           tx[0].data_ptr(), tx[1].data_ptr(), tx[2].data_ptr(),
                                           None if scale is None else scale.data_ptr(), None if shift is None else shift.data_ptr(),
                                           B, Cc, IH, IW, OH, OW, ty[2].shape[1], tx[2].shape[1], _lib.stream_ptr()), "resample2d")
    _lib.count_launch()
    return out


_NORM_CONSTS = {}


def _norm_consts(mean, std, dev):
    """device copies of 1/std and -mean/std, made once (a torch.tensor(..., device=cuda) per call is a stream sync)"""
    key = (mean, std, str(dev))
    if key not in _NORM_CONSTS:
        _NORM_CONSTS[key] = (torch.tensor([1.0 / s for s in std], dtype=torch.float32, device=dev),
                             torch.tensor([-m / s for m, s in zip(mean, std)], dtype=torch.float32, device=dev))
    return _NORM_CONSTS[key]


class _ResizeNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, size, mean, std):
        B, Cc, IH, IW = images.shape
        dev = images.device
        inv_std, shift = _norm_consts(tuple(mean), tuple(std), dev)
        ctx.meta = (IH, IW, size, inv_std, images.dtype)
        return _resample(images, _tables("f", IH, size, dev), _tables("f", IW, size, dev), size, size, inv_std, shift)

    @staticmethod
    def backward(ctx, g):
        IH, IW, size, inv_std, dt = ctx.meta
        dev = g.device
        gi = _resample(g, _tables("t", IH, size, dev), _tables("t", IW, size, dev), IH, IW, inv_std, None)
        return gi.to(dt), None, None, None


def resize_bicubic_aa_normalize(images: torch.Tensor, size: int, mean, std) -> torch.Tensor:
    """(B,C,H,W) -> (B,C,size,size) fp32: Resize(BICUBIC, antialias=True) then Normalize(mean, std); differentiable."""
    return _ResizeNorm.apply(images, size, tuple(mean), tuple(std))
