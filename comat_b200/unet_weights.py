"""Weight packing: diffusers-shaped parameters -> the K-major 16-bit layouts the tcgen05 GEMM/conv kernel consumes.

Forward and data-gradient (dgrad) copies are both materialised once (base weights are frozen:
training_utils/pipeline.py:66-71); 860 M params x 2 B x 2 copies = 3.4 GB of the B200's 180 GB.
"""
from __future__ import annotations

import torch

TAPS_3x3 = [(dh, dw) for dh in (-1, 0, 1) for dw in (-1, 0, 1)]
TAPS_1x1 = [(0, 0)]
TAPS_S2D = [(-1, -1), (-1, 0), (0, -1), (0, 0)]       # 2x2 taps over space-to-depth blocks (stride-2 conv)
TAPS_S2D_T = [(1, 1), (1, 0), (0, 1), (0, 0)]         # its transpose (dgrad)


def pad_channels(c: int, m: int = 64) -> int:
    return (c + m - 1) // m * m


def pack_conv3x3(w: torch.Tensor, cin_pad: int = 0):
    """(Cout, Cin, 3, 3) -> [Cout, 9*Cin'] with k = (kh*3+kw)*Cin' + c  (Cin' = Cin zero-padded to cin_pad)."""
    co, ci, kh, kw = w.shape
    cp = cin_pad or ci
    out = w.new_zeros(co, kh * kw, cp)
    out[:, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
    return out.reshape(co, kh * kw * cp).contiguous()


def pack_conv3x3_dgrad(w: torch.Tensor, cout_pad: int = 0):
    """dgrad of a stride-1 'same' conv = conv of dY with spatially flipped, channel-transposed weights:
    [Cin, 9*Cout'] with k = ((2-kh)*3 + (2-kw))*Cout' + co."""
    co, ci, kh, kw = w.shape
    cp = cout_pad or co
    wf = w.flip(2, 3).permute(1, 2, 3, 0)                       # (Cin, kh', kw', Cout)
    out = w.new_zeros(ci, kh * kw, cp)
    out[:, :, :co] = wf.reshape(ci, kh * kw, co)
    return out.reshape(ci, kh * kw * cp).contiguous()


def pack_conv_stride2(w: torch.Tensor):
    """conv3x3 stride 2 pad 1 on X == 2x2-tap stride-1 conv on space_to_depth(X) (channels 4C, block (py*2+px)):
    output o reads input rows 2o-1 (block o-1, phase 1), 2o (block o, phase 0), 2o+1 (block o, phase 1).
    returns (packed [Cout, 4 taps * 4C], taps)."""
    co, ci, _, _ = w.shape
    out = w.new_zeros(co, 4, 4, ci)                             # (Cout, tap(bh,bw), phase(py,px), C)
    for kh in range(3):
        bh, py = (0, 1) if kh == 0 else (1, kh - 1)             # tap row index (0: block o-1, 1: block o), phase
        for kw in range(3):
            bw, px = (0, 1) if kw == 0 else (1, kw - 1)
            out[:, bh * 2 + bw, py * 2 + px, :] = w[:, :, kh, kw]
    return out.reshape(co, 16 * ci).contiguous(), TAPS_S2D


def pack_conv_stride2_dgrad(w: torch.Tensor):
    """dgrad wrt the space-to-depth input: d s2d(X)[block b, phase p, c] = sum_taps dY[b - tap] W[tap, p, c, :].
    A stride-1 conv over dY (Cout channels) producing 4C channels with taps mirrored.
    returns (packed [4C, 4 taps * Cout], taps)."""
    co, ci, _, _ = w.shape
    fw, _ = pack_conv_stride2(w)
    fw = fw.reshape(co, 4, 4 * ci)                              # (Cout, tap, 4C)
    out = fw.permute(2, 1, 0).contiguous()                      # (4C, tap, Cout); tap t of fwd at offset TAPS_S2D[t]
    return out.reshape(4 * ci, 4 * co).contiguous(), TAPS_S2D_T  # dgrad reads dY at +(-offset)


def to16(t: torch.Tensor, dtype):
    return t.detach().to(dtype).contiguous()
