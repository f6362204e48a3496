"""CPU: weight-packing maths (forward / dgrad / stride-2 via space-to-depth) checked with a torch emulation of the
conv-mode semantics of comat_gemm:  out[n,h,w,:] = sum_t A[n,h+dh_t,w+dw_t,:] @ B[:, t*C:(t+1)*C]^T  (zero outside)."""
import torch
import torch.nn.functional as F

from comat_b200 import unet_weights as UW


def emu_conv(a_nhwc, b, taps):
    n, H, W, C = a_nhwc.shape
    out = torch.zeros(n, H, W, b.shape[0], dtype=torch.float64)
    ap = F.pad(a_nhwc.double(), (0, 0, 2, 2, 2, 2))
    for t, (dh, dw) in enumerate(taps):
        sl = ap[:, 2 + dh:2 + dh + H, 2 + dw:2 + dw + W, :]
        out += sl @ b[:, t * C:(t + 1) * C].double().t()
    return out


def s2d(x):
    n, H, W, C = x.shape
    return x.reshape(n, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(n, H // 2, W // 2, 4 * C)


def d2s(x):
    n, h, w, C4 = x.shape
    C = C4 // 4
    return x.reshape(n, h, w, 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * h, 2 * w, C)


def test_conv3x3_fwd_and_dgrad_packing():
    torch.manual_seed(0)
    x = torch.randn(2, 6, 5, 8, dtype=torch.float64, requires_grad=True)
    w = torch.randn(12, 8, 3, 3, dtype=torch.float64)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
    out = emu_conv(x.detach(), UW.pack_conv3x3(w), UW.TAPS_3x3)
    assert torch.allclose(out, ref.detach(), atol=1e-10)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    dx = emu_conv(dy, UW.pack_conv3x3_dgrad(w), UW.TAPS_3x3)
    assert torch.allclose(dx, x.grad, atol=1e-10)
    # channel padding (conv_in: 4 -> 64 input channels)
    xp = F.pad(x.detach(), (0, 8))
    assert torch.allclose(emu_conv(xp, UW.pack_conv3x3(w, 16), UW.TAPS_3x3), ref.detach(), atol=1e-10)


def test_stride2_conv_via_s2d_fwd_and_dgrad():
    torch.manual_seed(1)
    x = torch.randn(2, 8, 6, 4, dtype=torch.float64, requires_grad=True)
    w = torch.randn(5, 4, 3, 3, dtype=torch.float64)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, stride=2, padding=1).permute(0, 2, 3, 1)
    wk, taps = UW.pack_conv_stride2(w)
    out = emu_conv(s2d(x.detach()), wk, taps)
    assert torch.allclose(out, ref.detach(), atol=1e-10)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    wd, tapsd = UW.pack_conv_stride2_dgrad(w)
    dxs = emu_conv(dy, wd, tapsd)                     # gradient wrt the space-to-depth tensor
    assert torch.allclose(d2s(dxs), x.grad, atol=1e-10)
