"""GPU parity: tcgen05 fused attention forward (through the C ABI) vs a plain PyTorch fp32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_attn(q, k, v, heads):
    n, Lq, C = q.shape
    d = C // heads
    sp = lambda x: x.float().reshape(n, x.shape[1], heads, d).permute(0, 2, 1, 3)
    s = sp(q) @ sp(k).transpose(-1, -2) * d ** -0.5
    p = s.softmax(-1)
    o = (p @ sp(v)).permute(0, 2, 1, 3).reshape(n, Lq, C)
    return o, p.reshape(n * heads, Lq, -1), torch.logsumexp(s, -1).reshape(n * heads, Lq)


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 3e-3), (torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("n,Lq,Lk,H,d", [(2, 256, 256, 8, 40), (2, 1024, 1024, 8, 80), (1, 256, 256, 8, 160), (2, 64, 64, 8, 160),
                                         (1, 4096, 4096, 2, 40), (2, 577, 577, 4, 64), (3, 300, 77, 8, 40), (2, 20, 577, 12, 64),
                                         (2, 128, 200, 4, 16), (1, 384, 130, 2, 32)])
def test_attention_forward(n, Lq, Lk, H, d, dtype, tol):
    from comat_b200 import attention as A
    torch.manual_seed(Lq + Lk + d)
    q = torch.randn(n, Lq, H * d, device="cuda").to(dtype)
    k = torch.randn(n, Lk, H * d, device="cuda").to(dtype)
    v = torch.randn(n, Lk, H * d, device="cuda").to(dtype)
    o, _, lse = A.attention_fwd_native(q, k, v, H, export_probs=False, need_lse=True)
    o_ref, _, lse_ref = ref_attn(q, k, v, H)
    assert rel(o.float(), o_ref) < tol, rel(o.float(), o_ref)
    assert rel(lse, lse_ref) < 1e-3


@pytest.mark.parametrize("n,HW,H,d", [(2, 4096, 8, 40), (2, 1024, 8, 80), (3, 256, 8, 160), (2, 64, 8, 160), (2, 256, 10, 64)])
def test_cross_attention_with_probability_export(n, HW, H, d):
    """UNet cross-attention: 77 text tokens, P exported fp32 (n*H, HW, 77) — rows sum to 1, match the reference softmax."""
    from comat_b200 import attention as A
    torch.manual_seed(HW)
    q = torch.randn(n, HW, H * d, device="cuda").half()
    k = torch.randn(n, 77, H * d, device="cuda").half()
    v = torch.randn(n, 77, H * d, device="cuda").half()
    o, p, _ = A.attention_fwd_native(q, k, v, H, export_probs=True)
    o_ref, p_ref, _ = ref_attn(q, k, v, H)
    assert p.shape == (n * H, HW, 77) and p.dtype == torch.float32
    assert float((p.sum(-1) - 1).abs().max()) < 1e-5
    assert rel(p, p_ref) < 2e-3 and rel(o.float(), o_ref) < 3e-3


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 6e-3), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize("n,Lq,Lk,H,d,ext", [(2, 256, 256, 8, 40, False), (1, 1024, 1024, 4, 80, False), (2, 256, 256, 4, 160, False),
                                             (2, 64, 64, 8, 160, False), (1, 4096, 4096, 1, 40, False), (2, 577, 577, 2, 64, False),
                                             (2, 20, 577, 4, 64, False), (3, 300, 77, 8, 40, True), (2, 1024, 77, 8, 80, True),
                                             (2, 256, 77, 8, 160, True), (2, 100, 130, 2, 32, False), (2, 4096, 77, 8, 40, True)])
def test_attention_backward(n, Lq, Lk, H, d, ext, dtype, tol):
    from comat_b200 import attention as A
    torch.manual_seed(Lq * 3 + Lk + d)
    q = torch.randn(n, Lq, H * d, device="cuda").to(dtype)
    k = torch.randn(n, Lk, H * d, device="cuda").to(dtype)
    v = torch.randn(n, Lk, H * d, device="cuda").to(dtype)
    do = torch.randn(n, Lq, H * d, device="cuda").to(dtype)
    dp = (torch.randn(n * H, Lq, Lk, device="cuda") * 0.5) if ext else None
    o, probs, lse = A.attention_fwd_native(q, k, v, H, export_probs=ext, need_lse=True)
    dq, dk, dv = A.attention_bwd_native(q, k, v, o, lse, probs, H, do, dp)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    o_ref, p_ref, _ = ref_attn(qr, kr, vr, H)
    outs, grads = [o_ref], [do.float()]
    if ext:
        outs.append(p_ref); grads.append(dp)
    gq, gk, gv = torch.autograd.grad(outs, (qr, kr, vr), grads)
    for name, a, b in (("dq", dq, gq), ("dk", dk, gk), ("dv", dv, gv)):
        assert rel(a.float(), b) < tol, (name, rel(a.float(), b))
