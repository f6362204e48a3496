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


@pytest.mark.parametrize("n,L,T,H,d", [(2, 1024, 77, 8, 40), (2, 256, 77, 8, 80), (1, 300, 77, 10, 64), (2, 64, 77, 8, 160)])
def test_attention_forward_reads_fused_projection_slices_in_place(n, L, T, H, d):
    """q|k|v column slices of one (tokens, 3C) projection output (self-attention) and k|v slices of the cached (n*T, 2C)
    context projection (cross-attention) go to the kernel with their row pitch - results equal the contiguous call."""
    from comat_b200 import attention as A
    torch.manual_seed(L + d)
    C = H * d
    qkv = torch.randn(n, L, 3 * C, device="cuda").half()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    assert A._rows(k)[1] == 3 * C and A._rows(k)[0].data_ptr() == k.data_ptr()
    o, _, lse = A.attention_fwd_native(q, k, v, H, need_lse=True)
    o2, _, lse2 = A.attention_fwd_native(q.contiguous(), k.contiguous(), v.contiguous(), H, need_lse=True)
    assert torch.equal(o, o2) and torch.equal(lse, lse2)
    o_ref, _, _ = ref_attn(q, k, v, H)
    assert rel(o.float(), o_ref) < 3e-3
    kv = torch.randn(n, T, 2 * C, device="cuda").half()
    qc = torch.randn(n, L, C, device="cuda").half()
    o, p, _ = A.attention_fwd_native(qc, kv[..., :C], kv[..., C:], H, export_probs=True)
    o2, p2, _ = A.attention_fwd_native(qc, kv[..., :C].contiguous(), kv[..., C:].contiguous(), H, export_probs=True)
    assert torch.equal(o, o2) and torch.equal(p, p2)


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


@pytest.mark.parametrize("n,L,H,d,causal", [(3, 24, 12, 64, True), (2, 150, 4, 64, True), (3, 40, 4, 64, False), (2, 300, 4, 64, False),
                                            (2, 330, 8, 40, False)])
def test_attention_padding_and_causal_masks_fwd_bwd(n, L, H, d, causal):
    """BLIP text decoder self-attention: causal mask x key-padding mask (HF modeling_blip_text.py:496-545)."""
    from comat_b200 import attention as A
    torch.manual_seed(L)
    dtype = torch.float16
    q, k, v, do = (torch.randn(n, L, H * d, device="cuda").to(dtype) for _ in range(4))
    lens = torch.tensor([L, max(1, L // 2), max(1, L - 3)][:n], dtype=torch.int32, device="cuda")
    o, _, lse = A.attention_fwd_native(q, k, v, H, need_lse=True, kv_lens=lens, causal=causal)
    dq, dk, dv = A.attention_bwd_native(q, k, v, o, lse, None, H, do, None, kv_lens=lens, causal=causal)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    sp = lambda x: x.reshape(n, L, H, d).permute(0, 2, 1, 3)
    s = sp(qr) @ sp(kr).transpose(-1, -2) * d ** -0.5
    idx = torch.arange(L, device="cuda")
    mask = idx[None, None, None, :] < lens[:, None, None, None]
    if causal:
        mask = mask & (idx[None, None, None, :] <= idx[None, None, :, None])
    s = s.masked_fill(~mask, float("-inf"))
    o_ref = (s.softmax(-1) @ sp(vr)).permute(0, 2, 1, 3).reshape(n, L, H * d)
    gq, gk, gv = torch.autograd.grad(o_ref, (qr, kr, vr), do.float())
    assert rel(o.float(), o_ref) < 3e-3
    for name, a, b in (("dq", dq, gq), ("dk", dk, gk), ("dv", dv, gv)):
        assert rel(a.float(), b) < 8e-3, (name, rel(a.float(), b))


def test_unfused_attention_vae_mid_block_geometry():
    """AutoencoderKL mid-block attention: 1 head, d=512, 4096 tokens — GEMM + row-softmax + GEMM on the native kernels."""
    from comat_b200 import attention as A
    torch.manual_seed(0)
    n, L, d = 2, 1024, 512
    q, k, v, do = (torch.randn(n, L, d, device="cuda").half() for _ in range(4))
    o, _, saved = A.attention_fwd(q, k, v, 1, need_bwd=True)
    assert saved[0] == "unfused"
    dq, dk, dv = A.attention_bwd(saved, do, None)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    o_ref, _, _ = ref_attn(qr, kr, vr, 1)
    gq, gk, gv = torch.autograd.grad(o_ref, (qr, kr, vr), do.float())
    assert rel(o.float(), o_ref) < 4e-3
    for name, a, b in (("dq", dq, gq), ("dk", dk, gk), ("dv", dv, gv)):
        assert rel(a.float(), b) < 1e-2, (name, rel(a.float(), b))


@pytest.mark.parametrize("n,L,T,H,d", [(2, 1024, 77, 8, 40), (1, 300, 77, 10, 64), (2, 256, 77, 8, 80)])
def test_attention_backward_reads_fused_projection_slices_in_place(n, L, T, H, d):
    from comat_b200 import attention as A
    torch.manual_seed(L + d + 1)
    C = H * d
    qkv = torch.randn(n, L, 3 * C, device="cuda").half()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    do = torch.randn(n, L, C, device="cuda").half()
    o, _, lse = A.attention_fwd_native(q, k, v, H, need_lse=True)
    got = A.attention_bwd_native(q, k, v, o, lse, None, H, do, None)
    want = A.attention_bwd_native(q.contiguous(), k.contiguous(), v.contiguous(), o, lse, None, H, do, None)
    close = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm()) < 1e-3
    assert got[0].is_contiguous() and torch.equal(got[0], want[0])
    for a, b in zip(got[1:], want[1:]):
        # dK / dV: with few key tiles the query range is split over CTAs and summed with fp32 atomics (order varies run to run)
        assert a.is_contiguous() and close(a, b)
    kv = torch.randn(n, T, 2 * C, device="cuda").half()
    qc = torch.randn(n, L, C, device="cuda").half()
    dp = torch.randn(n * H, L, T, device="cuda") * 0.5
    o, p, lse = A.attention_fwd_native(qc, kv[..., :C], kv[..., C:], H, export_probs=True, need_lse=True)
    got = A.attention_bwd_native(qc, kv[..., :C], kv[..., C:], o, lse, p, H, do, dp)
    want = A.attention_bwd_native(qc, kv[..., :C].contiguous(), kv[..., C:].contiguous(), o, lse, p, H, do, dp)
    assert torch.equal(got[0], want[0])
    for a, b in zip(got[1:], want[1:]):
        assert close(a, b)


def test_attention_exports_and_differentiates_the_conditional_half_only():
    """export_from / dp_from = n/2 (the merged attrcon call): probabilities of samples >= n/2 only, identical to the rows a full export
    gives, and a backward whose external dP covers the same samples equals the full-batch backward with zeros for the others."""
    from comat_b200 import attention as A
    torch.manual_seed(5)
    n, L, T, H, d = 4, 1024, 77, 8, 40
    C = H * d
    q = torch.randn(n, L, C, device="cuda").half()
    k, v = torch.randn(n, T, C, device="cuda").half(), torch.randn(n, T, C, device="cuda").half()
    do = torch.randn(n, L, C, device="cuda").half()
    o_f, p_f, lse_f = A.attention_fwd_native(q, k, v, H, export_probs=True, need_lse=True)
    o_h, p_h, lse_h = A.attention_fwd_native(q, k, v, H, export_probs=True, need_lse=True, export_from=n // 2)
    assert p_h.shape == ((n - n // 2) * H, L, T)
    assert torch.equal(o_f, o_h) and torch.equal(lse_f, lse_h) and torch.equal(p_h, p_f[(n // 2) * H:])
    dp_h = torch.randn_like(p_h) * 0.5
    dp_f = torch.cat([torch.zeros_like(dp_h), dp_h])
    got = A.attention_bwd_native(q, k, v, o_h, lse_h, p_h, H, do, dp_h, dp_from=n // 2)
    want = A.attention_bwd_native(q, k, v, o_f, lse_f, p_f, H, do, dp_f)
    assert torch.equal(got[0], want[0])
    for a, b in zip(got[1:], want[1:]):
        assert float((a.float() - b.float()).norm() / b.float().norm()) < 1e-3
    saved = A.attention_fwd(q, k, v, H, export_probs=True, need_bwd=True, export_from=n // 2)[2]
    again = A.attention_bwd(saved, do, dp_h)
    assert torch.equal(again[0], want[0])
