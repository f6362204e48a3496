"""GPU parity: tcgen05 GEMM / implicit-GEMM conv (through the C ABI) vs a plain PyTorch fp32 reference of the same op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K,bn", [(128, 160, 64, 0), (256, 320, 320, 0), (1000, 320, 768, 0), (4096, 1280, 640, 0),
                                      (77, 640, 768, 0), (300, 128, 128, 0), (512, 30524 // 4 * 4, 768, 128),
                                      (130, 4, 320, 0), (256, 96, 200, 0), (512, 512, 1024, 256), (384, 64, 64, 64)])
def test_plain_gemm(M, N, K, bn, dtype):
    from comat_b200 import ops
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").to(dtype)
    b = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dtype)
    out = ops.gemm([a], [b], force_bn=bn)
    ref = a.float() @ b.float().t()
    l2, mx = _rel(out.float(), ref)
    tol = 2e-3 if dtype == torch.float16 else 1.2e-2
    assert l2 < tol and mx < 4 * tol, (l2, mx)


@pytest.mark.parametrize("M,N,K,why", [
    (40 * 128, 1024, 320, "auto BN=256, persistent (short K), several tiles per CTA"),
    (300 * 128 + 17, 320, 192, "persistent, >2 tiles per CTA so both TMEM accumulator buffers wrap, ragged last m-tile"),
    (200 * 128, 320, 2048, "long K, several waves: one-tile kernel"),
    (640, 1280, 2560, "single wave, long K: persistent with the deep ring"),
    (150 * 128, 96, 64, "one k-block per tile, N < tile width"),
])
def test_gemm_kernel_selection_paths(M, N, K, why):
    """Both GEMM kernels (one-tile-per-CTA and persistent, see gemm_tc.cu) and the automatic 256-wide tile choice, each on a
    shape the host heuristic routes to it; exact small-integer data so any tile / accumulator-buffer mix-up is a wrong integer."""
    from comat_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randint(-2, 3, (M, K), device="cuda", generator=g).half()
    b = torch.randint(-2, 3, (N, K), device="cuda", generator=g).half()
    bias = torch.randint(-4, 5, (N,), device="cuda", generator=g).float()
    out = ops.gemm([a], [b], bias=bias, out_fp32=False)
    ref = a.float() @ b.float().t() + bias
    assert ref.abs().max() < 2048            # exactly representable in fp16
    assert torch.equal(out.float(), ref), (why, (out.float() - ref).abs().nonzero()[:8])


@pytest.mark.parametrize("K,M,N,split", [(64, 128, 128, 1), (256, 128, 64, 1), (4096, 320, 128, 0), (1000, 128, 320, 0),
                                         (8 * 4096, 1280, 128, 0), (616, 136, 776, 1), (130, 8, 24, 1)])
def test_gemm_tn_mn_major_operands_exact(K, M, N, split):
    """out = A^T B with both operands read MN-major (TMA panels + MN-major tcgen05 descriptors): the LoRA weight-gradient
    shapes.  Small-integer data, fp32 output: exact, so a wrong panel stride / k-step / major bit is a wrong integer."""
    from comat_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(K + M + N)
    a = torch.randint(-2, 3, (K, M), device="cuda", generator=g).half()
    b = torch.randint(-2, 3, (K, N), device="cuda", generator=g).half()
    # sparsify so sums stay exactly representable with long K
    a = a * (torch.rand(K, M, device="cuda", generator=g) < 0.25)
    out = ops.gemm_tn(a, b, split_k=split)
    ref = a.float().t() @ b.float()
    assert ref.abs().max() < 2 ** 24
    assert torch.equal(out, ref), (out - ref).abs().nonzero()[:8]
    out_b = ops.gemm_tn(a.bfloat16(), b.bfloat16(), out_fp32=False, split_k=split)
    assert torch.allclose(out_b.float(), ref, rtol=1e-2, atol=1e-2 * ref.abs().max().item())


def test_gemm_identity_pattern_exact():
    """small integers are exact in fp16/fp32: any descriptor / swizzle / row-mapping slip shows up as a wrong integer."""
    from comat_b200 import ops
    M, N, K = 256, 160, 128
    a = torch.randint(-3, 4, (M, K), device="cuda").half()
    b = torch.randint(-3, 4, (N, K), device="cuda").half()
    out = ops.gemm([a], [b], out_fp32=True)
    ref = a.float() @ b.float().t()
    assert torch.equal(out, ref), (out - ref).abs().nonzero()[:8]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_epilogue_bias_rowvec_act_residual_and_lora_segment(dtype):
    from comat_b200 import ops
    torch.manual_seed(0)
    M, N, K, r = 2 * 1024, 320, 320, 128
    x = torch.randn(M, K, device="cuda").to(dtype)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dtype)
    down = (torch.randn(r, K, device="cuda") / r).to(dtype)
    up = (torch.randn(N, r, device="cuda") * 0.05).to(dtype)
    bias = torch.randn(N, device="cuda")
    temb = torch.randn(2, N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(dtype)
    t = ops.gemm([x], [down])                                   # x . down^T
    for act, fn in (("none", lambda v: v), ("silu", F.silu), ("gelu", F.gelu)):
        out = ops.gemm([x, t], [w, up], bias=bias, rowvec=temb, rows_per_group=1024, act=act, residual=res, alpha=0.5)
        ref = fn(0.5 * (x.float() @ w.float().t() + t.float() @ up.float().t()) + bias + temb.repeat_interleave(1024, 0)) + res.float()
        l2, mx = _rel(out.float(), ref)
        assert l2 < (3e-3 if dtype == torch.float16 else 1.5e-2), (act, l2, mx)


@pytest.mark.parametrize("n,H,W,C,Cout", [(2, 64, 64, 320, 320), (3, 32, 32, 640, 320), (2, 16, 16, 1280, 640), (3, 8, 8, 1280, 1280),
                                          (1, 128, 128, 128, 128), (1, 256, 256, 64, 32), (2, 24, 40, 64, 64)])
def test_conv3x3_implicit_gemm(n, H, W, C, Cout):
    from comat_b200 import ops
    torch.manual_seed(H + C)
    dtype = torch.float16
    x = torch.randn(n, H, W, C, device="cuda").to(dtype)
    w = (torch.randn(Cout, C, 3, 3, device="cuda") / (9 * C) ** 0.5).to(dtype)
    bias = torch.randn(Cout, device="cuda")
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).contiguous()        # k = (kh*3+kw)*C + c
    out = ops.gemm([x], [wk], bias=bias, conv_taps=ops.TAPS_3x3)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    l2, mx = _rel(out.float(), ref)
    assert l2 < 2e-3, (l2, mx)


def test_conv_with_concat_segments_and_1x1():
    from comat_b200 import ops
    torch.manual_seed(1)
    dtype = torch.float16
    n, H, W, C1, C2, Cout = 2, 32, 32, 640, 320, 640
    h = torch.randn(n, H, W, C1, device="cuda").to(dtype)
    s = torch.randn(n, H, W, C2, device="cuda").to(dtype)
    w = (torch.randn(Cout, C1 + C2, 3, 3, device="cuda") / (9 * (C1 + C2)) ** 0.5).to(dtype)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * (C1 + C2)).contiguous()
    out = ops.gemm([h, s], [wk, wk], b_koff=(0, C1), conv_taps=ops.TAPS_3x3, c_total=C1 + C2)
    ref = F.conv2d(torch.cat([h, s], -1).float().permute(0, 3, 1, 2), w.float(), None, padding=1).permute(0, 2, 3, 1)
    l2, mx = _rel(out.float(), ref)
    assert l2 < 2e-3, (l2, mx)
    w1 = (torch.randn(Cout, C1 + C2, device="cuda") / (C1 + C2) ** 0.5).to(dtype)
    out1 = ops.gemm([h.reshape(-1, C1), s.reshape(-1, C2)], [w1, w1], b_koff=(0, C1))
    ref1 = torch.cat([h, s], -1).reshape(-1, C1 + C2).float() @ w1.float().t()
    assert _rel(out1.float(), ref1)[0] < 2e-3


def test_split_k_matches_single_pass_and_reference():
    from comat_b200 import ops
    torch.manual_seed(2)
    dt = torch.float16
    # LoRA wgrad shape: tiny output, very long K
    a = torch.randn(320, 32768, device="cuda").to(dt)
    b = (torch.randn(128, 32768, device="cuda") / 180).to(dt)
    ref = a.float() @ b.float().t()
    one = ops.gemm([a], [b], out_fp32=True, split_k=1)
    many = ops.gemm([a], [b], out_fp32=True, split_k=37)
    auto = ops.gemm([a], [b], out_fp32=True)
    for o in (one, many, auto):
        assert _rel(o, ref)[0] < 2e-3
    acc = torch.ones(320, 128, device="cuda")
    ops.gemm([a], [b], out=acc, split_k=8, accumulate=True)
    assert _rel(acc, ref + 1)[0] < 2e-3
    # conv at 8x8 with the fused epilogue applied by the reduce pass
    x = torch.randn(8, 8, 8, 1280, device="cuda").to(dt)
    w = (torch.randn(640, 1280, 3, 3, device="cuda") / (9 * 1280) ** 0.5).to(dt)
    wk = w.permute(0, 2, 3, 1).reshape(640, -1).contiguous()
    bias = torch.randn(640, device="cuda")
    res = torch.randn(8, 8, 8, 640, device="cuda").to(dt)
    temb = torch.randn(8, 640, device="cuda")
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1) + temb[:, None, None, :] + res.float()
    for sk in (1, 5, 0):
        out = ops.gemm([x], [wk], bias=bias, rowvec=temb, rows_per_group=64, residual=res, conv_taps=ops.TAPS_3x3, split_k=sk)
        assert _rel(out.float(), ref)[0] < 2e-3, sk


@pytest.mark.parametrize("M,N,K,bn,split", [
    (256, 256, 64, 256, 1), (512, 512, 256, 256, 1), (384, 256, 128, 256, 1), (40 * 128, 1024, 320, 256, 1),
    (301 * 128 + 17, 320, 192, 160, 1), (2048, 1280, 1280, 160, 1), (1024, 640, 2560, 128, 1), (130, 96, 64, 128, 1),
    (512, 1280, 2304, 256, 3), (8192, 2560, 320, 256, 1),
])
def test_pair_kernel_exact(M, N, K, bn, split):
    """CTA-pair kernel (cta_group::2: 256 x BN tiles over two SMs): small-integer data, exact results.  Covers an odd number of
    m-tiles (the peer CTA's tile is out of bounds), N not a multiple of the tile, several pair-tiles per cluster (both TMEM
    accumulator buffers wrap), split-K, bias."""
    from comat_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randint(-2, 3, (M, K), device="cuda", generator=g).half()
    b = torch.randint(-2, 3, (N, K), device="cuda", generator=g).half()
    a = a * (torch.rand(M, K, device="cuda", generator=g) < 0.5)
    bias = torch.randint(-4, 5, (N,), device="cuda", generator=g).float()
    out = ops.gemm([a], [b], bias=bias, force_bn=bn, kernel="pair", split_k=split)
    ref = a.float() @ b.float().t() + bias
    assert ref.abs().max() < 2048
    assert torch.equal(out.float(), ref), (out.float() - ref).abs().nonzero()[:8]


@pytest.mark.parametrize("n,H,W,C,Cout,bn", [(2, 64, 64, 320, 320, 160), (3, 32, 32, 640, 320, 160), (2, 16, 16, 1280, 640, 128),
                                             (3, 8, 8, 1280, 1280, 256), (1, 128, 128, 128, 128, 128), (2, 24, 40, 64, 64, 128)])
def test_pair_kernel_conv_and_fused_epilogue(n, H, W, C, Cout, bn):
    from comat_b200 import ops
    torch.manual_seed(H + C)
    dtype = torch.float16
    x = torch.randn(n, H, W, C, device="cuda").to(dtype)
    w = (torch.randn(Cout, C, 3, 3, device="cuda") / (9 * C) ** 0.5).to(dtype)
    bias = torch.randn(Cout, device="cuda")
    res = torch.randn(n, H, W, Cout, device="cuda").to(dtype)
    temb = torch.randn(n, Cout, device="cuda")
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).contiguous()
    out = ops.gemm([x], [wk], bias=bias, conv_taps=ops.TAPS_3x3, rowvec=temb, rows_per_group=H * W, residual=res, force_bn=bn, kernel="pair")
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1) + temb[:, None, None, :] + res.float()
    l2, mx = _rel(out.float(), ref)
    assert l2 < 2e-3, (l2, mx)
    # two K segments (LoRA branch / fused concat) through the pair kernel
    M, N, K, r = n * H * W, Cout, C, 128
    xa = x.reshape(M, K)
    w2 = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dtype)
    t = torch.randn(M, r, device="cuda").to(dtype)
    up = (torch.randn(N, r, device="cuda") * 0.05).to(dtype)
    out2 = ops.gemm([xa, t], [w2, up], act="silu", force_bn=bn, kernel="pair")
    ref2 = F.silu(xa.float() @ w2.float().t() + t.float() @ up.float().t())
    assert _rel(out2.float(), ref2)[0] < 3e-3


def test_hi_lo_split_projection():
    """fp32 G fed to the tensor cores as bf16 hi + lo, factors as bf16 hi + lo (one operand format per MMA): how the accumulated
    LoRA product gradient is projected onto the factors; ~16 mantissa bits at fp32 range."""
    from comat_b200 import ops, _lib
    g = torch.Generator(device="cuda").manual_seed(11)
    N, K, r = 320, 640, 128
    G = torch.randn(N, K, device="cuda", generator=g) * 3.7e3                  # loss-scaled magnitudes, beyond fp16 comfort
    hi, lo = torch.empty(N * K, device="cuda", dtype=torch.bfloat16), torch.empty(N * K, device="cuda", dtype=torch.bfloat16)
    ops.split_f32_bf16x2(G.reshape(-1), hi, lo, 0.25)
    rec = hi.float() + lo.float()
    assert float(((rec - 0.25 * G.reshape(-1)).abs() / (0.25 * G.reshape(-1)).abs().clamp_min(1e-3)).max()) < 2 ** -15
    hi, lo = hi.view(N, K), lo.view(N, K)
    down = torch.randn(r, K, device="cuda", generator=g) / r ** 0.5
    up = torch.randn(N, r, device="cuda", generator=g) * 0.05
    dh, uh = down.bfloat16(), up.bfloat16()
    dl, ul = (down - dh.float()).bfloat16(), (up - uh.float()).bfloat16()
    g_up = torch.ones(N, r, device="cuda")
    ops.gemm([hi, lo], [dh, dh], out=g_up, split_k=2, accumulate=True)
    ops.gemm([hi], [dl], out=g_up, split_k=2, accumulate=True)
    ref_up = 0.25 * G.double() @ down.double().t()
    assert _rel(g_up - 1, ref_up)[0] < 5e-5
    g_down = torch.zeros(r, K, device="cuda")
    ops.gemm_tn(uh, hi, accumulate_into=g_down)
    ops.gemm_tn(uh, lo, accumulate_into=g_down)
    ops.gemm_tn(ul, hi, accumulate_into=g_down)
    ref_down = up.double().t() @ (0.25 * G.double())
    assert _rel(g_down, ref_down)[0] < 5e-5
    with pytest.raises(_lib.ComatError):                                      # one operand format per MMA
        ops.gemm([hi], [dh.half()], out_fp32=True)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,inner,K,kernel", [(8192, 1280, 320, None), (2048, 5120, 1280, None), (300, 320, 320, None), (1024, 2560, 640, "tile"),
                                              (4096, 1280, 320, "persist"), (4096, 1280, 640, "pair"), (130, 48, 64, None)])
def test_gemm_fused_geglu_epilogue(M, inner, K, kernel, dtype):
    """act='geglu': the projection's weight rows are interleaved (2j = hidden_j, 2j+1 = gate_j) and the epilogue writes
    hidden * gelu(gate) as a half-width tensor (diffusers GEGLU.forward) - vs the unfused fp32 computation."""
    from comat_b200 import ops
    torch.manual_seed(M + inner + K)
    a = torch.randn(M, K, device="cuda").to(dtype)
    w = (torch.randn(2 * inner, K, device="cuda") / K ** 0.5).to(dtype)
    b = torch.randn(2 * inner, device="cuda") * 0.1
    w_il = torch.stack([w[:inner], w[inner:]], 1).reshape(2 * inner, K).contiguous()
    b_il = torch.stack([b[:inner], b[inner:]], 1).reshape(-1).contiguous()
    out = ops.gemm([a], [w_il], bias=b_il, act="geglu", kernel=kernel)
    assert out.shape == (M, inner)
    hg = a.float() @ w.float().t() + b
    ref = hg[:, :inner] * F.gelu(hg[:, inner:])
    l2, mx = _rel(out.float(), ref)
    tol = 2e-3 if dtype == torch.float16 else 1.2e-2
    assert l2 < tol and mx < 4 * tol, (l2, mx)


def _gn_ref_sums(out, n_img, G):
    """(sum, sum of squares) per (image, group) of the stored 16-bit result, in fp64"""
    C = out.shape[-1]
    y = out.double().reshape(n_img, -1, G, C // G)
    return torch.stack([y.sum((1, 3)), (y * y).sum((1, 3))], -1)


@pytest.mark.parametrize("n,H,W,C,Cout,kernel,bn,why", [
    (2, 32, 32, 320, 320, None, 0, "cpg 10: groups straddle the 32-column chunks"),
    (2, 32, 32, 320, 320, "tile", 160, "one-tile kernel (4 epilogue warps)"),
    (2, 32, 32, 320, 320, "persist", 128, "BN=128: groups straddle N tiles (128 % 10 != 0)"),
    (4, 16, 16, 640, 640, "pair", 160, "CTA-pair kernel, cpg 20"),
    (4, 16, 16, 640, 1280, "pair", 256, "CTA-pair kernel BN=256, cpg 40"),
    (6, 8, 8, 1280, 1280, "tile", 0, "8x8 maps: two images per 128-pixel tile"),
    (3, 24, 24, 128, 256, None, 0, "W=24 < TW=32: out-of-bounds tile rows must not count"),
    (1, 64, 64, 320, 320, None, 0, "several tiles per image"),
])
def test_conv_epilogue_groupnorm_statistics(n, H, W, C, Cout, kernel, bn, why):
    """NS-1: the GroupNorm statistics of a conv's output come out of its epilogue (comat_gemm_params.gn_sums) and the
    one-pass GroupNorm built on them matches the two-pass kernel and torch's group_norm on the same tensor."""
    from comat_b200 import ops
    torch.manual_seed(n * H + C)
    x = torch.randn(n, H, W, C, device="cuda").half()
    w = (torch.randn(Cout, 9 * C, device="cuda") / (9 * C) ** 0.5).half()
    bias = torch.randn(Cout, device="cuda") * 0.5 + 0.3            # non-zero mean: exercises E[x^2] - mean^2
    rowvec = torch.randn(n, Cout, device="cuda")
    res = torch.randn(n * H * W, Cout, device="cuda").half()
    G = 32
    ops.gn_arena_reset(x.device)
    out, sums = ops.gemm([x], [w], conv_taps=ops.TAPS_3x3, bias=bias, rowvec=rowvec, rows_per_group=H * W, residual=res,
                         kernel=kernel, force_bn=bn, split_k=1, gn=(G, H * W))
    assert sums is not None, why
    plain = ops.gemm([x], [w], conv_taps=ops.TAPS_3x3, bias=bias, rowvec=rowvec, rows_per_group=H * W, residual=res,
                     kernel=kernel, force_bn=bn, split_k=1)
    assert torch.equal(out, plain)                                 # the statistics do not touch the result
    ref = _gn_ref_sums(out, n, G)
    got = sums.double().reshape(n, G, 2)
    cnt = H * W * (Cout // G)
    # sums of the fp32 accumulators vs sums of their fp16 roundings: 2^-11 relative per element, averaging down with the count
    assert ((got[..., 0] - ref[..., 0]).abs() / cnt).max() < 2e-4, why
    assert ((got[..., 1] - ref[..., 1]).abs() / ref[..., 1]).max() < 1e-3, why
    gamma, beta = torch.randn(Cout, device="cuda"), torch.randn(Cout, device="cuda")
    y1, mr1 = ops.groupnorm_fwd_from_sums(out, sums, gamma, beta, G, 1e-5, True)
    y2, mr2 = ops.groupnorm_fwd(out, gamma, beta, G, 1e-5, True)
    yt = F.silu(F.group_norm(out.float().permute(0, 3, 1, 2), G, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert _rel(y1.float(), yt)[0] < 2e-3 and _rel(y1.float(), y2.float())[0] < 1e-3, why
    assert _rel(mr1, mr2)[0] < 1e-3, why


@pytest.mark.parametrize("n,L,K,N,why", [(8, 4096, 320, 320, "proj_out at the 64x64 level (persistent kernel, K = 320)"),
                                         (4, 1024, 640, 640, "32x32 level"), (4, 64, 1280, 1280, "8x8 level: 64 rows per image"),
                                         (2, 256, 2048, 1280, "pair kernel (long K)")])
def test_linear_epilogue_groupnorm_statistics(n, L, K, N, why):
    from comat_b200 import ops
    torch.manual_seed(L + K)
    a = torch.randn(n * L, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    res = (torch.randn(n * L, N, device="cuda") + 0.7).half()
    ops.gn_arena_reset(a.device)
    out, sums = ops.gemm([a], [w], residual=res, split_k=1, gn=(32, L))
    assert sums is not None, why
    ref = _gn_ref_sums(out.reshape(n, L, N), n, 32)
    got = sums.double().reshape(n, 32, 2)
    assert ((got[..., 0] - ref[..., 0]).abs() / (L * N // 32)).max() < 2e-4, why
    assert ((got[..., 1] - ref[..., 1]).abs() / ref[..., 1]).max() < 1e-3, why


def test_epilogue_groupnorm_statistics_refused_where_unsupported():
    """split-K problems, fp32 outputs and images of fewer than 32 rows cannot carry the statistics: gemm() hands back sums=None
    and the caller keeps the two-pass GroupNorm"""
    from comat_b200 import ops
    torch.manual_seed(0)
    a = torch.randn(256, 4096, device="cuda").half()
    w = (torch.randn(320, 4096, device="cuda") / 64).half()
    out, sums = ops.gemm([a], [w], split_k=4, gn=(32, 64))
    assert sums is None and _rel(out.float(), a.float() @ w.float().t())[0] < 2e-3
    a2, w2 = a[:, :320].contiguous(), w[:, :320].contiguous()
    out, sums = ops.gemm([a2], [w2], out_fp32=True, gn=(32, 64))
    assert sums is None
    out, sums = ops.gemm([a2], [w2], gn=(32, 16))            # 16 rows per image: a 32-row accumulator quarter spans two images
    assert sums is None
