"""GPU parity: normalisation / activation / rearrangement kernels vs plain PyTorch fp32 references (fwd + bwd)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.2e-2)])
@pytest.mark.parametrize("n,HW,C,silu,eps", [(2, 4096, 320, True, 1e-5), (3, 1024, 640, False, 1e-6), (2, 256, 1920, True, 1e-5),
                                             (2, 64, 2560, True, 1e-5), (1, 16384, 128, True, 1e-6), (2, 64, 32, True, 1e-5),
                                             # single-launch cluster kernel: slabs larger than shared memory (rows re-read), 8-CTA
                                             # clusters, one-vector channel sets (VAE), a row count that does not divide
                                             (8, 4096, 960, True, 1e-5), (1, 65536, 256, True, 1e-6), (2, 262144, 128, False, 1e-6),
                                             (2, 100, 64, True, 1e-5), (8, 1024, 1280, False, 1e-6)])
def test_groupnorm(n, HW, C, silu, eps, dtype, tol):
    from comat_b200 import ops
    torch.manual_seed(C)
    x = (torch.randn(n, HW, C, device="cuda") * 1.5 + 0.3).to(dtype)
    g, b = torch.randn(C, device="cuda") * 0.2 + 1, torch.randn(C, device="cuda") * 0.1
    dy = torch.randn(n, HW, C, device="cuda").to(dtype)
    y, mr = ops.groupnorm_fwd(x, g, b, 32, eps, silu)
    dx = ops.groupnorm_bwd(x, dy, g, b, mr, 32, silu)
    xr = x.float().requires_grad_(True)
    yr = F.group_norm(xr.permute(0, 2, 1), 32, g, b, eps).permute(0, 2, 1)
    if silu:
        yr = F.silu(yr)
    yr.backward(dy.float())
    assert rel(y.float(), yr) < tol
    assert rel(dx.float(), xr.grad) < 2 * tol
    xs = x.float().reshape(n, HW, 32, C // 32)
    mean, var = xs.mean((1, 3)), xs.var((1, 3), unbiased=False)
    mr = mr.reshape(n, 32, 2)
    assert rel(mr[..., 0], mean) < 1e-4 + 1e-3 * float(var.sqrt().mean() / mean.abs().mean().clamp_min(1e-6)) and rel(mr[..., 1], (var + eps).rsqrt()) < 1e-3


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.2e-2)])
@pytest.mark.parametrize("rows,C", [(8192, 320), (1000, 640), (577 * 2, 1024), (77, 768), (300, 1280)])
def test_layernorm(rows, C, dtype, tol):
    from comat_b200 import ops
    torch.manual_seed(C)
    x = (torch.randn(rows, C, device="cuda") * 2 - 0.5).to(dtype)
    g, b = torch.randn(C, device="cuda") * 0.2 + 1, torch.randn(C, device="cuda") * 0.1
    dy = torch.randn(rows, C, device="cuda").to(dtype)
    y, mr = ops.layernorm_fwd(x, g, b, 1e-5)
    dx = ops.layernorm_bwd(x, dy, g, mr)
    xr = x.float().requires_grad_(True)
    yr = F.layer_norm(xr, (C,), g, b, 1e-5)
    yr.backward(dy.float())
    assert rel(y.float(), yr) < tol and rel(dx.float(), xr.grad) < 2 * tol


def test_geglu_and_unary():
    from comat_b200 import ops
    torch.manual_seed(0)
    for dtype, tol in ((torch.float16, 2e-3), (torch.bfloat16, 1.2e-2)):
        hg = torch.randn(1000, 2 * 1280, device="cuda").to(dtype)
        dy = torch.randn(1000, 1280, device="cuda").to(dtype)
        out = ops.geglu_fwd(hg)
        d = ops.geglu_bwd(hg, dy)
        r = hg.float().requires_grad_(True)
        h, g = r.chunk(2, -1)
        ref = h * F.gelu(g)
        ref.backward(dy.float())
        assert rel(out.float(), ref) < tol and rel(d.float(), r.grad) < 2 * tol
        x = torch.randn(64, 1280, device="cuda").to(dtype)
        y = torch.randn(64, 1280, device="cuda").to(dtype)
        assert rel(ops.elementwise("silu", x).float(), F.silu(x.float())) < tol
        xr = x.float().requires_grad_(True)
        F.silu(xr).backward(y.float())
        assert rel(ops.elementwise("silu_bwd", x, y).float(), xr.grad) < 2 * tol
        assert rel(ops.elementwise("add", x, y).float(), x.float() + y.float()) < tol
        assert rel(ops.elementwise("axpby", x, y, 0.5, -2.0).float(), 0.5 * x.float() - 2 * y.float()) < tol


def test_spatial_and_layout():
    from comat_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(2, 8, 12, 64, device="cuda").half()
    up = ops.spatial(x, "up2")
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    g = torch.randn_like(up)
    assert rel(ops.spatial(g, "up2_bwd").float(), g.float().reshape(2, 8, 2, 12, 2, 64).sum((2, 4))) < 2e-3
    s = ops.spatial(x, "s2d")
    ref = x.reshape(2, 4, 2, 6, 2, 64).permute(0, 1, 3, 2, 4, 5).reshape(2, 4, 6, 256)
    assert torch.equal(s, ref)
    assert torch.equal(ops.spatial(s, "d2s"), x)
    a = torch.randn(300, 320, device="cuda").half()
    assert torch.equal(ops.transpose16(a), a.t().contiguous())
    b = torch.randn(2, 8, 12, 128, device="cuda").half()
    assert torch.equal(ops.concat_channels(x, b), torch.cat([x, b], -1))
    lat = torch.randn(3, 4, 16, 16, device="cuda")
    nh = ops.latent_to_nhwc(lat, torch.float16, 64, 2.0)
    assert nh.shape == (3, 16, 16, 64) and float(nh[..., 4:].abs().max()) == 0
    assert rel(nh[..., :4].float(), (2.0 * lat).permute(0, 2, 3, 1)) < 1e-3
    back = ops.nhwc_to_nchw_f32(nh, 4, 0.5)
    assert rel(back, lat) < 1e-3


def test_stride2_conv_via_space_to_depth():
    """Downsample2D (conv3x3 stride 2 pad 1) == 2x2-tap conv over the space-to-depth input with rearranged weights."""
    from comat_b200 import ops, unet_weights
    torch.manual_seed(0)
    x = torch.randn(2, 32, 32, 64, device="cuda").half()
    w = (torch.randn(128, 64, 3, 3, device="cuda") / 24).half()
    bias = torch.randn(128, device="cuda")
    wk, taps = unet_weights.pack_conv_stride2(w)
    out = ops.gemm([ops.spatial(x, "s2d")], [wk], bias=bias, conv_taps=taps)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=2, padding=1).permute(0, 2, 3, 1)
    assert rel(out.float(), ref) < 2e-3


def test_bicubic_aa_resize_normalize_fwd_bwd():
    """concept_mat_utils/caption_blip.py:33-36,45 — native table-driven kernel vs (a) the dense fp64 operator built from the same
    taps (validated against aten on CPU in tests/test_image_ops_cpu.py) and (b) aten upsample_bicubic2d_aa on the GPU."""
    from comat_b200 import image_ops as IO
    torch.manual_seed(0)
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
    mt = torch.tensor(mean, device="cuda", dtype=torch.float64).view(1, 3, 1, 1)
    st = torch.tensor(std, device="cuda", dtype=torch.float64).view(1, 3, 1, 1)
    for size_in in (510, 254):
        x = torch.rand(2, 3, size_in, size_in, device="cuda", requires_grad=True)
        y = IO.resize_bicubic_aa_normalize(x, 384, mean, std)
        g = torch.randn_like(y)
        y.backward(g)
        M = IO.dense_operator(size_in, 384).double().cuda()
        y64 = ((M @ x.detach().double() @ M.t()) - mt) / st
        gx64 = M.t() @ (g.double() / st) @ M
        e_f, e_b = rel(y, y64), rel(x.grad, gx64)
        xr = x.detach().clone().requires_grad_(True)
        ref = (F.interpolate(xr, size=(384, 384), mode="bicubic", antialias=True, align_corners=False) - mt.float()) / st.float()
        ref.backward(g)
        e_fa, e_ba = rel(y, ref), rel(x.grad, xr.grad)
        print("resize errors (dense fwd, dense bwd, aten fwd, aten bwd):", e_f, e_b, e_fa, e_ba)
        assert e_f < 1e-5 and e_b < 1e-5, (e_f, e_b)
        assert e_fa < 1e-4 and e_ba < 1e-4, (e_fa, e_ba)


def test_label_smoothed_ce_fwd_bwd():
    from comat_b200 import blip_engine as BE
    torch.manual_seed(0)
    R, V, Vpad, eps = 37, 30524, 30528, 0.1
    logits = (torch.randn(R, V, device="cuda") * 3).contiguous()
    labels = torch.randint(0, V, (R,), device="cuda")
    labels[::5] = -100
    stats, out2 = BE.ce_fwd(logits, labels, V, eps)
    lr = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr, labels, ignore_index=-100, label_smoothing=eps)
    ref.backward()
    assert abs(float(out2[0]) - float(ref)) < 1e-5 * abs(float(ref)) and int(out2[1]) == int((labels != -100).sum())
    d = BE.ce_bwd(logits, labels, stats, out2, torch.ones(1, device="cuda"), V, Vpad, eps, torch.float16)
    assert float(d[:, V:].abs().max()) == 0
    assert rel(d[:, :V].float(), lr.grad) < 2e-3


def test_groupnorm_fused_cluster_kernel_in_subprocess():
    """COMAT_GN=fused (read once per process) routes GroupNorm to the single-launch cluster kernel (csrc/groupnorm_fused.cu): the
    same parity cases, in a child interpreter."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, COMAT_GN="fused")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_elementwise_gpu.py", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "test_groupnorm and not subprocess"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "passed" in r.stdout and "failed" not in r.stdout


@pytest.mark.parametrize("n,hw,n_zero", [(4, 64, 0), (8, 64, 4), (2, 32, 1), (3, 17, 0)])
def test_gan_head_bce_fwd_bwd_vs_torch(n, hw, n_zero):
    """comat_gan_head_bce: permute -> Linear(4, 1) -> BCEWithLogits of gan_sdxl.py:84-89 / :118-132, loss and all three gradients."""
    from comat_b200 import ops
    torch.manual_seed(n + hw)
    eps = (torch.randn(n, 4, hw, hw, device="cuda") * 2).requires_grad_(True)
    lin = torch.nn.Linear(4, 1).cuda()
    loss = ops.gan_head_bce(eps, lin.weight, lin.bias, n_zero)
    g = torch.autograd.grad(loss * 3.0, [eps, lin.weight, lin.bias])
    eps2 = eps.detach().clone().requires_grad_(True)
    pred = lin(eps2.permute(0, 2, 3, 1))
    target = torch.ones_like(pred)
    target[:n_zero] = 0
    ref = F.binary_cross_entropy_with_logits(pred, target)
    gr = torch.autograd.grad(ref * 3.0, [eps2, lin.weight, lin.bias])
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    for a, b in zip(g, gr):
        assert a.shape == b.shape and rel(a, b) < 1e-4
