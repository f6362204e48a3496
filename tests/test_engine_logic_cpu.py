"""CPU: executor logic (tape, backward formulas, packing, segments) vs the oracle, with the C-ABI ops replaced by a
torch emulation of their semantics (tests/cpu_ops_emulation.py — test infrastructure only)."""
import pytest
import torch

from oracle import sd_modules as sdm
from tests import cpu_ops_emulation as EMU


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _tiny(sdxl=False):
    torch.manual_seed(3)
    unet = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(sdxl=sdxl, width=64, cross_attention_dim=64))
    unet.requires_grad_(False)
    sdm.install_lora(unet, 8, up_std=0.05, seed=4)
    return unet


@pytest.mark.parametrize("impl", ["product", "explicit"])
@pytest.mark.parametrize("sdxl", [False, True])
def test_unet_executor_matches_oracle(monkeypatch, sdxl, impl):
    """taped forward + backward with LoRA weight gradients, both gradient forms: 'product' (G = dy^T x accumulated, projected
    onto the factors at the end) and 'explicit' (the literal LoRA branch)."""
    EMU.install(monkeypatch)
    from comat_b200 import engine as E, ops
    unet = _tiny(sdxl)
    dtype = torch.float32
    eng = E.UNetEngine(unet, dtype)
    eng.lora_train_impl = impl
    g = torch.Generator().manual_seed(0)
    n, hw = 2, 16
    x = torch.randn(n, 4, hw, hw, generator=g)
    ctx = torch.randn(n, 77, 64, generator=g)
    added = dict(text_embeds=torch.randn(n, 16, generator=g), time_ids=torch.tensor([[512., 512, 0, 0, 512, 512]] * n)) if sdxl else None
    t = torch.tensor(951)
    dy = torch.randn(n, 4, hw, hw, generator=g)
    xr = x.clone().requires_grad_(True)
    params = [p for p in unet.parameters() if p.requires_grad]
    kw = dict(added_cond_kwargs=added) if sdxl else {}
    out_ref = unet(xr, t, ctx, return_dict=False, **kw)[0]
    grads_ref = torch.autograd.grad(out_ref, [xr] + params, dy)
    cap = E.AttnCapture(["up_8", "up_16"])
    tape = E.Tape()
    xv = E.Var(ops.latent_to_nhwc(x, dtype, 64))
    out = eng.forward(tape, xv, t, ctx, capture=cap, added_cond=added)
    assert rel(ops.nhwc_to_nchw_f32(out.v, 4), out_ref) < 1e-4
    assert cap.count == len(unet.attn_processors)
    out.g = dy.permute(0, 2, 3, 1).contiguous()
    tape.backward()
    assert rel(ops.nhwc_to_nchw_f32(xv.g, 4), grads_ref[0]) < 1e-4
    eg = eng.finalize_lora_grads(1.0, into_param_grads=False) if impl == "product" else eng.lora_grads()
    assert len(eg) == len(params)
    # 'product' feeds G to the projection GEMMs as bf16 hi + lo (16 mantissa bits)
    assert max(rel(a, b) for a, b in zip(eg, grads_ref[1:])) < 1e-3
    if impl == "product":                                        # a second pass accumulates: twice the gradient
        tape = E.Tape()
        xv = E.Var(ops.latent_to_nhwc(x, dtype, 64))
        for _ in range(2):
            out = eng.forward(tape, xv, t, ctx, added_cond=added)
            out.g = dy.permute(0, 2, 3, 1).contiguous()
            tape.backward()
        eg2 = eng.finalize_lora_grads(1.0, into_param_grads=False)
        assert max(rel(a, 2 * b) for a, b in zip(eg2, grads_ref[1:])) < 1e-3
        assert float(eng._G.abs().max()) == 0.0 and not eng.G_dirty


def test_vae_executor_matches_oracle(monkeypatch):
    EMU.install(monkeypatch)
    from comat_b200 import engine as E, ops
    torch.manual_seed(5)
    vae = sdm.AutoencoderKL(block_out_channels=(64, 64, 128, 128))
    vae.requires_grad_(False)
    eng = E.VAEDecoderEngine(vae, torch.float32)
    z = torch.randn(1, 4, 8, 8)
    zr = z.clone().requires_grad_(True)
    ref = vae.decode(zr / vae.config.scaling_factor, return_dict=False)[0]
    dy = torch.randn_like(ref)
    gref = torch.autograd.grad(ref, zr, dy)[0]
    tape = E.Tape()
    zv = E.Var(ops.latent_to_nhwc(z, torch.float32, 64, 1.0 / vae.config.scaling_factor))
    out = eng.forward(tape, zv)
    assert rel(ops.nhwc_to_nchw_f32(out.v, 3), ref) < 1e-4
    out.g = dy.permute(0, 2, 3, 1).contiguous()
    tape.backward()
    assert rel(ops.nhwc_to_nchw_f32(zv.g, 4, 1.0 / vae.config.scaling_factor), gref) < 1e-4


@pytest.mark.parametrize("sdxl", [False, True])
def test_unet_merged_lora_modes_match_oracle(monkeypatch, sdxl):
    """LoRA folded into the projection weights: 'merged' (no tape: fused q|k|v and k|v GEMMs, cached context projections)
    and 'frozen' (taped, data gradients only - the discriminator's generator-side pass, gan_sdxl.py:52-89) must equal
    the oracle's explicit-LoRA forward / input gradient; a LoRA update must invalidate the folded weights."""
    EMU.install(monkeypatch)
    from comat_b200 import engine as E, ops
    unet = _tiny(sdxl)
    dtype = torch.float32
    eng = E.UNetEngine(unet, dtype)
    g = torch.Generator().manual_seed(1)
    n, hw = 2, 16
    x = torch.randn(n, 4, hw, hw, generator=g)
    ctx = torch.randn(n, 77, 64, generator=g)
    added = dict(text_embeds=torch.randn(n, 16, generator=g), time_ids=torch.tensor([[512., 512, 0, 0, 512, 512]] * n)) if sdxl else None
    kw = dict(added_cond_kwargs=added) if sdxl else {}
    t = torch.tensor(500)
    dy = torch.randn(n, 4, hw, hw, generator=g)
    xr = x.clone().requires_grad_(True)
    out_ref = unet(xr, t, ctx, return_dict=False, **kw)[0]
    gx_ref = torch.autograd.grad(out_ref, xr, dy)[0]
    # merged, context projections computed inside the call
    cap = E.AttnCapture(["up_8", "up_16"])
    out = eng.forward(None, E.Var(ops.latent_to_nhwc(x, dtype, 64), False), t, ctx, capture=cap, added_cond=added)
    assert rel(ops.nhwc_to_nchw_f32(out.v, 4), out_ref) < 1e-4
    assert cap.count == len(unet.attn_processors)
    # merged, with the cached per-layer k|v of the context
    kv = eng.cross_kv(ctx)
    assert len(kv) == len(unet.attn_processors) // 2
    out = eng.forward(None, E.Var(ops.latent_to_nhwc(x, dtype, 64), False), t, ctx, added_cond=added, cross_kv=kv)
    assert rel(ops.nhwc_to_nchw_f32(out.v, 4), out_ref) < 1e-4
    # frozen: taped, folded weights, input gradient only
    tape = E.Tape()
    xv = E.Var(ops.latent_to_nhwc(x, dtype, 64))
    out = eng.forward(tape, xv, t, ctx, added_cond=added, lora_mode="frozen")
    assert rel(ops.nhwc_to_nchw_f32(out.v, 4), out_ref) < 1e-4
    out.g = dy.permute(0, 2, 3, 1).contiguous()
    eng.zero_lora_grads()
    tape.backward()
    assert rel(ops.nhwc_to_nchw_f32(xv.g, 4), gx_ref) < 1e-4
    assert all(gr is None for gr in eng.lora_grads())
    # a LoRA update (optimiser step + refresh_lora) must be seen by the folded weights and the cached projections
    with torch.no_grad():
        for p in eng.lora_params():
            p.mul_(1.5)
    eng.refresh_lora()
    out_ref2 = unet(x, t, ctx, return_dict=False, **kw)[0]
    assert rel(out_ref2, out_ref) > 1e-3
    out = eng.forward(None, E.Var(ops.latent_to_nhwc(x, dtype, 64), False), t, ctx, added_cond=added, cross_kv=eng.cross_kv(ctx))
    assert rel(ops.nhwc_to_nchw_f32(out.v, 4), out_ref2) < 1e-4


@pytest.mark.parametrize("mode", ["product", "explicit", "frozen"])
def test_unet_executor_context_gradient_matches_oracle(monkeypatch, mode):
    """d(encoder_hidden_states): off on the CoMat path (frozen text encoders) and formed only when the caller passes the context as
    a grad-requiring Var - through the folded k / v projections in 'product' mode, through ``linear`` otherwise."""
    EMU.install(monkeypatch)
    from comat_b200 import engine as E, ops
    unet = _tiny(False)
    eng = E.UNetEngine(unet, torch.float32)
    if mode != "frozen":
        eng.lora_train_impl = mode
    g = torch.Generator().manual_seed(1)
    n, hw = 2, 16
    x, ctx = torch.randn(n, 4, hw, hw, generator=g), torch.randn(n, 77, 64, generator=g)
    t, dy = torch.tensor(400), torch.randn(n, 4, hw, hw, generator=g)
    cr = ctx.clone().requires_grad_(True)
    g_ref = torch.autograd.grad(unet(x, t, cr, return_dict=False)[0], cr, dy)[0]
    tape = E.Tape()
    cv = E.Var(ctx, needs_grad=True)
    out = eng.forward(tape, E.Var(ops.latent_to_nhwc(x, torch.float32, 64), False), t, cv, lora_mode="frozen" if mode == "frozen" else "train")
    out.g = dy.permute(0, 2, 3, 1).contiguous()
    tape.backward()
    assert cv.g is not None and cv.g.shape == ctx.shape and rel(cv.g, g_ref) < 1e-4
    # default: a plain tensor context gets no gradient work at all
    tape = E.Tape()
    out = eng.forward(tape, E.Var(ops.latent_to_nhwc(x, torch.float32, 64), False), t, ctx, lora_mode="frozen" if mode == "frozen" else "train")
    out.g = dy.permute(0, 2, 3, 1).contiguous()
    tape.backward()


def test_engine_unet_module_returns_context_gradient_to_autograd(monkeypatch):
    EMU.install(monkeypatch)
    from comat_b200.modules import EngineUNet
    unet = _tiny(False)
    mod = EngineUNet(unet, torch.float32)
    g = torch.Generator().manual_seed(2)
    x, ctx = torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 77, 64, generator=g)
    t, dy = torch.tensor(300), torch.randn(2, 4, 16, 16, generator=g)
    cr = ctx.clone().requires_grad_(True)
    g_ref = torch.autograd.grad(unet(x, t, cr, return_dict=False)[0], cr, dy)[0]
    c2 = ctx.clone().requires_grad_(True)
    eps = mod(x, t, encoder_hidden_states=c2)[0]
    (eps * dy).sum().backward()
    assert rel(c2.grad, g_ref) < 1e-3
    c3 = ctx.clone()                                             # no grad requested: the call returns None for the context
    xr = x.clone().requires_grad_(True)
    (mod(xr, t, encoder_hidden_states=c3)[0] * dy).sum().backward()
    assert c3.grad is None and xr.grad is not None


@pytest.mark.parametrize("sdxl", [False, True])
def test_groupnorm_statistics_routing_over_the_unet_graph(monkeypatch, sdxl):
    """NS-1 wiring on the executor level (emulated ops, so no split-K refusals): every GroupNorm whose input is the output of ONE GEMM / conv
    - norm2 of every ResBlock, norm1 of the down / mid ResBlocks, every Transformer2DModel.norm, conv_norm_out - takes the statistics
    its producer accumulated; exactly the up-path norm1's (input = concatenation with the skip tensor) run the statistics pass.  Turning
    the routing off changes nothing in the result."""
    EMU.install(monkeypatch)
    from comat_b200 import engine as E, ops
    unet = _tiny(sdxl)
    eng = E.UNetEngine(unet, torch.float32)
    g = torch.Generator().manual_seed(1)
    n, hw = 1, (32 if sdxl else 64)                # every level keeps >= 32 rows per image (the epilogue's condition): 8 x 8 at the bottom
    x = torch.randn(n, 4, hw, hw, generator=g)
    ctx = torch.randn(n, 77, 64, generator=g)
    added = dict(text_embeds=torch.randn(n, 16, generator=g), time_ids=torch.tensor([[512., 512, 0, 0, 512, 512]] * n)) if sdxl else None
    t = torch.tensor(400)
    calls = {"fused": 0, "two_pass": 0}
    f0, f1 = ops.groupnorm_fwd_from_sums, ops.groupnorm_fwd
    monkeypatch.setattr(ops, "groupnorm_fwd_from_sums", lambda *a, **k: (calls.__setitem__("fused", calls["fused"] + 1), f0(*a, **k))[1])
    monkeypatch.setattr(ops, "groupnorm_fwd", lambda *a, **k: (calls.__setitem__("two_pass", calls["two_pass"] + 1), f1(*a, **k))[1])
    out = eng.forward(None, E.Var(ops.latent_to_nhwc(x, torch.float32, 64), False), t, ctx, added_cond=added)
    n_res = len(eng._res_all)
    n_up_res = sum(len(rs) for rs, _, _ in eng.up)
    n_tr = sum(len(a) for _, a, _ in eng.down + eng.up if a is not None) + len(eng.mid[1])
    assert calls["fused"] + calls["two_pass"] == 2 * n_res + n_tr + 1
    assert calls["two_pass"] == n_up_res, calls
    monkeypatch.setattr(E, "FUSE_GN_STATS", False)
    calls["fused"] = calls["two_pass"] = 0
    ref = eng.forward(None, E.Var(ops.latent_to_nhwc(x, torch.float32, 64), False), t, ctx, added_cond=added)
    assert calls["fused"] == 0 and calls["two_pass"] == 2 * n_res + n_tr + 1
    assert rel(out.v, ref.v) < 1e-5
