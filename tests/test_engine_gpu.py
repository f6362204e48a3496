"""GPU parity: explicit forward/backward executors (UNet, VAE decoder) vs the fp32 oracle modules with torch autograd."""
import pytest
import torch

from oracle import sd_modules as sdm

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _tiny(sdxl=False):
    torch.manual_seed(3)
    unet = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(sdxl=sdxl, width=64, cross_attention_dim=64))
    unet.requires_grad_(False)
    params = sdm.install_lora(unet, 8, up_std=0.05, seed=4)
    return unet.cuda(), params


@pytest.mark.parametrize("impl", ["product", "explicit"])
@pytest.mark.parametrize("dtype,tol", [(torch.float16, 8e-3), (torch.bfloat16, 5e-2)])
def test_unet_engine_fwd_bwd_vs_oracle(dtype, tol, impl):
    from comat_b200 import engine as E, ops
    unet, _ = _tiny()
    eng = E.UNetEngine(unet, dtype)
    eng.lora_train_impl = impl
    g = torch.Generator().manual_seed(0)
    n, hw = 2, 32
    x = torch.randn(n, 4, hw, hw, generator=g).cuda()
    ctx = torch.randn(n, 77, 64, generator=g).cuda()
    t = torch.tensor(951, device="cuda")
    dy = torch.randn(n, 4, hw, hw, generator=g).cuda()
    # oracle (fp32, autograd)
    xr = x.clone().requires_grad_(True)
    params = [p for p in unet.parameters() if p.requires_grad]
    out_ref = unet(xr, t, ctx, return_dict=False)[0]
    grads_ref = torch.autograd.grad(out_ref, [xr] + params, dy)
    # engine
    tape = E.Tape()
    xv = E.Var(ops.latent_to_nhwc(x, dtype, 64))
    out = eng.forward(tape, xv, t, ctx.to(dtype))
    out_nchw = ops.nhwc_to_nchw_f32(out.v, 4)
    assert rel(out_nchw, out_ref) < tol
    out.g = dy.permute(0, 2, 3, 1).contiguous().to(dtype)
    tape.backward()
    dx = ops.nhwc_to_nchw_f32(xv.g, 4)
    assert rel(dx, grads_ref[0]) < 3 * tol
    eg = eng.finalize_lora_grads(1.0, into_param_grads=False) if impl == "product" else eng.lora_grads()
    assert len(eg) == len(params)
    worst = max(rel(a, b) for a, b in zip(eg, grads_ref[1:]))
    assert worst < 6 * tol, worst
    if impl == "product":
        # accumulation over passes + projection straight into param.grad (what the trainer does)
        for p_ in params:
            p_.grad = torch.zeros_like(p_)
        for _ in range(2):
            tape = E.Tape()
            xv = E.Var(ops.latent_to_nhwc(x, dtype, 64))
            out = eng.forward(tape, xv, t, ctx.to(dtype))
            out.g = dy.permute(0, 2, 3, 1).contiguous().to(dtype)
            tape.backward()
        assert eng.G_dirty
        eng.finalize_lora_grads(0.5, into_param_grads=True)
        assert not eng.G_dirty and float(eng._G.abs().max()) == 0.0
        worst = max(rel(p_.grad, b) for p_, b in zip(params, grads_ref[1:]))
        assert worst < 6 * tol, worst


def test_unet_engine_sdxl_geometry_and_capture():
    from comat_b200 import engine as E, ops
    unet, _ = _tiny(sdxl=True)
    dtype = torch.float16
    eng = E.UNetEngine(unet, dtype)
    g = torch.Generator().manual_seed(1)
    n, hw = 2, 32
    x = torch.randn(n, 4, hw, hw, generator=g).cuda()
    ctx = torch.randn(n, 77, 64, generator=g).cuda()
    added = dict(text_embeds=torch.randn(n, 16, generator=g).cuda(),
                 time_ids=torch.tensor([[512., 512, 0, 0, 512, 512]] * n).cuda())
    t = torch.tensor(501, device="cuda")
    ref = unet(x, t, ctx, added_cond_kwargs=added, return_dict=False)[0]
    cap = E.AttnCapture(["up_8", "up_16"])
    tape = E.Tape()
    out = eng.forward(tape, E.Var(ops.latent_to_nhwc(x, dtype, 64)), t, ctx.to(dtype), capture=cap, added_cond=added)
    assert rel(ops.nhwc_to_nchw_f32(out.v, 4), ref) < 8e-3
    maps, _ = cap.attn_dict()
    assert cap.count == len(unet.attn_processors)
    assert set(maps) == {"up_16", "up_8"}
    for k, v in maps.items():
        for m in v:
            assert m.dtype == torch.float32 and abs(float(m.sum(-1).mean()) - 1) < 1e-3


def test_vae_decoder_engine_fwd_bwd_vs_oracle():
    from comat_b200 import engine as E, ops
    torch.manual_seed(5)
    vae = sdm.AutoencoderKL(block_out_channels=(64, 64, 128, 128)).cuda()
    vae.requires_grad_(False)
    dtype = torch.float16
    eng = E.VAEDecoderEngine(vae, dtype)
    z = torch.randn(2, 4, 16, 16, device="cuda")
    zr = z.clone().requires_grad_(True)
    ref = vae.decode(zr / vae.config.scaling_factor, return_dict=False)[0]
    dy = torch.randn_like(ref)
    gref = torch.autograd.grad(ref, zr, dy)[0]
    tape = E.Tape()
    zv = E.Var(ops.latent_to_nhwc(z, dtype, 64, 1.0 / vae.config.scaling_factor))
    out = eng.forward(tape, zv)
    img = ops.nhwc_to_nchw_f32(out.v, 3)
    assert rel(img, ref) < 8e-3
    out.g = dy.permute(0, 2, 3, 1).contiguous().to(dtype)
    tape.backward()
    dz = ops.nhwc_to_nchw_f32(zv.g, 4, 1.0 / vae.config.scaling_factor)
    assert rel(dz, gref) < 3e-2


@pytest.mark.parametrize("sdxl", [False, True])
def test_unet_engine_folded_lora_modes_vs_oracle(sdxl):
    """'merged' (no tape: LoRA folded into the weights, fused q|k|v / k|v GEMMs, cached context projections) and 'frozen'
    (taped, data gradient only) against the fp32 oracle's explicit-LoRA module; a LoRA update invalidates the folded weights."""
    from comat_b200 import engine as E, ops
    unet, _ = _tiny(sdxl)
    dtype, tol = torch.float16, 8e-3
    eng = E.UNetEngine(unet, dtype)
    g = torch.Generator().manual_seed(2)
    n, hw = 2, 32
    x = torch.randn(n, 4, hw, hw, generator=g).cuda()
    ctx = torch.randn(n, 77, 64, generator=g).cuda()
    added = dict(text_embeds=torch.randn(n, 16, generator=g).cuda(),
                 time_ids=torch.tensor([[512., 512, 0, 0, 512, 512]] * n).cuda()) if sdxl else None
    kw = dict(added_cond_kwargs=added) if sdxl else {}
    t = torch.tensor(501, device="cuda")
    dy = torch.randn(n, 4, hw, hw, generator=g).cuda()
    xr = x.clone().requires_grad_(True)
    ref = unet(xr, t, ctx, return_dict=False, **kw)[0]
    gx_ref = torch.autograd.grad(ref, xr, dy)[0]
    ctx16 = ctx.to(dtype)
    out = eng.forward(None, E.Var(ops.latent_to_nhwc(x, dtype, 64), False), t, ctx16, added_cond=added)
    assert rel(ops.nhwc_to_nchw_f32(out.v, 4), ref) < tol
    out_kv = eng.forward(None, E.Var(ops.latent_to_nhwc(x, dtype, 64), False), t, ctx16, added_cond=added, cross_kv=eng.cross_kv(ctx16))
    assert rel(out_kv.v, out.v) < 2e-3          # not bitwise: GroupNorm partial sums use shared-memory float atomics
    tape = E.Tape()
    xv = E.Var(ops.latent_to_nhwc(x, dtype, 64))
    outf = eng.forward(tape, xv, t, ctx16, added_cond=added, lora_mode="frozen")
    assert rel(ops.nhwc_to_nchw_f32(outf.v, 4), ref) < tol
    outf.g = dy.permute(0, 2, 3, 1).contiguous().to(dtype)
    eng.zero_lora_grads()
    tape.backward()
    assert rel(ops.nhwc_to_nchw_f32(xv.g, 4), gx_ref) < 3 * tol
    assert all(gr is None for gr in eng.lora_grads())
    with torch.no_grad():
        for p in eng.lora_params():
            p.mul_(1.5)
    eng.refresh_lora()
    ref2 = unet(x, t, ctx, return_dict=False, **kw)[0]
    assert rel(ref2, ref) > 2 * tol
    out = eng.forward(None, E.Var(ops.latent_to_nhwc(x, dtype, 64), False), t, ctx16, added_cond=added)
    assert rel(ops.nhwc_to_nchw_f32(out.v, 4), ref2) < tol


def test_engine_unet_module_graph_and_eager_share_context_cache():
    """EngineUNet no-grad call: CUDA-graph replay and eager execution agree; the cached context projections follow the
    caller's tensor (identity + in-place version) and the LoRA version."""
    from comat_b200.modules import EngineUNet
    unet, _ = _tiny()
    mod = EngineUNet(unet, torch.float16)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 32, 32, generator=g).cuda()
    ctx = torch.randn(2, 77, 64, generator=g).cuda()
    ctx_b = torch.randn(2, 77, 64, generator=g).cuda()
    t = torch.tensor(301, device="cuda")
    with torch.no_grad():
        e0 = mod(x, t, encoder_hidden_states=ctx)[0].clone()
        e0b = mod(x, t, encoder_hidden_states=ctx_b)[0].clone()
        mod.use_graphs = True
        assert rel(e0b, e0) > 5e-2
        for c, want in ((ctx, e0), (ctx, e0), (ctx_b, e0b), (ctx, e0)):
            assert rel(mod(x, t, encoder_hidden_states=c)[0], want) < 2e-3
        ctx.mul_(0.5)                                       # in-place edit of the caller's tensor: version changes
        e1 = mod(x, t, encoder_hidden_states=ctx)[0].clone()
        mod.use_graphs = False
        assert rel(mod(x, t, encoder_hidden_states=ctx)[0], e1) < 2e-3
        assert rel(e1, e0) > 2e-2
        for p in mod.lora_parameters():
            p.mul_(1.5)
        mod.refresh_lora()
        e2 = mod(x, t, encoder_hidden_states=ctx)[0].clone()
        mod.use_graphs = True
        assert rel(mod(x, t, encoder_hidden_states=ctx)[0], e2) < 2e-3
        assert rel(e2, e1) > 2e-2
    ref = unet(x, t, ctx, return_dict=False)[0]
    assert rel(e2, ref) < 8e-3


def test_engine_unet_module_graph_sdxl_conditioning():
    """SDXL geometry: the CUDA-graphed no-grad forward takes the pooled-text / time-id conditioning as static inputs and
    follows their values call by call."""
    from comat_b200.modules import EngineUNet
    unet, _ = _tiny(sdxl=True)
    mod = EngineUNet(unet, torch.float16)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 4, 32, 32, generator=g).cuda()
    ctx = torch.randn(2, 77, 64, generator=g).cuda()
    t = torch.tensor(301, device="cuda")
    mk = lambda: dict(text_embeds=torch.randn(2, 16, generator=g).cuda(), time_ids=torch.tensor([[512., 512, 0, 0, 512, 512]] * 2).cuda())
    a1, a2 = mk(), mk()
    with torch.no_grad():
        e1 = mod(x, t, encoder_hidden_states=ctx, added_cond_kwargs=a1)[0].clone()
        e2 = mod(x, t, encoder_hidden_states=ctx, added_cond_kwargs=a2)[0].clone()
        assert rel(e2, e1) > 1e-2
        mod.use_graphs = True
        for a, want in ((a1, e1), (a2, e2), (a2, e2), (a1, e1)):
            assert rel(mod(x, t, encoder_hidden_states=ctx, added_cond_kwargs=a)[0], want) < 2e-3
        assert len(mod._graphs) == 1
    ref = unet(x, t, ctx, added_cond_kwargs=a1, return_dict=False)[0]
    assert rel(e1, ref) < 8e-3


def test_full_size_sd15_unet_vs_oracle():
    """BASELINE geometry: the 859.5 M-parameter SD1.5 UNet (LoRA r = 128) at a 64x64 latent, CFG batch of 2 - engine (fp16,
    tcgen05 kernels, LoRA folded for the no-grad call / product-form gradients for the taped call) against the fp32 oracle
    module with torch autograd on the same weights: noise prediction, input gradient and every LoRA gradient."""
    from comat_b200 import engine as E, ops
    torch.manual_seed(42)
    with torch.device("cuda"):
        unet = sdm.UNet2DConditionModel(**sdm.SD15_UNET_CONFIG)
    unet.requires_grad_(False)
    params = sdm.install_lora(unet, 128, up_std=0.02, seed=1)
    unet = unet.cuda()
    params = [p for p in unet.parameters() if p.requires_grad]
    assert sum(p.numel() for p in unet.parameters() if not p.requires_grad) == 859_520_964
    g = torch.Generator().manual_seed(5)
    n, hw = 2, 64
    x = torch.randn(n, 4, hw, hw, generator=g).cuda()
    ctx = torch.randn(n, 77, 768, generator=g).cuda()
    dy = torch.randn(n, 4, hw, hw, generator=g).cuda()
    t = torch.tensor(601, device="cuda")
    xr = x.clone().requires_grad_(True)
    ref = unet(xr, t, ctx, return_dict=False)[0]
    grads_ref = torch.autograd.grad(ref, [xr] + params, dy)
    dtype = torch.float16
    eng = E.UNetEngine(unet, dtype)
    ctx16 = ctx.to(dtype)
    out = eng.forward(None, E.Var(ops.latent_to_nhwc(x, dtype, 64), False), t, ctx16, cross_kv=eng.cross_kv(ctx16))
    assert rel(ops.nhwc_to_nchw_f32(out.v, 4), ref) < 1.5e-2
    tape = E.Tape()
    xv = E.Var(ops.latent_to_nhwc(x, dtype, 64))
    out = eng.forward(tape, xv, t, ctx16)
    assert rel(ops.nhwc_to_nchw_f32(out.v, 4), ref) < 1.5e-2
    S = 1.0                                                    # dy ~ N(0, 1): no loss scaling needed for this check
    out.g = (dy * S).permute(0, 2, 3, 1).contiguous().to(dtype)
    tape.backward()
    assert rel(ops.nhwc_to_nchw_f32(xv.g, 4, 1.0 / S), grads_ref[0]) < 3e-2
    eg = eng.finalize_lora_grads(1.0 / S, into_param_grads=False)
    errs = [rel(a, b) for a, b in zip(eg, grads_ref[1:])]
    assert len(errs) == 256 and max(errs) < 0.1 and sum(errs) / len(errs) < 4e-2, (max(errs), sum(errs) / len(errs))
    got = torch.cat([a.reshape(-1) for a in eg]).double()
    want = torch.cat([b.reshape(-1) for b in grads_ref[1:]]).double()
    assert float((got * want).sum() / (got.norm() * want.norm())) > 0.998


@pytest.mark.parametrize("with_capture", [False, True])
def test_engine_unet_graphed_taped_calls_match_eager(with_capture):
    """graph_taped: the back-propagated UNet call replayed from a (forward graph, backward graph) pair - noise prediction, exported
    cross-attention probabilities, input gradient and the LoRA gradients (product arena -> finalize_lora_grads) equal the eager
    taped call's, call after call with new inputs; two calls in flight use two instances."""
    from comat_b200 import engine as E
    from comat_b200.modules import EngineUNet
    unet, _ = _tiny()
    g = torch.Generator().manual_seed(11)
    ins = [(torch.randn(2, 4, 32, 32, generator=g).cuda(), torch.randn(2, 77, 64, generator=g).cuda(), torch.tensor(t, device="cuda"),
            torch.randn(2, 4, 32, 32, generator=g).cuda()) for t in (801, 401, 1, 601)]

    def run(graphed):
        mod = EngineUNet(unet, torch.float16)
        for p in mod.lora_parameters():
            p.grad = torch.zeros_like(p)
        mod.direct_lora_grads, mod.graph_taped = True, graphed
        if with_capture:
            mod.capture = E.AttnCapture(["up_8", "up_16"])
        res = []
        for k in range(0, len(ins), 2):                      # two taped calls, then both backward passes (K > 1 steps in flight)
            mod.new_step()
            outs, leaves = [], []
            for x, ctx, t, w in ins[k:k + 2]:
                xr = x.clone().requires_grad_(True)
                eps = mod(xr, t, encoder_hidden_states=ctx)[0]
                loss = (eps * w).sum()
                if with_capture:
                    maps, _ = mod.capture.attn_dict()
                    loss = loss + sum((m * m).sum() for v in maps.values() for m in v) * 1e-2
                    res.append(torch.cat([m.reshape(-1) for v in maps.values() for m in v]).detach().clone())
                outs.append(loss)
                leaves.append(xr)
                res.append(eps.detach().clone())
            for loss, xr in zip(reversed(outs), reversed(leaves)):
                loss.backward()
                res.append(xr.grad.clone())
            mod.finalize_lora_grads()
            res.append(torch.cat([p.grad.reshape(-1) for p in mod.lora_parameters()]).clone())
            for p in mod.lora_parameters():
                p.grad.zero_()
        n_inst = sum(len(s["inst"]) for s in mod._taped.values())
        return res, n_inst
    ref, n0 = run(False)
    got, n1 = run(True)
    assert n0 == 0 and n1 == 2                               # first pair eager (its backward marks the signature ready), second pair: two captured instances
    # same kernels, same order - but fp32 atomics (split-K gradient accumulation, GroupNorm partial sums) reorder sums from run to
    # run and a flipped fp16 rounding early in the network moves the outputs by ~1e-3 and the gradients by a few 1e-3
    for a, b in zip(got, ref):
        assert rel(a, b) < 1.5e-2, rel(a, b)


def test_bf16_engine_sees_a_young_lora_adapter():
    """folding the adapter into bf16 weights (W' = round(W + up.down)) rounds a young adapter away - a delta below half an
    ulp of W (2^-9 |W|) never reaches the output.  bf16 engines therefore run the explicit branch y = W x + up(down(x)) in every pass:
    on one projection, the output change caused by a small adapter equals x (up.down)^T, while the folded weight loses most of it."""
    from comat_b200 import engine as E, ops
    unet, _ = _tiny()
    eng = E.UNetEngine(unet, torch.bfloat16)
    assert not eng.fold_lora and eng.lora_train_impl == "explicit"
    a = eng._attns[0]
    lora = a.loras[0]
    g = torch.Generator().manual_seed(8)
    x = torch.randn(4096, a.q.k, generator=g).cuda().bfloat16()
    with torch.no_grad():
        up0 = lora.up.detach().clone()
        lora.up.zero_()
        eng.refresh_lora()
        y0 = E.linear(None, E.Var(x, False), a.q, lora).v.float()
        a.refresh_merged(False)
        f0 = ops.gemm([x], [a.mq.w]).float()
        lora.up.copy_(torch.randn(lora.up.shape, generator=g).cuda() * 2e-4)          # |up.down| ~ 1e-4 |W|-ish: a few optimiser steps old
        eng.refresh_lora()
        y1 = E.linear(None, E.Var(x, False), a.q, lora).v.float()
        a.refresh_merged(False)
        f1 = ops.gemm([x], [a.mq.w]).float()
        want = x.float() @ (lora.up.float() @ lora.down.float()).t()
        lora.up.copy_(up0)
        eng.refresh_lora()
    cos = lambda d: float((d.double() * want.double()).sum() / (d.double().norm() * want.double().norm()).clamp_min(1e-30))
    e_cos, f_cos = cos(y1 - y0), cos(f1 - f0)
    e_ratio, f_ratio = float((y1 - y0).norm() / want.norm()), float((f1 - f0).norm() / want.norm())
    print(f"[measured] bf16 young adapter on one projection: explicit cos {e_cos:.3f} ratio {e_ratio:.3f}; folded cos {f_cos:.3f} ratio {f_ratio:.3f}")
    assert float(want.norm() / y0.norm()) < 2e-2          # a small perturbation ...
    assert e_cos > 0.5                                     # ... visible through the explicit branch (limited by bf16 output rounding)
    assert e_cos > f_cos + 0.1 or f_ratio < 0.5 * e_ratio  # ... and mostly lost by the fold


def test_unet_groupnorm_statistics_come_from_the_producing_gemm(monkeypatch):
    """NS-1 wiring: in a UNet call every GroupNorm whose input is the output of one GEMM / conv (norm2 of every ResBlock, norm1 of
    the down / mid ResBlocks, every Transformer2DModel.norm, conv_norm_out) takes its statistics from that GEMM's epilogue; only
    the up-path norm1's (which normalise a concatenation with the skip tensor) and shapes the epilogue refuses (split-K) run the
    statistics pass.  The result equals the all-two-pass executor's to rounding."""
    from comat_b200 import engine as E, ops
    unet, _ = _tiny()
    eng = E.UNetEngine(unet, torch.float16)
    g = torch.Generator().manual_seed(0)
    n, hw = 2, 32
    x = torch.randn(n, 4, hw, hw, generator=g).cuda()
    ctx = torch.randn(n, 77, 64, generator=g).cuda().half()
    t = torch.tensor(500, device="cuda")
    calls = {"fused": 0, "two_pass": 0}
    f0, f1 = ops.groupnorm_fwd_from_sums, ops.groupnorm_fwd
    monkeypatch.setattr(ops, "groupnorm_fwd_from_sums", lambda *a, **k: (calls.__setitem__("fused", calls["fused"] + 1), f0(*a, **k))[1])
    monkeypatch.setattr(ops, "groupnorm_fwd", lambda *a, **k: (calls.__setitem__("two_pass", calls["two_pass"] + 1), f1(*a, **k))[1])
    out = eng.forward(None, E.Var(ops.latent_to_nhwc(x, torch.float16, 64), False), t, ctx)
    n_res = len(eng._res_all)
    n_up_res = sum(len(rs) for rs, _, _ in eng.up)
    n_tr = sum(len(a) for _, a, _ in eng.down + eng.up if a is not None) + len(eng.mid[1])
    total = 2 * n_res + n_tr + 1
    assert calls["fused"] + calls["two_pass"] == total
    assert calls["two_pass"] >= n_up_res                       # concatenated inputs
    # split-K problems refuse the statistics: at this tiny geometry (width 64, 32x32 latent, n = 2) most convs have too few tiles
    # to fill the GPU and run split-K, so only a lower bound is asserted here; test_full_size_* covers the BASELINE geometry
    assert calls["fused"] >= 16, calls
    monkeypatch.setattr(E, "FUSE_GN_STATS", False)
    calls["fused"] = 0
    ref = eng.forward(None, E.Var(ops.latent_to_nhwc(x, torch.float16, 64), False), t, ctx)
    assert calls["fused"] == 0
    assert rel(out.v.float(), ref.v.float()) < 2e-3
