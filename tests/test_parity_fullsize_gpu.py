"""GPU parity at BASELINE geometry (VERDICT r01 "next round" item 1): the product on the sm_100a kernels against the fp32
oracle run on the same B200 (TF32 off), same 16-bit-rounded frozen weights, same noise / steps / crop.

* the optimiser kernels vs torch.optim.AdamW + clip_grad_norm_ (row a20), incl. the overflow guard;
* the discriminator step vs the oracle (row a13);
* one generator step of BASELINE configs[1] geometry - SD1.5 859.5 M UNet, S = 20, K = 5, attrcon + GAN, BlipEngine on BLIP-large
  24 + 12 layers, B = 1 so the fp32 oracle's autograd graph fits beside the product - north-star tolerance 1e-3 relative on the
  concept-matching loss and the per-token attention loss;
* BLIP-large at full depth, the full-width VAE decoder, the full-size SDXL UNet (rows a11 / a17 / a15-SDXL) and an SDXL attrcon
  rollout (rows a3 / a5).
Measured errors are printed ([measured] ...) and appended to gpurun_out/r02_parity.jsonl when that directory exists."""
import json
import os
import random

import pytest
import torch

from oracle import comat_ref as R
from oracle import fixtures as FX
from oracle import sd_modules as sdm
from tests import parity_world as PW

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, **vals):
    print(f"[measured] {name}: " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in vals.items()))
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "r02_parity.jsonl"), "a") as fh:
            fh.write(json.dumps({"test": name, **vals}) + "\n")


@pytest.fixture(autouse=True)
def _fp32_oracle():
    """the oracle is the fp32 statement of the arithmetic: no TF32 in its convolutions / matmuls"""
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------------------- a20
def _lora_shaped_params(dev, seed=0):
    """the SD1.5 r=128 LoRA parameter set (training_utils/pipeline.py:84-121): 128 (down, up) pairs, 25.5 M fp32 values"""
    g = torch.Generator().manual_seed(seed)
    shapes = []
    for C, n_self, ctx in ((320, 5, 768), (640, 5, 768), (1280, 6, 768)):
        for _ in range(n_self):
            shapes += [(128, C), (C, 128)] * 5 + [(128, ctx), (C, 128)] * 2 + [(128, C), (C, 128)]     # attn1 q,k,v,o + attn2 q(=C),k,v,o
    return [torch.nn.Parameter((torch.randn(s, generator=g) * 0.05).to(dev)) for s in shapes]


@pytest.mark.parametrize("max_norm", [0.1, 0.0])
def test_flat_adamw_vs_torch_adamw_and_clip(max_norm):
    from comat_b200.optim import FlatAdamW
    dev = torch.device("cuda")
    ours = _lora_shaped_params(dev)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    n = sum(p.numel() for p in ours)
    hp = dict(lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    opt = FlatAdamW(ours, max_grad_norm=max_norm, **hp)
    topt = torch.optim.AdamW(ref, **hp)
    start = opt.flat.clone()
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        scale = (1e-4, 3e-3, 1e-6)[step]                       # total norms ~0.5 / ~15 (clipped at 0.1) and ~0.005 (not clipped)
        grads = [(torch.randn(p.shape, generator=g) * scale).to(dev) for p in ours]
        opt.zero_grad()
        for p, q, gr in zip(ours, ref, grads):
            p.grad.copy_(gr)
            q.grad = gr.clone()
        if max_norm > 0:
            total = torch.nn.utils.clip_grad_norm_(ref, max_norm)            # training_script.py:661
        else:
            total = torch.cat([q.grad.reshape(-1) for q in ref]).norm()
        topt.step()                                                          # :662
        opt.step()
        assert abs(float(opt.grad_norm()) - float(total)) <= 1e-5 * float(total)
    flat_ref = torch.cat([q.detach().reshape(-1) for q in ref])
    m_ref = torch.cat([topt.state[q]["exp_avg"].reshape(-1) for q in ref])
    v_ref = torch.cat([topt.state[q]["exp_avg_sq"].reshape(-1) for q in ref])
    upd = PW.rel_l2(opt.flat - start, flat_ref - start)
    record(f"adamw_clip max_norm={max_norm}", n=n, update_rel=upd, m_rel=PW.rel_l2(opt.m, m_ref), v_rel=PW.rel_l2(opt.v, v_ref),
           max_abs=float((opt.flat - flat_ref).abs().max()))
    assert n > 25_000_000 and opt.step_count == 3
    assert upd < 5e-6 and PW.rel_l2(opt.m, m_ref) < 1e-6 and PW.rel_l2(opt.v, v_ref) < 1e-6
    assert float((opt.flat - flat_ref).abs().max()) < 5e-7
    assert ours[0].data_ptr() == opt.flat.data_ptr() and ours[0].grad.data_ptr() == opt.grad.data_ptr()     # parameters live in the flat buffers


@pytest.mark.parametrize("bad", [float("inf"), float("nan")])
def test_flat_adamw_skips_a_non_finite_gradient(bad):
    """GradScaler semantics (accelerate fp16, training_script.py:659-663): an inf / NaN gradient leaves parameters, moments
    and the step count untouched; the next clean step proceeds with bias correction for step 2, not 3."""
    from comat_b200.optim import FlatAdamW
    dev = torch.device("cuda")
    ours = [torch.nn.Parameter(torch.randn(1000, 37, device=dev)), torch.nn.Parameter(torch.randn(513, device=dev))]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    hp = dict(lr=1e-2, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    opt, topt = FlatAdamW(ours, max_grad_norm=1.0, **hp), torch.optim.AdamW(ref, **hp)

    def both(gen_seed, poison=None):
        g = torch.Generator().manual_seed(gen_seed)
        grads = [torch.randn(p.shape, generator=g).to(dev) for p in ours]
        opt.zero_grad()
        for p, q, gr in zip(ours, ref, grads):
            p.grad.copy_(gr)
            q.grad = gr.clone()
        if poison is not None:
            opt.grad[4321] = poison
        else:
            torch.nn.utils.clip_grad_norm_(ref, 1.0)
            topt.step()
        opt.step()
    both(0)
    snap = (opt.flat.clone(), opt.m.clone(), opt.v.clone())
    both(1, poison=bad)
    assert torch.equal(opt.flat, snap[0]) and torch.equal(opt.m, snap[1]) and torch.equal(opt.v, snap[2])
    assert opt.step_count == 1 and opt.counters.tolist() == [1, 1, 1]
    both(2)
    assert opt.step_count == 2 and opt.counters.tolist() == [2, 1, 0]
    flat_ref = torch.cat([q.detach().reshape(-1) for q in ref])
    assert torch.isfinite(opt.flat).all() and PW.rel_l2(opt.flat, flat_ref) < 1e-6
    got = None
    for _ in range(200):                                       # the non-blocking read-back the trainer's loss-scale schedule uses
        got = opt.poll_overflow() or got
        torch.cuda.synchronize()
    assert got is not None and opt._skipped_seen == 1


# ---------------------------------------------------------------------------------------------------------------- a13
def test_discriminator_step_vs_oracle():
    w = PW.sd15_world("cuda", torch.float16, tiny=True, B=2, S=4, K=2, res=256, rank=8)
    d = PW.d_step_compare(w)
    record("D step (tiny geometry)", **d)
    assert d["D_loss"] < 2e-3 and d["grad_cos"] > 0.98 and abs(d["grad_norm_ratio"] - 1) < 0.1 and d["head_grad_rel"] < 3e-2, d


# ------------------------------------------------------------------------------------------------- a1-a14 at full size
def test_generator_step_at_baseline_geometry_vs_fp32_oracle():
    """BASELINE configs[1] geometry with B = 1: SD1.5 full UNet / VAE decoder / BLIP-large, S = 20 DDPM steps, K = 5 back-propagated,
    2 attrcon steps over ['mid_8','up_16','up_32','up_64'], GAN generator loss through the second full UNet."""
    w = PW.sd15_world("cuda", torch.float16, tiny=False, B=1, S=20, K=5, res=512, rank=128, lora_up_std=0.02)
    r = PW.g_step_compare(w)
    record("G step SD1.5 full geometry S=20 K=5 B=1", **{k: v for k, v in r.items() if not isinstance(v, dict)},
           **{f"product_{k}": v for k, v in r["product"].items()}, **{f"oracle_{k}": v for k, v in r["oracle"].items()})
    assert r["Blip"] < 1e-3 and r["token_loss"] < 1e-3, r                     # north-star tolerance
    assert r["loss"] < 1e-3 and r["pixel_loss"] < 5e-3 and r["G_loss"] < 1e-2, r
    assert r["grad_cos"] > 0.95 and abs(r["grad_norm_ratio"] - 1) < 0.15, r
    d = PW.d_step_compare(w)
    record("D step SD1.5 full geometry", **d)
    assert d["D_loss"] < 2e-3 and d["grad_cos"] > 0.95, d


# -------------------------------------------------------------------------------------------------------- a11 / a18
def test_blip_large_full_depth_loss_and_image_grad():
    from comat_b200.blip_engine import BlipEngine
    from comat_b200.caption import Blip
    model = PW.round_frozen_(R.make_blip(large=True, seed=3).cuda(), torch.float16)
    g = torch.Generator().manual_seed(5)
    B = 2
    images = torch.rand(B, 3, 510, 510, generator=g).cuda()
    ids, mask = FX.blip_token_batch(g, B, 14)
    ids, mask = ids.cuda(), mask.cuda()
    ref_img = images.clone().requires_grad_(True)
    r_ref = R.blip_score(model, ref_img, ids, mask, 4)
    (-r_ref).backward()
    img = images.clone().requires_grad_(True)
    r = Blip(BlipEngine(model, torch.float16)).score(img, None, input_ids=ids, attention_mask=mask)
    (-r).backward()
    cos = PW.cosine(img.grad, ref_img.grad)
    record("BLIP-large 24+12 layers", reward_rel=abs(float(r) - float(r_ref)) / abs(float(r_ref)), grad_cos=cos,
           grad_norm_ratio=float(img.grad.norm() / ref_img.grad.norm()))
    assert abs(float(r) - float(r_ref)) < 1e-3 * abs(float(r_ref))
    assert cos > 0.95 and abs(float(img.grad.norm() / ref_img.grad.norm()) - 1) < 0.15


# ---------------------------------------------------------------------------------------------------------------- a17
def test_vae_decoder_full_width_vs_oracle():
    from comat_b200 import engine as E, ops
    torch.manual_seed(5)
    with torch.device("cuda"):
        vae = sdm.AutoencoderKL()
    vae.requires_grad_(False)
    PW.round_frozen_(vae, torch.float16)
    assert sum(p.numel() for p in vae.parameters()) == 49_490_199
    dtype = torch.float16
    eng = E.VAEDecoderEngine(vae, dtype)
    z = torch.randn(1, 4, 64, 64, device="cuda") * 0.18215
    zr = z.clone().requires_grad_(True)
    ref = vae.decode(zr / vae.config.scaling_factor, return_dict=False)[0]
    dy = torch.randn_like(ref)
    gref = torch.autograd.grad(ref, zr, dy)[0]
    tape = E.Tape()
    zv = E.Var(ops.latent_to_nhwc(z, dtype, 64, 1.0 / vae.config.scaling_factor))
    out = eng.forward(tape, zv)
    img = ops.nhwc_to_nchw_f32(out.v, 3)
    out.g = dy.permute(0, 2, 3, 1).contiguous().to(dtype)
    tape.backward()
    dz = ops.nhwc_to_nchw_f32(zv.g, 4, 1.0 / vae.config.scaling_factor)
    record("VAE decoder full width 64^2 -> 512^2", image_rel=PW.rel_l2(img, ref), dz_rel=PW.rel_l2(dz, gref))
    assert PW.rel_l2(img, ref) < 1.5e-2 and PW.rel_l2(dz, gref) < 5e-2


# ------------------------------------------------------------------------------------------------------- a15 (SDXL)
def test_full_size_sdxl_unet_vs_oracle():
    """the 2.567 B-parameter SDXL UNet (LoRA r = 128 on 140 attention layers x 4 projections) at a 64x64 latent, n = 1."""
    from comat_b200 import engine as E, ops
    torch.manual_seed(42)
    with torch.device("cuda"):
        unet = sdm.UNet2DConditionModel(**sdm.SDXL_UNET_CONFIG)
    unet.requires_grad_(False)
    assert sum(p.numel() for p in unet.parameters()) == 2_567_463_684
    PW.round_frozen_(unet, torch.float16)
    sdm.install_lora(unet, 128, up_std=0.02, seed=1)
    unet = unet.cuda()
    params = [p for p in unet.parameters() if p.requires_grad]
    g = torch.Generator().manual_seed(5)
    n, hw = 1, 64
    x = torch.randn(n, 4, hw, hw, generator=g).cuda()
    ctx = torch.randn(n, 77, 2048, generator=g).cuda()
    added = dict(text_embeds=torch.randn(n, 1280, generator=g).cuda(), time_ids=torch.tensor([[512., 512, 0, 0, 512, 512]] * n).cuda())
    dy = torch.randn(n, 4, hw, hw, generator=g).cuda()
    t = torch.tensor(601, device="cuda")
    xr = x.clone().requires_grad_(True)
    ref = unet(xr, t, ctx, added_cond_kwargs=added, return_dict=False)[0]
    grads_ref = torch.autograd.grad(ref, [xr] + params, dy)
    ref, grads_ref = ref.detach(), [gr.detach() for gr in grads_ref]
    torch.cuda.empty_cache()
    dtype = torch.float16
    eng = E.UNetEngine(unet, dtype)
    ctx16 = ctx.to(dtype)
    out = eng.forward(None, E.Var(ops.latent_to_nhwc(x, dtype, 64), False), t, ctx16, added_cond=added, cross_kv=eng.cross_kv(ctx16))
    e_merged = PW.rel_l2(ops.nhwc_to_nchw_f32(out.v, 4), ref)
    tape = E.Tape()
    xv = E.Var(ops.latent_to_nhwc(x, dtype, 64))
    out = eng.forward(tape, xv, t, ctx16, added_cond=added)
    e_taped = PW.rel_l2(ops.nhwc_to_nchw_f32(out.v, 4), ref)
    out.g = dy.permute(0, 2, 3, 1).contiguous().to(dtype)
    tape.backward()
    e_dx = PW.rel_l2(ops.nhwc_to_nchw_f32(xv.g, 4), grads_ref[0])
    eg = eng.finalize_lora_grads(1.0, into_param_grads=False)
    errs = [PW.rel_l2(a, b) for a, b in zip(eg, grads_ref[1:])]
    cos = PW.cosine(torch.cat([a.reshape(-1) for a in eg]), torch.cat([b.reshape(-1) for b in grads_ref[1:]]))
    record("SDXL UNet full size n=1 64^2", eps_merged_rel=e_merged, eps_taped_rel=e_taped, dx_rel=e_dx, lora_grad_max_rel=max(errs),
           lora_grad_mean_rel=sum(errs) / len(errs), lora_grad_cos=cos, n_lora_tensors=len(errs))
    assert len(errs) == 1120
    assert e_merged < 2.5e-2 and e_taped < 2.5e-2 and e_dx < 5e-2
    assert cos > 0.99 and sum(errs) / len(errs) < 6e-2


# --------------------------------------------------------------------------------------------------------- a3 / a5
def test_sdxl_attrcon_rollout_and_step_vs_oracle():
    """AttrConcenTrainableSDXLPipeline on the device (tiny SDXL geometry): rollout image / latents / captured maps, the token and
    pixel losses on them and the LoRA gradients of an image + attention-map loss vs the oracle rollout (pinned to the reference's own
    AttrConcenTrainableSDXLPipeline.forward by tests/golden/sdxl_pipeline.pt)."""
    from comat_b200 import attn_loss
    from comat_b200 import pipelines as PL
    from comat_b200.modules import EngineUNet, EngineVAE
    dev, dtype = torch.device("cuda"), torch.float16
    B, S, hw = 2, 4, 32
    unet = FX.make_tiny_unet(31, rank=8, sdxl=True, width=64).to(dev)
    unet2 = FX.make_tiny_unet(31, rank=8, sdxl=True, width=64).to(dev)
    torch.manual_seed(32)
    vae = sdm.AutoencoderKL(block_out_channels=(64, 64, 128, 128), scaling_factor=0.13025).to(dev)
    vae.requires_grad_(False)
    for m in (unet, unet2, vae):
        PW.round_frozen_(m, dtype)
    g = torch.Generator().manual_seed(33)
    pe, ne = torch.randn(B, 77, 64, generator=g).to(dev), torch.randn(B, 77, 64, generator=g).to(dev)
    pooled, npooled = torch.randn(B, 16, generator=g).to(dev), torch.randn(B, 16, generator=g).to(dev)
    lat0 = torch.randn(B, 4, hw, hw, generator=g).to(dev)
    noises = [torch.randn(B, 4, hw, hw, generator=g).to(dev) for _ in range(S)]
    T, A = R.select_training_steps(S, 2, random.Random(2), 2)
    layers = ["up_8", "up_16"]
    words = [[[3, 7], [11]], [[5]]]
    mg = torch.Generator().manual_seed(3)
    masks = [[FX.random_mask(mg, hw * 8).to(dev) for _ in ws] for ws in words]
    ctrl = R.AttentionStore(layers)
    R.register_attention_control(unet, ctrl)
    ids = torch.tensor([[256., 256, 0, 0, 256, 256]]).repeat(B, 1).to(dev)
    added = {"text_embeds": torch.cat([npooled, pooled]), "time_ids": torch.cat([ids, ids])}
    img_o, lat_o, attn_o = R.rollout(unet, vae, sdm.DDPMScheduler(), pe, ne, lat0.clone(), noises, S, T, 7.5, 0.0, A, ctrl,
                                     added_cond_kwargs=added, sdxl=True, return_latents=True)
    tok_o, pix_o = R.mask_loss(attn_o, words, masks, layers, img_o.detach())
    loss_o = (img_o ** 2).mean() + 1e-1 * tok_o + 1e-2 * pix_o
    g_o = torch.autograd.grad(loss_o, [p for p in unet.parameters() if p.requires_grad], allow_unused=True)
    pipe = PL.AttrConcenTrainableSDXLPipeline(EngineVAE(vae, dtype), EngineUNet(unet2, dtype))
    PL.register_attention_control(pipe, PL.AttentionStore(layers))
    img, lat = pipe.forward(prompt=["p"] * B, height=hw * 8, width=hw * 8, training_timesteps=T, num_inference_steps=S,
                            guidance_scale=7.5, prompt_embeds=pe, negative_prompt_embeds=ne, pooled_prompt_embeds=pooled,
                            negative_pooled_prompt_embeds=npooled, latents=lat0.clone(), return_latents=True,
                            attrcon_train_steps=A, noises=noises)
    assert sorted(pipe.attn_dict) == sorted(attn_o)
    map_err = max(PW.rel_l2(a, b) for t in attn_o for k in attn_o[t] for a, b in zip(pipe.attn_dict[t][k], attn_o[t][k]))
    tok, pix = attn_loss.get_mask_loss(pipe.attn_dict, words, masks, layers)
    loss = (img ** 2).mean() + 1e-1 * tok + 1e-2 * pix
    gp = torch.autograd.grad(loss, pipe.unet.lora_parameters(), allow_unused=True)
    cat = lambda gs, ps: torch.cat([(a if a is not None else torch.zeros_like(p)).reshape(-1) for a, p in zip(gs, ps)])
    got, want = cat(gp, pipe.unet.lora_parameters()), cat(g_o, [p for p in unet.parameters() if p.requires_grad])
    res = dict(latents_rel=PW.rel_l2(lat, lat_o), image_rel=PW.rel_l2(img, img_o), maps_rel=map_err,
               token_rel=PW.rel_scalar(tok, tok_o), pixel_rel=PW.rel_scalar(pix, pix_o), grad_cos=PW.cosine(got, want),
               grad_norm_ratio=float(got.norm() / want.norm()))
    record("SDXL attrcon rollout (tiny geometry)", **res)
    assert res["latents_rel"] < 1e-2 and res["image_rel"] < 3e-2 and res["maps_rel"] < 2e-2
    assert res["token_rel"] < 1e-3 and res["pixel_rel"] < 2e-3
    assert res["grad_cos"] > 0.98 and abs(res["grad_norm_ratio"] - 1) < 0.1
