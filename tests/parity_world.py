"""TEST INFRASTRUCTURE: the product trainer and the oracle modules built on identical weights and inputs, at the tiny test
geometry or at BASELINE geometry (SD1.5 859.5 M UNet, full VAE decoder, BLIP-large 24 + 12 layers), so the same comparison
runs on CPU (emulated ops, tiny: checks this harness) and on the B200 (real kernels, full size: the north-star parity claim).

Frozen weights are rounded to the 16-bit engine dtype on BOTH sides before anything is built: the reference itself runs
``pipeline.unet.to(weight_dtype)`` / ``vae.to(weight_dtype)`` (training_utils/pipeline.py:60-65), so weight quantisation is
part of the reference's numbers, not an error of the port; the oracle then does the fp32 arithmetic on those values."""
import random

import torch

from oracle import comat_ref as R
from oracle import sd_modules as sdm


def rel_scalar(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64).cpu(), torch.as_tensor(b, dtype=torch.float64).cpu()
    return (a - b).abs().max().item() / max(1e-12, b.abs().max().item())


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def cosine(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))


def round_frozen_(module, dtype):
    if dtype == torch.float32:
        return module
    with torch.no_grad():
        for p in module.parameters():
            if not p.requires_grad:
                p.copy_(p.to(dtype).to(p.dtype))
    return module


def _oracle_unet(container, rank, dev, tiny):
    cfg = sdm.tiny_unet_config(width=64, cross_attention_dim=64) if tiny else sdm.SD15_UNET_CONFIG
    with torch.device(dev):
        o = sdm.UNet2DConditionModel(**cfg)
    o.requires_grad_(False)
    sdm.install_lora(o, rank)
    o.to(dev)
    o.load_state_dict(container.state_dict())
    return o


def sd15_world(dev, dtype, *, tiny, B, S, K, res, rank, n_attrcon=2, layers=None, blip_layers=None, seed=7, lora_up_std=0.05,
               blip_wrap=None):
    """returns dict(trainer, batch, oracle=dict(unet, vae, d, head, blip, ctrl, batch, cfg))"""
    from comat_b200 import synthetic
    from comat_b200.blip_engine import BlipEngine
    from comat_b200.caption import Blip, CaptionModelWrapper
    from comat_b200.gan import D_sd
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import AttentionStore, AttrConcenTrainableSDPipeline, register_attention_control
    from comat_b200.trainer import CoMatTrainer
    dev = torch.device(dev)
    unet_p, vae_p = synthetic.build_sd15(dev, dtype, rank=rank, seed=seed, tiny=tiny, lora_up_std=lora_up_std)
    d_p, _ = synthetic.build_sd15(dev, dtype, rank=rank, seed=seed + 1, tiny=tiny, lora_up_std=lora_up_std)
    for m in (unet_p, vae_p, d_p):
        round_frozen_(m, dtype)
    o_unet, o_d = _oracle_unet(unet_p, rank, dev, tiny), _oracle_unet(d_p, rank, dev, tiny)
    with torch.device(dev):
        o_vae = sdm.AutoencoderKL(block_out_channels=(64, 64, 128, 128)) if tiny else sdm.AutoencoderKL()
    o_vae.load_state_dict(vae_p.state_dict())
    o_vae.requires_grad_(False)
    blip_model = round_frozen_(R.make_blip(large=not tiny, layers=blip_layers, seed=seed + 2).to(dev), dtype)
    head = torch.nn.Sequential(torch.nn.Linear(4, 1)).to(dev)
    ctx_dim = 64 if tiny else 768
    layers = layers or (["up_8", "up_16", "up_32"] if tiny else ["mid_8", "up_16", "up_32", "up_64"])
    args = synthetic.default_args(pretrain_model_name="sd_1_5_attrcon", train_batch_size=B, K=K, total_step=S, gan_loss=True,
                                  gan_model_arch="gansd_1_5", attrcon_train_steps=n_attrcon, resolution=res, max_grad_norm=0.1, seed=3,
                                  lora_rank=rank)
    args.train_layer_ls = layers
    pipe = AttrConcenTrainableSDPipeline(EngineVAE(vae_p, dtype), EngineUNet(unet_p, dtype))
    register_attention_control(pipe, AttentionStore(layers))
    D = D_sd(EngineUNet(d_p, dtype), mlp=head)
    blip = Blip(BlipEngine(blip_model, dtype) if blip_wrap is None else blip_wrap(blip_model))
    tr = CoMatTrainer(args, pipe, CaptionModelWrapper(["Blip"], [1.0], blip), D)
    hb = synthetic.synthetic_batch(B, 5, ctx_dim, res, True, True)
    for m in hb["masks_host"]:                 # no "object not detected" (all-false) masks here: they make the token loss a constant
        for i in range(m.shape[0]):
            if not bool(m[i].any()):
                m[i, :, res // 5: res // 2 + 16 * i, res // 4: 3 * res // 4] = True
    batch, _ = synthetic.batch_to_device(hb, dev)
    g = torch.Generator().manual_seed(9)
    lat = res // 8
    batch["init_latents"] = torch.randn(B, 4, lat, lat, generator=g).to(dev)
    batch["noises"] = [torch.randn(B, 4, lat, lat, generator=g).to(dev) for _ in range(S)]
    batch["training_steps"], batch["attrcon_steps"] = R.select_training_steps(S, K, random.Random(1), n_attrcon)
    batch["crop"] = (1, 0)
    ctrl = R.AttentionStore(layers)
    R.register_attention_control(o_unet, ctrl)
    ob = dict(prompt_embeds=batch["prompt_embeds"], null_embeds=batch["null_embeds"], latents=batch["init_latents"], noises=batch["noises"],
              training_steps=batch["training_steps"], attrcon_steps=batch["attrcon_steps"], crop=(1, 0),
              blip_ids=batch["blip"]["input_ids"], blip_mask=batch["blip"]["attention_mask"], gan_null_embeds=batch["gan_null_embeds"],
              words=batch["words"], masks=batch["masks"])
    return dict(trainer=tr, batch=batch, D=D,
                oracle=dict(unet=o_unet, vae=o_vae, d=o_d, head=head, blip=blip_model, ctrl=ctrl, batch=ob,
                            cfg=dict(S=S, resolution=res, train_layer_ls=layers)))


def g_step_compare(world, with_grads=True):
    """product ``CoMatTrainer.g_losses`` (+ backward into the flat LoRA-gradient buffer) vs ``oracle.comat_ref.g_step_loss`` (+ autograd).
    returns {name: relative error} plus 'grad_cos' / 'grad_norm_ratio' and the raw scalars."""
    tr, batch, o = world["trainer"], world["batch"], world["oracle"]
    logs = tr.g_losses(batch)
    out = {"product": {k: float(logs[k].detach()) for k in ("Blip", "G_loss", "token_loss", "pixel_loss", "loss")}}
    image_p = logs["_image"].detach().float().clone()
    got = None
    if with_grads:
        tr.optimizer.zero_grad()
        logs["loss"].backward()
        tr.pipeline.unet.finalize_lora_grads()      # trainer protocol: accumulated dy^T x products -> d up / d down, once per step
        got = tr.optimizer.grad.double().clone()
    del logs
    o["d"].eval()
    ref = R.g_step_loss(o["unet"], o["vae"], sdm.DDPMScheduler(), o["blip"], o["batch"], o["cfg"], controller=o["ctrl"],
                        d_unet=o["d"], d_head=o["head"])
    out["oracle"] = {k: float(ref[k].detach()) for k in ("Blip", "G_loss", "token_loss", "pixel_loss", "loss")}
    for k in ("Blip", "G_loss", "token_loss", "pixel_loss", "loss"):
        out[k] = rel_scalar(out["product"][k], out["oracle"][k])
    out["image"] = rel_l2(image_p, ref["image"].detach())
    if with_grads:
        params = [p for p in o["unet"].parameters() if p.requires_grad]
        g_ref = torch.autograd.grad(ref["loss"], params, allow_unused=True)
        want = torch.cat([(gr if gr is not None else torch.zeros_like(p)).reshape(-1) for p, gr in zip(params, g_ref)]).double()
        out["grad_cos"] = cosine(got, want)
        out["grad_norm_ratio"] = float(got.norm() / want.norm().clamp_min(1e-30))
    return out


def d_step_compare(world, lat_seed=11):
    """discriminator side (gan_sdxl.py:92-132): loss and the gradients of the D LoRA factors + Linear(4,1) head."""
    tr, batch, o, D = world["trainer"], world["batch"], world["oracle"], world["D"]
    a = tr.args
    g = torch.Generator().manual_seed(lat_seed)
    fake = torch.randn(batch["real_latents"].shape, generator=g).to(batch["real_latents"].device)
    d_loss = D.D_sd_pipeline_forward(fake, side="D", negative_prompt_embeds=batch["gan_null_embeds"], num_inference_steps=a.total_step,
                                     batch={"latents": batch["real_latents"]})
    tr.D_optimizer.zero_grad()
    d_loss.backward()
    D.unet.finalize_lora_grads()
    got = tr.D_optimizer.grad.double().clone()
    o["d"].train()
    for p in o["head"].parameters():
        p.requires_grad_(True)
    ref = R.d_forward(o["d"], o["head"], sdm.DDPMScheduler(), fake, batch["gan_null_embeds"], a.total_step, "D", batch["real_latents"])
    params = [p for p in o["d"].parameters() if p.requires_grad] + list(o["head"].parameters())
    g_ref = torch.autograd.grad(ref, params, allow_unused=True)
    want = torch.cat([(gr if gr is not None else torch.zeros_like(p)).reshape(-1) for p, gr in zip(params, g_ref)]).double()
    n_head = sum(p.numel() for p in o["head"].parameters())
    return {"D_loss": rel_scalar(d_loss, ref), "product": float(d_loss.detach()), "oracle": float(ref.detach()),
            "grad_cos": cosine(got[:-n_head], want[:-n_head]), "grad_norm_ratio": float(got[:-n_head].norm() / want[:-n_head].norm().clamp_min(1e-30)),
            "head_grad_rel": rel_l2(got[-n_head:], want[-n_head:])}
