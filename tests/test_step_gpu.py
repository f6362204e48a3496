"""GPU parity of the whole training-step loss assembly: product trainer on the CUDA kernels vs the
fp32 oracle on the same weights / noise / steps / crop.  Tolerance: north star = 1e-3 relative on the concept-matching loss
and the per-token attention loss."""
import random

import pytest
import torch

from oracle import comat_ref as R
from oracle import sd_modules as sdm

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64).cpu(), torch.as_tensor(b, dtype=torch.float64).cpu()
    return (a - b).abs().max().item() / max(1e-12, b.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.float16])
def test_train_step_losses_and_grads_vs_oracle(dtype):
    from comat_b200 import synthetic
    from comat_b200.blip_engine import BlipEngine
    from comat_b200.caption import Blip, CaptionModelWrapper
    from comat_b200.gan import D_sd
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import AttentionStore, AttrConcenTrainableSDPipeline, register_attention_control
    from comat_b200.trainer import CoMatTrainer
    dev = torch.device("cuda")
    B, S, K, res = 2, 4, 2, 256
    unet_p, vae_p = synthetic.build_sd15(dev, dtype, rank=8, seed=7, tiny=True, lora_up_std=0.05)
    d_p, _ = synthetic.build_sd15(dev, dtype, rank=8, seed=8, tiny=True)
    o_unet = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(width=64, cross_attention_dim=64)).to(dev)
    o_vae = sdm.AutoencoderKL(block_out_channels=(64, 64, 128, 128)).to(dev)
    o_d = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(width=64, cross_attention_dim=64)).to(dev)
    for o, p in ((o_unet, unet_p), (o_d, d_p)):
        o.requires_grad_(False)
        sdm.install_lora(o, 8)
        o.to(dev)
        o.load_state_dict(p.state_dict())
    o_vae.load_state_dict(vae_p.state_dict())
    o_vae.requires_grad_(False)
    blip_model = R.make_blip(large=False).to(dev)
    head = torch.nn.Sequential(torch.nn.Linear(4, 1)).to(dev)
    args = synthetic.default_args(pretrain_model_name="sd_1_5_attrcon", train_batch_size=B, K=K, total_step=S, gan_loss=True,
                                  gan_model_arch="gansd_1_5", attrcon_train_steps=2, resolution=res, max_grad_norm=0.1, seed=3)
    args.train_layer_ls = ["up_8", "up_16", "up_32"]
    pipe = AttrConcenTrainableSDPipeline(EngineVAE(vae_p, dtype), EngineUNet(unet_p, dtype))
    register_attention_control(pipe, AttentionStore(args.train_layer_ls))
    D = D_sd(EngineUNet(d_p, dtype), mlp=head)
    tr = CoMatTrainer(args, pipe, CaptionModelWrapper(["Blip"], [1.0], Blip(BlipEngine(blip_model, dtype))), D)
    batch, _ = synthetic.batch_to_device(synthetic.synthetic_batch(B, 5, 64, res, True, True), dev)
    g = torch.Generator().manual_seed(9)
    lat = res // 8
    batch["init_latents"] = torch.randn(B, 4, lat, lat, generator=g).to(dev)
    batch["noises"] = [torch.randn(B, 4, lat, lat, generator=g).to(dev) for _ in range(S)]
    batch["training_steps"], batch["attrcon_steps"] = R.select_training_steps(S, K, random.Random(1), 2)
    batch["crop"] = (1, 0)
    logs = tr.g_losses(batch)
    ctrl = R.AttentionStore(args.train_layer_ls)
    R.register_attention_control(o_unet, ctrl)
    ob = dict(prompt_embeds=batch["prompt_embeds"], null_embeds=batch["null_embeds"], latents=batch["init_latents"], noises=batch["noises"],
              training_steps=batch["training_steps"], attrcon_steps=batch["attrcon_steps"], crop=(1, 0),
              blip_ids=batch["blip"]["input_ids"], blip_mask=batch["blip"]["attention_mask"], gan_null_embeds=batch["gan_null_embeds"],
              words=batch["words"], masks=batch["masks"])
    o_d.eval()
    ref = R.g_step_loss(o_unet, o_vae, sdm.DDPMScheduler(), blip_model, ob, dict(S=S, resolution=res, train_layer_ls=args.train_layer_ls),
                        controller=ctrl, d_unet=o_d, d_head=head)
    errs = {k: rel(logs[k], ref[k]) for k in ("Blip", "G_loss", "token_loss", "pixel_loss", "loss")}
    print("step parity rel errors:", errs)
    assert errs["Blip"] < 1e-3 and errs["token_loss"] < 1e-3, errs          # north-star tolerance
    assert errs["pixel_loss"] < 2e-3 and errs["G_loss"] < 5e-3 and errs["loss"] < 1e-3, errs
    assert rel(logs["_image"], ref["image"]) < 3e-2
    g_ref = torch.autograd.grad(ref["loss"], [p for p in o_unet.parameters() if p.requires_grad], allow_unused=True)
    tr.optimizer.zero_grad()
    logs["loss"].backward()
    tr.pipeline.unet.finalize_lora_grads()          # trainer protocol: accumulated dy^T x products -> d up / d down, once per step
    got = tr.optimizer.grad.double()
    want = torch.cat([(gr if gr is not None else torch.zeros_like(p)).reshape(-1) for p, gr in zip(tr.G_parameters, g_ref)]).double()
    cos = float((got * want).sum() / (got.norm() * want.norm()))
    print("LoRA grad cosine:", cos, "norm ratio:", float(got.norm() / want.norm()))
    assert cos > 0.98 and abs(float(got.norm() / want.norm()) - 1) < 0.1
    out = tr.train_step(batch)
    assert torch.isfinite(out["step_loss"]) and torch.isfinite(out["D_loss"])


def _make_trainer(dtype, overlap):
    from comat_b200 import synthetic
    from comat_b200.blip_engine import BlipEngine
    from comat_b200.caption import Blip, CaptionModelWrapper
    from comat_b200.gan import D_sd
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import AttentionStore, AttrConcenTrainableSDPipeline, register_attention_control
    from comat_b200.trainer import CoMatTrainer
    dev = torch.device("cuda")
    B, S, K, res = 2, 4, 2, 256
    unet_p, vae_p = synthetic.build_sd15(dev, dtype, rank=8, seed=7, tiny=True, lora_up_std=0.05)
    d_p, _ = synthetic.build_sd15(dev, dtype, rank=8, seed=8, tiny=True)
    torch.manual_seed(21)
    blip_model = R.make_blip(large=False).to(dev)
    head = torch.nn.Sequential(torch.nn.Linear(4, 1)).to(dev)
    args = synthetic.default_args(pretrain_model_name="sd_1_5_attrcon", train_batch_size=B, K=K, total_step=S, gan_loss=True,
                                  gan_model_arch="gansd_1_5", attrcon_train_steps=2, resolution=res, max_grad_norm=0.1, seed=3,
                                  learning_rate=1e-3, learning_rate_D=1e-3)
    args.train_layer_ls = ["up_8", "up_16", "up_32"]
    pipe = AttrConcenTrainableSDPipeline(EngineVAE(vae_p, dtype), EngineUNet(unet_p, dtype))
    register_attention_control(pipe, AttentionStore(args.train_layer_ls))
    D = D_sd(EngineUNet(d_p, dtype), mlp=head)
    tr = CoMatTrainer(args, pipe, CaptionModelWrapper(["Blip"], [1.0], Blip(BlipEngine(blip_model, dtype))), D, rng=random.Random(5))
    tr.overlap_updates = overlap
    batches = [synthetic.batch_to_device(synthetic.synthetic_batch(B, 40 + i, 64, res, True, True), dev)[0] for i in range(3)]
    return tr, batches


def test_side_stream_updates_match_serial_updates():
    """the optimiser tails (gradient projection, clip + AdamW, operand refresh) issued on the side stream next to the
    discriminator step / the next rollout give the same parameters and losses as the serial schedule over several steps."""
    results = []
    for overlap in (False, True):
        tr, batches = _make_trainer(torch.float16, overlap)
        losses = []
        for i in range(4):
            out = tr.train_step(batches[i % len(batches)])
            losses.append((float(out["step_loss"]), float(out["D_loss"])))
        tr.sync()
        torch.cuda.synchronize()
        results.append((losses, tr.optimizer.flat.clone(), tr.D_optimizer.flat.clone()))
    (l0, g0, d0), (l1, g1, d1) = results
    # same arithmetic on both schedules; fp32 atomics / GroupNorm shared-memory atomics reorder sums, hence tolerances
    for (a, b), (c, d) in zip(l0, l1):
        assert abs(a - c) <= 2e-3 * max(1.0, abs(a)) and abs(b - d) <= 2e-3 * max(1.0, abs(b)), (l0, l1)
    assert float((g0 - g1).abs().max()) <= 5e-2 * float(g0.abs().max())
    assert float((d0 - d1).abs().max()) <= 5e-2 * float(d0.abs().max())
    moved = float((g0 - _make_trainer(torch.float16, False)[0].optimizer.flat).abs().max())
    assert moved > 1e-3                                   # the four steps did change the parameters
