"""CPU: step assembly of the product trainer (training_script.py:556-664 mirror) vs the oracle's g_step_loss on the same
weights, noise, steps and crop — with the CUDA-only pieces replaced by semantic emulations (test infrastructure)."""
import random

import pytest
import torch

from oracle import comat_ref as R
from oracle import sd_modules as sdm
from tests import cpu_ops_emulation as EMU
from tests.hf_blip import HFBlipComparator


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return (a - b).abs().max().item() / max(1e-12, b.abs().max().item())


def _emulate_cuda_only(monkeypatch):
    EMU.install(monkeypatch)
    from comat_b200 import attn_loss, image_ops, optim
    import torch.nn.functional as F

    def resize_norm(images, size, mean, std):
        x = F.interpolate(images.float(), size=(size, size), mode="bicubic", antialias=True, align_corners=False)
        m = torch.tensor(mean).view(1, -1, 1, 1)
        sd = torch.tensor(std).view(1, -1, 1, 1)
        return (x - m) / sd
    monkeypatch.setattr(image_ops, "resize_bicubic_aa_normalize", resize_norm)

    def get_mask_loss(attn_map, words, masks, layers, tokens=77):
        some = next(iter(next(iter(attn_map.values())).values()))[0]
        return R.mask_loss(attn_map, words, masks, layers, some)
    monkeypatch.setattr(attn_loss, "get_mask_loss", get_mask_loss)

    def step(self, handle=None):
        self.step_count += 1
        g = self.grad
        if self.max_norm > 0:
            g = g * min(1.0, self.max_norm / (float(g.norm()) + 1e-6))
        b1, b2 = self.betas
        self.flat.mul_(1 - self.lr * self.wd)
        self.m.mul_(b1).add_(g, alpha=1 - b1)
        self.v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = self.v.sqrt() / (1 - b2 ** self.step_count) ** 0.5 + self.eps
        self.flat.addcdiv_(self.m, denom, value=-self.lr / (1 - b1 ** self.step_count))
    monkeypatch.setattr(optim.FlatAdamW, "step", step)


def test_train_step_matches_oracle_step(monkeypatch):
    _emulate_cuda_only(monkeypatch)
    from comat_b200 import containers as Cn, synthetic
    from comat_b200.caption import Blip, CaptionModelWrapper
    from comat_b200.gan import D_sd
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import AttentionStore, AttrConcenTrainableSDPipeline, register_attention_control
    from comat_b200.trainer import CoMatTrainer
    B, S, K, res = 2, 3, 2, 128
    torch.manual_seed(0)
    # product-side containers, oracle modules loaded with the SAME state dict
    unet_p, vae_p = synthetic.build_sd15("cpu", torch.float32, rank=4, seed=7, tiny=True, lora_up_std=0.05)
    d_p, _ = synthetic.build_sd15("cpu", torch.float32, rank=4, seed=8, tiny=True)
    o_unet = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(width=64, cross_attention_dim=64))
    o_vae = sdm.AutoencoderKL(block_out_channels=(64, 64, 128, 128))
    o_d = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(width=64, cross_attention_dim=64))
    for o, p in ((o_unet, unet_p), (o_d, d_p)):
        o.requires_grad_(False)
        sdm.install_lora(o, 4)
        o.load_state_dict(p.state_dict())
    o_vae.load_state_dict(vae_p.state_dict())
    o_vae.requires_grad_(False)
    blip_model = R.make_blip(large=False)
    head = torch.nn.Sequential(torch.nn.Linear(4, 1))
    args = synthetic.default_args(pretrain_model_name="sd_1_5_attrcon", train_batch_size=B, K=K, total_step=S, gan_loss=True,
                                  gan_model_arch="gansd_1_5", attrcon_train_steps=2, resolution=res, max_grad_norm=0.1, seed=3)
    args.train_layer_ls = ["up_8", "up_16"]
    pipe = AttrConcenTrainableSDPipeline(EngineVAE(vae_p, torch.float32), EngineUNet(unet_p, torch.float32))
    register_attention_control(pipe, AttentionStore(args.train_layer_ls))
    D = D_sd(EngineUNet(d_p, torch.float32), mlp=head)
    tr = CoMatTrainer(args, pipe, CaptionModelWrapper(["Blip"], [1.0], Blip(HFBlipComparator(blip_model))), D)
    hb = synthetic.synthetic_batch(B, 5, 64, res, True, True)
    batch, _ = synthetic.batch_to_device(hb, "cpu")
    g = torch.Generator().manual_seed(9)
    lat = res // 8
    batch["init_latents"] = torch.randn(B, 4, lat, lat, generator=g)
    batch["noises"] = [torch.randn(B, 4, lat, lat, generator=g) for _ in range(S)]
    batch["training_steps"], batch["attrcon_steps"] = R.select_training_steps(S, K, random.Random(1), 2)
    batch["crop"] = (1, 0)
    logs = tr.g_losses(batch)
    # oracle on identical inputs
    ctrl = R.AttentionStore(args.train_layer_ls)
    R.register_attention_control(o_unet, ctrl)
    ob = dict(prompt_embeds=batch["prompt_embeds"], null_embeds=batch["null_embeds"], latents=batch["init_latents"], noises=batch["noises"],
              training_steps=batch["training_steps"], attrcon_steps=batch["attrcon_steps"], crop=(1, 0),
              blip_ids=batch["blip"]["input_ids"], blip_mask=batch["blip"]["attention_mask"], gan_null_embeds=batch["gan_null_embeds"],
              words=batch["words"], masks=batch["masks"])
    cfgd = dict(S=S, resolution=res, train_layer_ls=args.train_layer_ls)
    o_d.eval()
    ref = R.g_step_loss(o_unet, o_vae, sdm.DDPMScheduler(), blip_model, ob, cfgd, controller=ctrl, d_unet=o_d, d_head=head)
    for k_p, k_o in (("Blip", "Blip"), ("G_loss", "G_loss"), ("token_loss", "token_loss"), ("pixel_loss", "pixel_loss"), ("loss", "loss")):
        assert rel(logs[k_p], ref[k_o]) < 2e-4, (k_p, float(logs[k_p]), float(ref[k_o]))
    # gradients of the total loss w.r.t. the LoRA parameters
    g_ref = torch.autograd.grad(ref["loss"], [p for p in o_unet.parameters() if p.requires_grad], allow_unused=True)
    tr.optimizer.zero_grad()
    logs["loss"].backward()
    tr.pipeline.unet.finalize_lora_grads()      # trainer protocol: accumulated dy^T x products -> d up / d down, once per step
    off = 0
    for p, gr in zip(tr.G_parameters, g_ref):
        got = tr.optimizer.grad[off:off + p.numel()].view_as(p)
        off += p.numel()
        if gr is not None and float(gr.abs().max()) > 0:
            assert rel(got, gr) < 5e-3
    # a full step (G + D update) runs and changes the parameters
    before = tr.optimizer.flat.clone()
    out = tr.train_step(batch)
    assert "D_loss" in out and torch.isfinite(out["step_loss"])
    assert float((tr.optimizer.flat - before).abs().max()) > 0


def test_dynamic_loss_scale_follows_the_overflow_counter():
    """GradScaler's schedule driven by the optimiser's device-side skip counter (one step late, no host sync): halve after an
    overflow-skipped step, double after ``scale_growth_interval`` clean steps, bf16 executors (scale 1) are never touched."""
    from types import SimpleNamespace
    from comat_b200.trainer import CoMatTrainer
    tr = object.__new__(CoMatTrainer)
    fp16_a, fp16_b, bf16 = SimpleNamespace(grad_scale=4096.0), SimpleNamespace(grad_scale=16384.0), SimpleNamespace(grad_scale=1.0)
    tr._scaled = [m for m in (fp16_a, fp16_b, bf16) if m.grad_scale != 1.0]
    tr.scale_growth_interval, tr._clean_steps, tr.scale_min, tr.scale_max = 3, 0, 1.0, 65536.0
    feed = []
    tr.optimizer = SimpleNamespace(poll_overflow=lambda: feed.pop(0) if feed else None)
    tr.D_optimizer = None
    tr._update_loss_scale()                                     # nothing landed yet
    assert fp16_a.grad_scale == 4096.0
    feed[:] = [(1, 5)]
    tr._update_loss_scale()                                     # one skipped step seen -> halve
    assert (fp16_a.grad_scale, fp16_b.grad_scale, bf16.grad_scale) == (2048.0, 8192.0, 1.0)
    feed[:] = [(0, 6), (0, 7), (0, 8)]
    for _ in range(3):
        tr._update_loss_scale()
    assert (fp16_a.grad_scale, fp16_b.grad_scale) == (4096.0, 16384.0) and tr._clean_steps == 0
    fp16_a.grad_scale = 1.0
    feed[:] = [(2, 8)]
    tr._update_loss_scale()
    assert fp16_a.grad_scale == 1.0                             # floor


def test_close_gives_the_garbage_collector_back():
    """the step freezes the long-lived objects and disables the automatic collector; close() (called by train.py at the end
    of training) restores both."""
    import gc
    from comat_b200.trainer import CoMatTrainer
    t = object.__new__(CoMatTrainer)
    t.manual_gc_interval, t._ev_G, t._ev_D = 25, None, None
    assert gc.isenabled()
    try:
        t._gc_before_step()
        assert not gc.isenabled() and gc.get_freeze_count() > 0
        t.close()
        assert gc.isenabled() and gc.get_freeze_count() == 0
    finally:
        gc.unfreeze()
        gc.enable()
