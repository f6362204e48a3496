"""GPU: the GAN ground-truth producer - text encoding, CUDA-graphed 16-bit sampling, file output - vs the fp32 oracle
rollout on the same weights and the same CUDA generator stream; and the checkpoint round trip on a live device trainer (8f-3).

Tolerance: an fp16 emulation of this schedule (tests/cpu_ops_emulation.py) differs from the fp32 oracle by 2e-3 relative L2 after
10 sampler steps; 2e-2 leaves room for the tensor-core accumulation order."""
import os

import pytest
import torch

from oracle import comat_ref as R
from oracle import fixtures as FX
from oracle import sd_modules as sdm

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


class _Tok(FX.ClipTokenizerStub):
    def __call__(self, *a, **k):
        t = super().__call__(*a, **k)
        t.input_ids = t.input_ids.cuda()
        return t


def _world(dtype=torch.float16):
    from comat_b200 import containers as Cn, synthetic
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import TrainableSDPipeline
    from comat_b200.text_encoder import EngineCLIPText
    torch.manual_seed(0)
    with torch.device("cuda"):
        unet = Cn.UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=128)
        vae = Cn.AutoencoderKL(block_out_channels=(64, 64, 128, 128))
    unet.requires_grad_(False); vae.requires_grad_(False)
    unet.install_lora(8, up_std=0.05)
    clip = R.make_clip_text("clip_l", tiny=True, seed=21, device="cuda")
    o_unet = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(width=64, cross_attention_dim=128)).cuda()
    o_unet.requires_grad_(False)
    sdm.install_lora(o_unet, 8)
    o_unet.cuda().load_state_dict(unet.state_dict())
    pipe = TrainableSDPipeline(EngineVAE(vae, dtype), EngineUNet(unet, dtype), text_encoder=EngineCLIPText(clip, dtype),
                               tokenizer=synthetic.SyntheticClipTokenizer())
    return pipe, o_unet, clip


def test_gan_ground_truth_producer_vs_oracle(tmp_path):
    from comat_b200 import _lib, gan_data as GD
    pipe, o_unet, clip = _world()
    prompts = ["a red apple", "two dogs on a sofa", "a blue car"]
    index = str(tmp_path / "train_data" / "gan_train_data.jsonl")
    S, hw = 10, 256
    l0 = _lib.LAUNCH_COUNT
    n = GD.generate_gan_ground_truth(pipe, prompts, index, batch_size=2, num_inference_steps=S, guidance_scale=7.5, height=hw, width=hw,
                                     generator=torch.Generator(device="cuda").manual_seed(11))
    assert n == 3 and _lib.LAUNCH_COUNT > l0
    assert len(pipe.unet._graphs) == 2                              # one captured forward per batch shape (2 and 1 prompts)
    recs = GD.read_jsonl(index)
    assert [r["prompt"] for r in recs] == prompts
    got = [torch.load(r["file_path"]) for r in recs]
    assert all(t.shape == (4, hw // 8, hw // 8) and t.dtype == torch.float32 and t.device.type == "cpu" for t in got)
    gen = torch.Generator(device="cuda").manual_seed(11)
    k = 0
    for i in range(0, 3, 2):
        chunk = prompts[i:i + 2]
        pe, npe, _ = R.encode_prompt_sd(clip, _Tok(), chunk, 1, True)
        z = torch.randn(len(chunk), 4, hw // 8, hw // 8, generator=gen, device="cuda")
        noises = [torch.randn(z.shape, generator=gen, device="cuda") for _ in range(S)]
        _, lat, _ = R.rollout(o_unet, None, sdm.DDPMScheduler(), pe, npe, z, noises, S, [], 7.5, decode=False)
        for j in range(len(chunk)):
            print(f"[measured] producer latent {k}: {rel(got[k].cuda(), lat[j]):.2e}")
            assert rel(got[k].cuda(), lat[j]) < 2e-2, (k, rel(got[k].cuda(), lat[j]))
            k += 1
    # the reader hands the same tensors back, batched for the discriminator step
    from types import SimpleNamespace
    ds = GD.Gan_Dataset(SimpleNamespace(training_prompts=index))
    b = GD.collate_gan_batch([ds[0], ds[2]])
    assert torch.equal(b["real_latents"][1], got[2]) and b["text"] == [prompts[0], prompts[2]]
    # validation-style sampling: decoded, clamped image
    img = pipe(prompts[:1], height=hw, width=hw, num_inference_steps=2, generator=torch.Generator(device="cuda").manual_seed(1),
               output_type="pt").images
    assert img.shape == (1, 3, hw, hw) and img.is_cuda and float(img.min()) >= 0 and float(img.max()) <= 1 and torch.isfinite(img).all()


def test_checkpoint_round_trip_on_device(tmp_path):
    """save after real optimiser steps' worth of state, resume into a fresh trainer: identical parameters, moments, 16-bit operand
    images, and an identical next UNet output."""
    import random
    from comat_b200 import checkpoint as CK, synthetic
    from comat_b200.gan import D_sd
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import TrainableSDPipeline
    from comat_b200.trainer import CoMatTrainer

    def make(seed):
        unet, vae = synthetic.build_sd15("cuda", torch.float16, rank=8, seed=seed, tiny=True, lora_up_std=0.05)
        d, _ = synthetic.build_sd15("cuda", torch.float16, rank=8, seed=seed + 100, tiny=True, lora_up_std=0.02)
        args = synthetic.default_args(pretrain_model_name="sd_1_5", gan_loss=True, seed=seed)
        pipe = TrainableSDPipeline(EngineVAE(vae, torch.float16), EngineUNet(unet, torch.float16))
        return CoMatTrainer(args, pipe, None, D_sd(EngineUNet(d, torch.float16)), rng=random.Random(seed))
    a = make(1)
    g = torch.Generator(device="cuda").manual_seed(3)
    for opt in (a.optimizer, a.D_optimizer):                      # two real fused clip+AdamW steps on random gradients
        for _ in range(2):
            opt.grad.copy_(torch.randn(opt.n, generator=g, device="cuda") * 1e-3)
            opt.step()
    a.pipeline.unet.refresh_lora(); a.D.unet.refresh_lora()
    a.global_step = 2
    path = CK.save_checkpoint(a, str(tmp_path))
    assert sorted(os.listdir(path)) == ["D_sd", "pytorch_lora_weights.safetensors", "trainer_state.pt"]
    b = make(2)
    assert CK.load_checkpoint(b, str(tmp_path), "latest") == 2
    for x, y in ((a.optimizer, b.optimizer), (a.D_optimizer, b.D_optimizer)):
        assert torch.equal(x.flat, y.flat) and torch.equal(x.m, y.m) and torch.equal(x.v, y.v) and y.step_count == 2
    x = torch.randn(2, 4, 32, 32, device="cuda")
    ehs = torch.randn(2, 77, 64, device="cuda")
    t = torch.tensor(500, device="cuda")
    with torch.no_grad():
        # base weights differ between the two trainers (different seeds): compare through the LoRA operand images instead
        for la, lb in zip(a.pipeline.unet.engine.loras, b.pipeline.unet.engine.loras):
            assert torch.equal(la.down16, lb.down16) and torch.equal(la.up16, lb.up16)
        ya = a.pipeline.unet(x, t, encoder_hidden_states=ehs)[0]
        assert torch.isfinite(ya).all()


def test_train_step_from_prompt_strings_on_device():
    """training_script.py:513-525,575-588 on the device: '' encoded once by the trainer, batch['text'] encoded inside forward;
    the loss equals the one obtained from the oracle's fp32 embeddings of the same strings (16-bit encoder error only)."""
    from comat_b200 import synthetic
    from comat_b200.blip_engine import BlipEngine
    from comat_b200.caption import Blip, CaptionModelWrapper
    from comat_b200.trainer import CoMatTrainer
    pipe, _, clip = _world()
    B, S, res = 2, 2, 256
    args = synthetic.default_args(pretrain_model_name="sd_1_5", train_batch_size=B, K=1, total_step=S, gan_loss=False, resolution=res, seed=3)
    blip = Blip(BlipEngine(R.make_blip(large=False).cuda(), torch.float16))
    tr = CoMatTrainer(args, pipe, CaptionModelWrapper(["Blip"], [1.0], blip), None)
    prompts = ["a red apple on a table", "two dogs"]
    g = torch.Generator().manual_seed(9)
    ids, mask = FX.blip_token_batch(g, B, 8)
    base = dict(blip={"input_ids": ids.cuda(), "attention_mask": mask.cuda()}, init_latents=torch.randn(B, 4, res // 8, res // 8, generator=g).cuda(),
                noises=[torch.randn(B, 4, res // 8, res // 8, generator=g).cuda() for _ in range(S)], training_steps=[1], attrcon_steps=None, crop=(0, 0))
    l_text = tr.g_losses(dict(base, text=prompts))["loss"].detach()
    pe, _, _ = R.encode_prompt_sd(clip, _Tok(), prompts, 1, False)
    null, _, _ = R.encode_prompt_sd(clip, _Tok(), "", B, False)
    assert tr.null_embed.shape == (B, 77, 128) and rel(tr.null_embed, null) < 5e-3
    l_emb = tr.g_losses(dict(base, prompt_embeds=pe, null_embeds=null))["loss"].detach()
    print(f"[measured] step loss from strings {float(l_text):.6f} vs from oracle embeddings {float(l_emb):.6f}")
    assert abs(float(l_text) - float(l_emb)) < 5e-3 * abs(float(l_emb))
    out = tr.train_step(dict(base, text=prompts))                    # full step (backward + fused AdamW) from strings
    tr.sync()
    assert torch.isfinite(out["step_loss"])
