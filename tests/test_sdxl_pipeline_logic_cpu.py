"""CPU: SDXL twin of the drop-in pipeline (TrainableSDPipeline.py:657-846 / AttrConcenTrainableSDXLPipeline.py:234-496):
always-detached UNet input, pooled-text + time-id conditioning, un-rescaled image when return_latents — vs the oracle rollout
(pinned to the reference's own SD1.5 and SDXL pipelines: tests/golden/pipeline.pt, sdxl_pipeline.pt), with the C-ABI ops emulated."""
import random

import torch

from oracle import comat_ref as R
from oracle import fixtures as FX
from oracle import sd_modules as sdm
from tests import cpu_ops_emulation as EMU


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def test_sdxl_attrcon_pipeline_matches_oracle(monkeypatch):
    EMU.install(monkeypatch)
    from comat_b200 import pipelines as PL
    from comat_b200.modules import EngineUNet, EngineVAE
    B, S, hw = 2, 3, 32
    unet = FX.make_tiny_unet(31, rank=4, sdxl=True)
    vae = FX.make_tiny_vae(32, sdxl=True)
    g = torch.Generator().manual_seed(33)
    pe, ne = torch.randn(B, 77, 64, generator=g), torch.randn(B, 77, 64, generator=g)
    pooled, npooled = torch.randn(B, 16, generator=g), torch.randn(B, 16, generator=g)
    lat0 = torch.randn(B, 4, hw, hw, generator=g)
    noises = [torch.randn(B, 4, hw, hw, generator=g) for _ in range(S)]
    T, A = R.select_training_steps(S, 1, random.Random(2), 2)
    layers = ["up_8", "up_16"]
    # oracle
    ctrl = R.AttentionStore(layers)
    R.register_attention_control(unet, ctrl)
    ids = torch.tensor([[256., 256, 0, 0, 256, 256]]).repeat(B, 1)
    added = {"text_embeds": torch.cat([npooled, pooled]), "time_ids": torch.cat([ids, ids])}
    img_o, lat_o, attn_o = R.rollout(unet, vae, sdm.DDPMScheduler(), pe, ne, lat0.clone(), noises, S, T, 7.5, 0.0, A, ctrl,
                                     added_cond_kwargs=added, sdxl=True, return_latents=True)
    # product
    unet2 = FX.make_tiny_unet(31, rank=4, sdxl=True)
    pipe = PL.AttrConcenTrainableSDXLPipeline(EngineVAE(vae, torch.float32), EngineUNet(unet2, torch.float32))
    PL.register_attention_control(pipe, PL.AttentionStore(layers))
    img, lat = pipe.forward(prompt=["p"] * B, height=hw * 8, width=hw * 8, training_timesteps=T, num_inference_steps=S,
                            guidance_scale=7.5, prompt_embeds=pe, negative_prompt_embeds=ne, pooled_prompt_embeds=pooled,
                            negative_pooled_prompt_embeds=npooled, latents=lat0.clone(), return_latents=True,
                            attrcon_train_steps=A, noises=noises)
    assert rel(lat, lat_o) < 1e-4 and rel(img, img_o) < 1e-4                 # image NOT rescaled to [0,1] (:838-840 quirk)
    assert sorted(pipe.attn_dict) == sorted(attn_o)
    for t in attn_o:
        for k in attn_o[t]:
            for a, b in zip(pipe.attn_dict[t][k], attn_o[t][k]):
                assert rel(a, b) < 1e-4
    loss = (img ** 2).mean()
    loss_o = (img_o ** 2).mean()
    g1 = torch.autograd.grad(loss, pipe.unet.lora_parameters(), allow_unused=True)
    g2 = torch.autograd.grad(loss_o, [p for p in unet.parameters() if p.requires_grad], allow_unused=True)
    for a, b in zip(g1, g2):
        if b is not None and float(b.abs().max()) > 0:
            assert rel(a, b) < 5e-3
