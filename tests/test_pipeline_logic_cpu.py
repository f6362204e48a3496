"""CPU: the drop-in pipelines (gradient windows, CFG, DDPM step, attrcon capture, autograd wiring through the explicit
executors) against the goldens produced by the reference's OWN AttrConcenTrainableSDPipeline.forward
(oracle/pin_against_reference.py), with the C-ABI ops emulated in torch (tests/cpu_ops_emulation.py)."""
import pytest
import torch

from oracle import fixtures as FX
from tests import cpu_ops_emulation as EMU


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return (a - b).abs().max().item() / max(1e-12, b.abs().max().item())


@pytest.mark.parametrize("case", FX.PIPELINE_CASES, ids=FX.case_key)
def test_attrcon_pipeline_matches_reference_golden(monkeypatch, case, golden):
    EMU.install(monkeypatch)
    from comat_b200 import pipelines as PL
    from comat_b200.modules import EngineUNet, EngineVAE
    g = golden("pipeline")[FX.case_key(case)]
    w = FX.pipeline_world(**case)
    unet = w["make_unet"]()
    pipe = PL.AttrConcenTrainableSDPipeline(EngineVAE(w["vae"], torch.float32), EngineUNet(unet, torch.float32))
    ctrl = PL.AttentionStore(w["train_layer_ls"])
    assert PL.register_attention_control(pipe, ctrl) == g["num_att_layers"]
    gen = torch.Generator().manual_seed(case["seed"] + 77)
    noises = [torch.randn(w["latents"].shape, generator=gen) for _ in range(case["S"])]
    image, lat = pipe.forward(prompt=["p"] * case["B"], height=case["hw"] * 8, width=case["hw"] * 8,
                              training_timesteps=w["training_steps"], detach_gradient=True, num_inference_steps=case["S"],
                              guidance_scale=7.5, guidance_rescale=case.get("rescale", 0.0),
                              negative_prompt_embeds=w["null_embeds"], prompt_embeds=w["prompt_embeds"],
                              latents=w["latents"].clone(), return_latents=True, bp_on_trained=True,
                              attrcon_train_steps=w["attrcon_steps"], noises=noises)
    assert sorted(pipe.attn_dict.keys()) == g["timesteps"]
    assert {k: len(v) for k, v in next(iter(pipe.attn_dict.values())).items()} == g["keyset"]
    assert rel(lat, g["latents"]) < 1e-4
    assert rel(image.double().mean(), g["image_mean"]) < 1e-4 and rel(image.double().norm(), g["image_l2"]) < 1e-4
    loss = (image.float() ** 2).mean() + sum((m.float() ** 2).sum() for d in pipe.attn_dict.values() for v in d.values() for m in v) * 1e-3
    assert rel(loss, g["loss"]) < 1e-4
    params = pipe.unet.lora_parameters()
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    ref_params = [p for p in unet.parameters() if p.requires_grad]
    assert [id(p) for p in params] == [id(p) for p in ref_params]          # same order as training_utils/pipeline.py:123-143
    for gr, l2 in zip(grads, g["grad_l2"]):
        got = 0.0 if gr is None else float(gr.double().norm())
        assert abs(got - l2) <= 2e-3 * max(l2, 1e-6) + 1e-9, (got, l2)
