"""GPU parity of the text-LoRA / context-gradient path (training_script.py:227-255, :569-573): d(encoder_hidden_states)
out of the UNet executor, the taped CLIP text encoder with LoRA, and a trainer step with ``--train_text_encoder_lora``.
First run on a B200 in round 2 (profiles/r02_gpu_tests_run1.log).  Tolerances follow the measured 16-bit errors of the sibling tests
(UNet input gradient 3e-2, text encoder 5e-3)."""
import pytest
import torch

from oracle import comat_ref as R
from oracle import fixtures as FX
from oracle import sd_modules as sdm

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("mode", ["product", "explicit", "frozen"])
def test_unet_context_gradient_on_device(mode):
    from comat_b200 import engine as E, ops
    torch.manual_seed(3)
    unet = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(width=64, cross_attention_dim=64))
    unet.requires_grad_(False)
    sdm.install_lora(unet, 8, up_std=0.05, seed=4)
    unet = unet.cuda()
    g = torch.Generator().manual_seed(0)
    x, ctx = torch.randn(2, 4, 32, 32, generator=g).cuda(), torch.randn(2, 77, 64, generator=g).cuda()
    dy, t = torch.randn(2, 4, 32, 32, generator=g).cuda(), torch.tensor(501, device="cuda")
    cr = ctx.clone().requires_grad_(True)
    g_ref = torch.autograd.grad(unet(x, t, cr, return_dict=False)[0], cr, dy)[0]
    eng = E.UNetEngine(unet, torch.float16)
    if mode != "frozen":
        eng.lora_train_impl = mode
    tape = E.Tape()
    cv = E.Var(ctx.half(), needs_grad=True)
    out = eng.forward(tape, E.Var(ops.latent_to_nhwc(x, torch.float16, 64), False), t, cv, lora_mode="frozen" if mode == "frozen" else "train")
    out.g = dy.permute(0, 2, 3, 1).contiguous().half()
    tape.backward()
    print(f"[measured] d ctx ({mode}): {rel(cv.g, g_ref):.2e}")
    assert rel(cv.g, g_ref) < 3e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 5e-3), (torch.bfloat16, 3e-2)])
def test_text_lora_taped_executor_on_device(dtype, tol):
    from comat_b200.text_encoder import EngineCLIPText, install_text_lora
    model = R.make_clip_text("clip_l", tiny=False, seed=3, device="cuda", layers=4)
    params = install_text_lora(model, 16, up_std=0.05)
    hooks = R.add_text_lora_hooks(model)
    ids = FX.ClipTokenizerStub()(["a photo of a cat", "two red cubes on a blue sphere", ""]).input_ids.cuda()
    # a loss-sized upstream gradient: N(0, 1) x 4096 (the fp16 loss scale) overflows fp16 after a few layers, which is the guard's
    # business (test_parity_fullsize_gpu.py::test_flat_adamw_skips_a_non_finite_gradient), not this comparison's
    dy = (torch.randn(3, 77, 768, generator=torch.Generator().manual_seed(3)) * 1e-3).cuda()
    ref = model(ids).last_hidden_state
    g_ref = torch.autograd.grad(ref, params, dy)
    for h in hooks:
        h.remove()
    enc = EngineCLIPText(model, dtype)
    last = enc(ids)[0]
    got = torch.autograd.grad(last, params, dy)
    errs = [rel(a, b) for a, b in zip(got, g_ref)]
    print(f"[measured] text LoRA {dtype}: out {rel(last, ref):.2e}, grads max {max(errs):.2e}")
    assert rel(last, ref) < tol and max(errs) < 10 * tol


def test_trainer_step_with_text_lora_on_device():
    from comat_b200 import containers as Cn, synthetic
    from comat_b200.blip_engine import BlipEngine
    from comat_b200.caption import Blip, CaptionModelWrapper
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import TrainableSDPipeline
    from comat_b200.text_encoder import EngineCLIPText, install_text_lora
    from comat_b200.trainer import CoMatTrainer
    B, S, res = 2, 2, 256
    torch.manual_seed(0)
    with torch.device("cuda"):
        unet = Cn.UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=128)
        vae = Cn.AutoencoderKL(block_out_channels=(64, 64, 128, 128))
    unet.requires_grad_(False); vae.requires_grad_(False)
    unet.install_lora(8, up_std=0.05)
    clip = R.make_clip_text("clip_l", tiny=True, seed=21, device="cuda")
    tparams = install_text_lora(clip, 8, up_std=0.05)
    args = synthetic.default_args(pretrain_model_name="sd_1_5", train_batch_size=B, K=1, total_step=S, gan_loss=False, resolution=res, seed=3,
                                  train_text_encoder_lora=True)
    enc = EngineCLIPText(clip, torch.float16)
    pipe = TrainableSDPipeline(EngineVAE(vae, torch.float16), EngineUNet(unet, torch.float16), text_encoder=enc,
                               tokenizer=synthetic.SyntheticClipTokenizer())
    tr = CoMatTrainer(args, pipe, CaptionModelWrapper(["Blip"], [1.0], Blip(BlipEngine(R.make_blip(large=False).cuda(), torch.float16))), None)
    g = torch.Generator().manual_seed(9)
    ids, mask = FX.blip_token_batch(g, B, 8)
    base = dict(blip={"input_ids": ids.cuda(), "attention_mask": mask.cuda()}, init_latents=torch.randn(B, 4, res // 8, res // 8, generator=g).cuda(),
                noises=[torch.randn(B, 4, res // 8, res // 8, generator=g).cuda() for _ in range(S)], training_steps=[1], attrcon_steps=None, crop=(0, 0))
    prompts = ["a red apple on a table", "two dogs"]
    tr.optimizer.zero_grad()
    loss = tr.g_losses(dict(base, text=prompts))["loss"]
    loss.backward()
    got = [p.grad.clone() for p in tparams]
    tok = synthetic.SyntheticClipTokenizer()
    pe, null = enc(tok(prompts).input_ids)[0], enc(tok([""] * B).input_ids)[0]
    pe_leaf, null_leaf = pe.detach().requires_grad_(True), null.detach().requires_grad_(True)
    tr.optimizer.zero_grad()
    loss2 = tr.g_losses(dict(base, prompt_embeds=pe_leaf, null_embeds=null_leaf))["loss"]
    d_pe, d_null = torch.autograd.grad(loss2, [pe_leaf, null_leaf])
    want = torch.autograd.grad([pe, null], tparams, [d_pe, d_null])
    cos = [float((a.double() * b.double()).sum() / (a.double().norm() * b.double().norm()).clamp_min(1e-30)) for a, b in zip(got, want)]
    print(f"[measured] text-LoRA gradient cosine, one backward vs chain rule: min {min(cos):.4f}")
    assert min(cos) > 0.98
    before = [p.detach().clone() for p in tparams]
    tr.train_step(dict(base, text=prompts))
    tr.sync()
    assert any(not torch.equal(a, b) for a, b in zip(before, tparams))
