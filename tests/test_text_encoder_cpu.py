"""CPU: prompt encoding - oracle vs the golden outputs of the reference's own ``encode_prompt``
(tests/golden/encode_prompt.pt, written by oracle/pin_against_reference.py), and the product's text-encoder executor +
``encode_prompt`` host logic vs the oracle with the CUDA ops emulated in torch (test infrastructure)."""
import pytest
import torch

from oracle import comat_ref as R
from oracle import fixtures as FX
from tests import cpu_ops_emulation as EMU
from tests.hf_blip import HFBlipComparator


def _neg(case):
    return None if case["negative"] is None else [case["negative"]] * len(case["prompts"])


def test_oracle_encode_prompt_matches_reference_golden(golden):
    gold = golden("encode_prompt")
    assert len(gold) == len(FX.ENCODE_PROMPT_CASES)
    for case in FX.ENCODE_PROMPT_CASES:
        g = gold["seed%d" % case["seed"]]
        enc = R.make_clip_text("clip_l", tiny=True, seed=case["seed"])
        pe, npe, ids = R.encode_prompt_sd(enc, FX.ClipTokenizerStub(), case["prompts"], case["n_per"], case["cfg"],
                                          negative_prompt=_neg(case), clip_skip=case["clip_skip"])
        assert torch.equal(ids, g["input_ids"])
        assert pe.shape == (len(case["prompts"]) * case["n_per"], 77, 128)
        torch.testing.assert_close(pe, g["prompt_embeds"], rtol=1e-5, atol=1e-6)
        if case["cfg"]:
            torch.testing.assert_close(npe, g["negative_prompt_embeds"], rtol=1e-5, atol=1e-6)
        else:
            assert npe is None and g["negative_prompt_embeds"] is None


def test_product_tokenizer_stub_matches_golden_ids(golden):
    from comat_b200.synthetic import SyntheticClipTokenizer
    gold = golden("encode_prompt")
    tok = SyntheticClipTokenizer()
    for case in FX.ENCODE_PROMPT_CASES:
        ids = tok(case["prompts"], padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
        assert torch.equal(ids, gold["seed%d" % case["seed"]]["input_ids"])
    t2 = SyntheticClipTokenizer(pad_token_id=0)(["a b"], padding="longest")
    assert t2.input_ids.shape == (1, 4) and t2.attention_mask.sum() == 4


@pytest.mark.parametrize("which", ["clip_l", "bigg"])
def test_text_executor_matches_hf(monkeypatch, which):
    EMU.install_blip(monkeypatch)
    from comat_b200.text_encoder import EngineCLIPText
    model = R.make_clip_text(which, tiny=True, seed=3)
    tok = FX.ClipTokenizerStub()
    t = tok(["a photo of a cat", "two red cubes on a blue sphere next to a green cone", ""])
    with torch.no_grad():
        ref = model(t.input_ids, output_hidden_states=True)
    out = EngineCLIPText(model, torch.float32)(t.input_ids, output_hidden_states=True)
    torch.testing.assert_close(out.last_hidden_state, ref.last_hidden_state, rtol=1e-4, atol=1e-5)
    assert len(out.hidden_states) == len(ref.hidden_states) == 3
    for a, b in zip(out.hidden_states, ref.hidden_states):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    if which == "bigg":
        assert out.keys() == ["text_embeds", "last_hidden_state", "hidden_states"]
        torch.testing.assert_close(out[0], ref.text_embeds, rtol=1e-4, atol=1e-5)
    else:
        assert out.keys() == ["last_hidden_state", "pooler_output", "hidden_states"]
        torch.testing.assert_close(out.pooler_output, ref.pooler_output, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(out[-1][-2], ref.hidden_states[-2], rtol=1e-4, atol=1e-5)
    # key-padding mask on top of the causal mask (configs with use_attention_mask, TrainableSDPipeline.py:312-322)
    with torch.no_grad():
        ref_m = model(t.input_ids, attention_mask=t.attention_mask)
    out_m = EngineCLIPText(model, torch.float32)(t.input_ids, attention_mask=t.attention_mask)
    for i in range(3):
        L = int(t.attention_mask[i].sum())
        torch.testing.assert_close(out_m[0 if which == "clip_l" else 1][i, :L], ref_m.last_hidden_state[i, :L], rtol=1e-4, atol=1e-5)


def _sd_pipe(enc, tok):
    from comat_b200.pipelines import TrainableSDPipeline
    pipe = TrainableSDPipeline.__new__(TrainableSDPipeline)
    TrainableSDPipeline.__init__(pipe, vae=None, unet=None, text_encoder=enc, tokenizer=tok)
    return pipe


def test_sd_encode_prompt_matches_reference_golden(monkeypatch, golden):
    EMU.install_blip(monkeypatch)
    from comat_b200.synthetic import SyntheticClipTokenizer
    from comat_b200.text_encoder import EngineCLIPText
    gold = golden("encode_prompt")
    for case in FX.ENCODE_PROMPT_CASES:
        g = gold["seed%d" % case["seed"]]
        pipe = _sd_pipe(EngineCLIPText(R.make_clip_text("clip_l", tiny=True, seed=case["seed"]), torch.float32), SyntheticClipTokenizer())
        pe, npe = pipe.encode_prompt(case["prompts"], torch.device("cpu"), case["n_per"], case["cfg"], negative_prompt=_neg(case),
                                     clip_skip=case["clip_skip"])
        torch.testing.assert_close(pe, g["prompt_embeds"], rtol=1e-4, atol=1e-5)
        if case["cfg"]:
            torch.testing.assert_close(npe, g["negative_prompt_embeds"], rtol=1e-4, atol=1e-5)
        else:
            assert npe is None


def test_sd_encode_prompt_errors_and_passthrough(monkeypatch):
    EMU.install_blip(monkeypatch)
    from comat_b200.synthetic import SyntheticClipTokenizer
    from comat_b200.text_encoder import EngineCLIPText
    pipe = _sd_pipe(EngineCLIPText(R.make_clip_text("clip_l", tiny=True, seed=1), torch.float32), SyntheticClipTokenizer())
    with pytest.raises(TypeError):                                   # TrainableSDPipeline.py:363-367
        pipe.encode_prompt(["a", "b"], torch.device("cpu"), 1, True, negative_prompt="x")
    with pytest.raises(ValueError):                                  # :370-375
        pipe.encode_prompt(["a", "b"], torch.device("cpu"), 1, True, negative_prompt=["x"])
    # the trainer's null-embedding call (training_script.py:519): '' repeated train_batch_size times, no guidance
    null = pipe.encode_prompt("", torch.device("cpu"), 3, False)[0]
    assert null.shape == (3, 77, 128) and torch.equal(null[0], null[2])
    # pre-computed embeddings skip the encoder entirely
    bare = _sd_pipe(None, None)
    pe, npe = bare.encode_prompt(None, torch.device("cpu"), 2, True, prompt_embeds=torch.ones(1, 77, 8), negative_prompt_embeds=torch.zeros(1, 77, 8))
    assert pe.shape == npe.shape == (2, 77, 8)
    with pytest.raises(NotImplementedError):
        bare.encode_prompt("a", torch.device("cpu"), 1, False)


@pytest.mark.parametrize("force_zeros,negative", [(True, None), (False, None), (True, "ugly")])
def test_sdxl_encode_prompt_matches_oracle(monkeypatch, force_zeros, negative):
    EMU.install_blip(monkeypatch)
    from comat_b200.pipelines import TrainableSDXLPipeline
    from comat_b200.synthetic import SyntheticClipTokenizer
    from comat_b200.text_encoder import EngineCLIPText
    e1, e2 = R.make_clip_text("clip_l", tiny=True, seed=11), R.make_clip_text("bigg", tiny=True, seed=12)
    prompts = ["a brown horse and a white fence", "three yellow birds"]
    ref = R.encode_prompt_sdxl(e1, e2, FX.ClipTokenizerStub(), FX.ClipTokenizerStub(pad_token_id=0), prompts, 2, True,
                               negative_prompt=negative, force_zeros_for_empty_prompt=force_zeros)
    pipe = TrainableSDXLPipeline.__new__(TrainableSDXLPipeline)
    TrainableSDXLPipeline.__init__(pipe, vae=None, unet=None, text_encoder=EngineCLIPText(e1, torch.float32),
                                   tokenizer=SyntheticClipTokenizer(), text_encoder_2=EngineCLIPText(e2, torch.float32),
                                   tokenizer_2=SyntheticClipTokenizer(pad_token_id=0), force_zeros_for_empty_prompt=force_zeros)
    got = pipe.encode_prompt(prompts, device=torch.device("cpu"), num_images_per_prompt=2, do_classifier_free_guidance=True,
                             negative_prompt=negative)
    assert got[0].shape == (4, 77, 256) and got[2].shape == (4, 64)
    for a, b in zip(got, ref):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    if force_zeros and negative is None:
        assert float(got[1].abs().max()) == 0.0 and float(got[3].abs().max()) == 0.0
    # the trainer's call (training_script.py:521): 4-tuple, negatives None without guidance
    null, nneg, pnull, pneg = pipe.encode_prompt("", device=torch.device("cpu"), num_images_per_prompt=3, do_classifier_free_guidance=False)
    assert null.shape == (3, 77, 256) and pnull.shape == (3, 64) and nneg is None and pneg is None


def test_trainer_step_from_prompt_strings_equals_step_from_oracle_embeddings(monkeypatch):
    """training_script.py:513-525 + :575-588: the trainer encodes '' once, the pipeline encodes batch['text'] inside forward."""
    from tests.test_trainer_logic_cpu import _emulate_cuda_only
    _emulate_cuda_only(monkeypatch)
    EMU.install_blip(monkeypatch)
    from comat_b200 import containers as Cn, synthetic
    from comat_b200.caption import Blip, CaptionModelWrapper
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import TrainableSDPipeline
    from comat_b200.text_encoder import EngineCLIPText
    from comat_b200.trainer import CoMatTrainer
    B, S, res = 2, 2, 64
    torch.manual_seed(0)
    unet = Cn.UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=128)
    vae = Cn.AutoencoderKL(block_out_channels=(64, 64, 128, 128))
    unet.requires_grad_(False); vae.requires_grad_(False)
    unet.install_lora(4, up_std=0.05)
    clip = R.make_clip_text("clip_l", tiny=True, seed=21)
    args = synthetic.default_args(pretrain_model_name="sd_1_5", train_batch_size=B, K=1, total_step=S, gan_loss=False, resolution=res, seed=3)
    pipe = TrainableSDPipeline(EngineVAE(vae, torch.float32), EngineUNet(unet, torch.float32), text_encoder=EngineCLIPText(clip, torch.float32),
                               tokenizer=synthetic.SyntheticClipTokenizer())
    tr = CoMatTrainer(args, pipe, CaptionModelWrapper(["Blip"], [1.0], Blip(HFBlipComparator(R.make_blip(large=False)))), None)
    prompts = ["a red apple on a table", "two dogs"]
    g = torch.Generator().manual_seed(9)
    ids, mask = FX.blip_token_batch(g, B, 8)
    base = dict(blip={"input_ids": ids, "attention_mask": mask}, init_latents=torch.randn(B, 4, res // 8, res // 8, generator=g),
                noises=[torch.randn(B, 4, res // 8, res // 8, generator=g) for _ in range(S)], training_steps=[1], attrcon_steps=None, crop=(0, 0))
    l_text = tr.g_losses(dict(base, text=prompts))["loss"]
    pe, _, _ = R.encode_prompt_sd(clip, FX.ClipTokenizerStub(), prompts, 1, False)
    null, _, _ = R.encode_prompt_sd(clip, FX.ClipTokenizerStub(), "", B, False)
    assert tr.null_embed.shape == (B, 77, 128)
    torch.testing.assert_close(tr.null_embed, null, rtol=1e-4, atol=1e-5)
    l_emb = tr.g_losses(dict(base, prompt_embeds=pe, null_embeds=null))["loss"]
    assert abs(float(l_text.detach()) - float(l_emb.detach())) < 1e-4 * abs(float(l_emb.detach()))
    args2 = synthetic.default_args(pretrain_model_name="sd_1_5", train_text_encoder_lora=True)
    with pytest.raises(NotImplementedError):
        CoMatTrainer(args2, pipe, None, None)


@pytest.mark.needs_reference
def test_discriminator_encode_prompt_vs_reference(monkeypatch):
    """gan_sdxl.py:134-155 run verbatim (reference D_sd.encode_prompt over the reference TrainableSDPipeline.encode_prompt, HF CLIP,
    tokenizer stub) vs the product's D_sd.encode_prompt: the '' embedding the trainer asks for once (training_script.py:516), and
    the release of the D pipeline's text encoder afterwards."""
    from types import SimpleNamespace
    from oracle import ref_shim
    EMU.install_blip(monkeypatch)
    ref_shim.install()
    gan = ref_shim.import_reference("training_utils.gan_sdxl")
    pl = ref_shim.import_reference("TrainableSDPipeline")
    clip = R.make_clip_text("clip_l", tiny=True, seed=31)
    moved = []

    class Enc(torch.nn.Module):                                  # records .to('cpu') like the reference parks it (:151)
        def __init__(self):
            super().__init__()
            self.m, self.config = clip, clip.config

        @property
        def dtype(self):
            return torch.float32

        def forward(self, *a, **k):
            return self.m(*a, **k)

        def to(self, *a, **k):
            moved.append(a)
            return self
    rp = pl.TrainableSDPipeline.__new__(pl.TrainableSDPipeline)
    ref_shim._PipelineBase.__init__(rp, vae=SimpleNamespace(to=lambda *a: moved.append(a)), text_encoder=Enc(), tokenizer=FX.ClipTokenizerStub(),
                                    unet=None, scheduler=None)
    self_ns = SimpleNamespace(D_sd_pipeline=rp)
    null_ref, pooled_ref = gan.D_sd.encode_prompt(self_ns, "", torch.device("cpu"), 3, do_classifier_free_guidance=False)
    assert pooled_ref is None and null_ref.shape == (3, 77, 128) and len(moved) == 2
    # product
    from comat_b200.gan import D_sd
    from comat_b200.pipelines import TrainableSDPipeline
    from comat_b200.synthetic import SyntheticClipTokenizer
    from comat_b200.text_encoder import EngineCLIPText
    dp = TrainableSDPipeline(None, None, text_encoder=EngineCLIPText(clip, torch.float32), tokenizer=SyntheticClipTokenizer())
    D = D_sd.__new__(D_sd)
    torch.nn.Module.__init__(D)
    D.D_sd_pipeline = dp
    null, pooled = D.encode_prompt("", torch.device("cpu"), 3, do_classifier_free_guidance=False)
    assert pooled is None and dp.text_encoder is None
    torch.testing.assert_close(null, null_ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("act", ["quick_gelu", "gelu"])
def test_text_lora_taped_executor_matches_hf_autograd(monkeypatch, act):
    """--train_text_encoder_lora building block: taped forward + backward of the CLIP executor with LoRA on q/k/v/out vs torch
    autograd through the HF module carrying the same adapters (forward hooks): output, every d down / d up, and the frozen
    (no-grad) call with adapters present."""
    EMU.install_blip(monkeypatch)
    from transformers import CLIPTextConfig, CLIPTextModel
    from comat_b200.text_encoder import EngineCLIPText, install_text_lora
    torch.manual_seed(0)
    model = CLIPTextModel(CLIPTextConfig(vocab_size=49408, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                                         max_position_embeddings=77, hidden_act=act, eos_token_id=2, bos_token_id=0, pad_token_id=1)).eval()
    model.requires_grad_(False)
    params = install_text_lora(model, 4, up_std=0.05)
    assert len(params) == 2 * 4 * 2 and all(p.requires_grad and p.dtype == torch.float32 for p in params)
    hooks = R.add_text_lora_hooks(model)
    t = FX.ClipTokenizerStub()(["a photo of a cat", "two red cubes on a blue sphere", ""])
    g = torch.Generator().manual_seed(3)
    dy = torch.randn(3, 77, 128, generator=g)
    ref = model(t.input_ids).last_hidden_state
    g_ref = torch.autograd.grad(ref, params, dy)
    for h in hooks:
        h.remove()
    enc = EngineCLIPText(model, torch.float32)
    assert [id(p) for p in enc.lora_parameters()] == [id(p) for p in params]
    out = enc(t.input_ids)
    last = out[0]
    assert last.requires_grad
    torch.testing.assert_close(last, ref, rtol=1e-4, atol=1e-5)
    got = torch.autograd.grad(last, params, dy)
    for a, b in zip(got, g_ref):
        assert float((a - b).norm() / b.norm().clamp_min(1e-12)) < 1e-4
    with torch.no_grad():                                        # adapters present, no gradient wanted
        frozen = enc(t.input_ids)[0]
    assert not frozen.requires_grad
    torch.testing.assert_close(frozen, ref.detach(), rtol=1e-4, atol=1e-5)
    with pytest.raises(NotImplementedError):
        enc(t.input_ids, output_hidden_states=True)
    # optimiser step -> refresh: the executor follows the new masters
    with torch.no_grad():
        for p in params:
            p.mul_(1.5)
    enc.refresh_lora()
    hooks = R.add_text_lora_hooks(model)
    with torch.no_grad():
        torch.testing.assert_close(enc(t.input_ids)[0], model(t.input_ids).last_hidden_state, rtol=1e-4, atol=1e-5)


def test_trainer_text_lora_gradients_equal_the_chain_rule(monkeypatch):
    """--train_text_encoder_lora end to end on emulated ops: one ``loss.backward()`` through trainer -> pipeline -> UNet executor
    (d encoder_hidden_states) -> taped text encoder must give the text-LoRA gradients that the chain rule gives when the same
    step is cut at the embeddings (d loss / d embeddings from a run with leaf embeddings, pushed through the encoder alone)."""
    from tests.test_trainer_logic_cpu import _emulate_cuda_only
    _emulate_cuda_only(monkeypatch)
    EMU.install_blip(monkeypatch)
    from comat_b200 import containers as Cn, synthetic
    from comat_b200.caption import Blip, CaptionModelWrapper
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import TrainableSDPipeline
    from comat_b200.text_encoder import EngineCLIPText, install_text_lora
    from comat_b200.trainer import CoMatTrainer
    B, S, res = 2, 2, 64
    torch.manual_seed(0)
    unet = Cn.UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=128)
    vae = Cn.AutoencoderKL(block_out_channels=(64, 64, 128, 128))
    unet.requires_grad_(False); vae.requires_grad_(False)
    unet.install_lora(4, up_std=0.05)
    clip = R.make_clip_text("clip_l", tiny=True, seed=21)
    tparams = install_text_lora(clip, 4, up_std=0.05)
    args = synthetic.default_args(pretrain_model_name="sd_1_5", train_batch_size=B, K=1, total_step=S, gan_loss=False, resolution=res, seed=3,
                                  train_text_encoder_lora=True)
    enc = EngineCLIPText(clip, torch.float32)
    pipe = TrainableSDPipeline(EngineVAE(vae, torch.float32), EngineUNet(unet, torch.float32), text_encoder=enc,
                               tokenizer=synthetic.SyntheticClipTokenizer())
    tr = CoMatTrainer(args, pipe, CaptionModelWrapper(["Blip"], [1.0], Blip(HFBlipComparator(R.make_blip(large=False)))), None)
    assert tr.train_text and len(tr.G_parameters) == 256 + len(tparams) and tr.optimizer.n == sum(p.numel() for p in tr.G_parameters)
    prompts = ["a red apple on a table", "two dogs"]
    g = torch.Generator().manual_seed(9)
    ids, mask = FX.blip_token_batch(g, B, 8)
    base = dict(blip={"input_ids": ids, "attention_mask": mask}, init_latents=torch.randn(B, 4, res // 8, res // 8, generator=g),
                noises=[torch.randn(B, 4, res // 8, res // 8, generator=g) for _ in range(S)], training_steps=[1], attrcon_steps=None, crop=(0, 0))
    # (1) one backward through everything
    tr.optimizer.zero_grad()
    loss = tr.g_losses(dict(base, text=prompts))["loss"]
    loss.backward()
    tr.pipeline.unet.finalize_lora_grads()
    got = [p.grad.clone() for p in tparams]
    assert max(float(x.abs().max()) for x in got) > 0
    # (2) the same step cut at the embeddings
    tok = synthetic.SyntheticClipTokenizer()
    pe = enc(tok(prompts).input_ids)[0]
    null = enc(tok([""] * B).input_ids)[0]
    pe_leaf, null_leaf = pe.detach().requires_grad_(True), null.detach().requires_grad_(True)
    tr.optimizer.zero_grad()
    # embeddings come in as leaves (train_text stays on so the pipeline keeps autograd enabled around the embedding plumbing)
    loss2 = tr.g_losses(dict(base, prompt_embeds=pe_leaf, null_embeds=null_leaf))["loss"]
    assert abs(float(loss2.detach()) - float(loss.detach())) < 1e-5 * abs(float(loss.detach()))
    d_pe, d_null = torch.autograd.grad(loss2, [pe_leaf, null_leaf])
    want = torch.autograd.grad([pe, null], tparams, [d_pe, d_null])
    for a, b in zip(got, want):
        assert float((a - b).norm()) <= 2e-3 * float(b.norm()) + 1e-9
    # (3) a full step moves the text adapters and refreshes the executor's operand images
    before = [p.detach().clone() for p in tparams]
    tr.train_step(dict(base, text=prompts))
    assert any(not torch.equal(a, b) for a, b in zip(before, tparams))
    l0 = enc.engine.loras[0]
    assert torch.equal(l0.down16.float(), l0.down.detach().to(l0.down16.dtype).float())
    # refusals
    with pytest.raises(NotImplementedError):
        CoMatTrainer(synthetic.default_args(pretrain_model_name="sd_1_5", tune_text_encoder=True), pipe, None, None)
    with pytest.raises(NotImplementedError):
        CoMatTrainer(synthetic.default_args(pretrain_model_name="sd_1_5", train_text_encoder_lora=True, textenc_lora_lr=1e-6), pipe, None, None)


def test_text_lora_checkpoint_entries_round_trip(tmp_path, monkeypatch):
    """training_script.py:386-401,182-185: text-encoder LoRA entries share pytorch_lora_weights.safetensors with the UNet's, under the
    ``text_encoder.`` prefix and diffusers' ``...self_attn.<proj>.lora_linear_layer.<down|up>.weight`` names (un-vendored, restated)."""
    EMU.install_blip(monkeypatch)
    from safetensors import safe_open
    from comat_b200 import checkpoint as CK, synthetic
    from comat_b200.text_encoder import EngineCLIPText, install_text_lora
    a, b = R.make_clip_text("clip_l", tiny=True, seed=1), R.make_clip_text("clip_l", tiny=True, seed=1)
    install_text_lora(a, 4, up_std=0.05)
    torch.manual_seed(9)
    install_text_lora(b, 4, up_std=0.01)
    unet, _ = synthetic.build_sd15("cpu", torch.float32, rank=4, seed=1, tiny=True)
    ea, eb = EngineCLIPText(a, torch.float32), EngineCLIPText(b, torch.float32)
    sd = CK.text_encoder_lora_state_dict(ea)
    assert len(sd) == 2 * 4 * 2 and "text_model.encoder.layers.1.self_attn.out_proj.lora_linear_layer.up.weight" in sd
    path = CK.save_lora_weights(str(tmp_path), CK.unet_lora_state_dict(unet), text_encoder_lora_layers=sd)
    with safe_open(path, "pt") as f:
        names = list(f.keys())
    assert sum(n.startswith("text_encoder.text_model.") for n in names) == 16 and sum(n.startswith("unet.unet.") for n in names) == 256
    assert len(CK.lora_state_dict(str(tmp_path))) == 256                       # the UNet reader skips the text entries
    assert CK.load_lora_into_text_encoder(CK.text_lora_state_dict(str(tmp_path)), eb) == 16
    for x, y in zip(ea.lora_parameters(), eb.lora_parameters()):
        assert torch.equal(x, y)
    ids = FX.ClipTokenizerStub()(["a cat"]).input_ids
    with torch.no_grad():
        torch.testing.assert_close(ea(ids)[0], eb(ids)[0])                         # the executor of b follows the loaded adapters
    with pytest.raises(KeyError):
        CK.load_lora_into_text_encoder({"text_model.encoder.layers.9.self_attn.q_proj.lora_linear_layer.up.weight": torch.zeros(1)}, eb)
