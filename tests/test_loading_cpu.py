"""CPU: ``PIPELINE.from_pretrained`` over a local diffusers-layout directory (comat_b200/loading.py; training_utils/pipeline.py:19-39):
config.json schema mapping, safetensors / fp16-variant files, legacy VAE attention names, strict key match, and the loaded
pipeline sampling exactly like one built from the same modules in memory (CUDA ops emulated in torch - test infrastructure)."""
import json
import os

import pytest
import torch
from safetensors.torch import save_file

from oracle import comat_ref as R
from tests import cpu_ops_emulation as EMU

SD15_CONFIG_JSON = {          # the fields of runwayml/stable-diffusion-v1-5 unet/config.json that define the geometry
    "_class_name": "UNet2DConditionModel", "act_fn": "silu", "attention_head_dim": 8, "block_out_channels": [320, 640, 1280, 1280],
    "center_input_sample": False, "cross_attention_dim": 768, "down_block_types": ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"],
    "downsample_padding": 1, "flip_sin_to_cos": True, "freq_shift": 0, "in_channels": 4, "layers_per_block": 2, "mid_block_scale_factor": 1,
    "norm_eps": 1e-05, "norm_num_groups": 32, "out_channels": 4, "sample_size": 64, "up_block_types": ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3}
SDXL_CONFIG_JSON = {          # stabilityai/stable-diffusion-xl-base-1.0 unet/config.json
    "_class_name": "UNet2DConditionModel", "addition_embed_type": "text_time", "addition_time_embed_dim": 256, "attention_head_dim": [5, 10, 20],
    "block_out_channels": [320, 640, 1280], "cross_attention_dim": 2048, "down_block_types": ["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"],
    "in_channels": 4, "layers_per_block": 2, "mid_block_type": "UNetMidBlock2DCrossAttn", "num_attention_heads": None, "out_channels": 4,
    "projection_class_embeddings_input_dim": 2816, "transformer_layers_per_block": [1, 2, 10],
    "up_block_types": ["CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"], "use_linear_projection": True, "upcast_attention": None,
    "class_embed_type": None, "encoder_hid_dim": None, "time_cond_proj_dim": None, "conv_in_kernel": 3, "dual_cross_attention": False,
    "only_cross_attention": False}


def test_config_schema_maps_to_the_published_geometries():
    from comat_b200 import containers as Cn
    from comat_b200.loading import unet_kwargs_from_config
    assert unet_kwargs_from_config(SD15_CONFIG_JSON) == Cn.SD15_UNET
    assert unet_kwargs_from_config(SDXL_CONFIG_JSON) == Cn.SDXL_UNET
    with pytest.raises(NotImplementedError):
        unet_kwargs_from_config({**SD15_CONFIG_JSON, "class_embed_type": "timestep"})
    with torch.device("meta"):                                  # and the containers built from them have the published sizes
        n15 = sum(p.numel() for p in Cn.UNet2DConditionModel(**unet_kwargs_from_config(SD15_CONFIG_JSON)).parameters())
        nxl = sum(p.numel() for p in Cn.UNet2DConditionModel(**unet_kwargs_from_config(SDXL_CONFIG_JSON)).parameters())
    assert (n15, nxl) == (859_520_964, 2_567_463_684)


def _write_checkpoint(root, unet, vae, clip, fp16_variant):
    os.makedirs(root / "unet"), os.makedirs(root / "vae")
    cfg = {"in_channels": 4, "out_channels": 4, "block_out_channels": [64, 128, 256, 256], "layers_per_block": 2, "attention_head_dim": 4,
           "cross_attention_dim": 128, "down_block_types": ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"],
           "up_block_types": ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3}
    json.dump(cfg, open(root / "unet" / "config.json", "w"))
    sd = {k: v.contiguous() for k, v in unet.state_dict().items()}
    if fp16_variant:
        save_file({k: v.half() for k, v in sd.items()}, str(root / "unet" / "diffusion_pytorch_model.fp16.safetensors"))
    else:
        save_file(sd, str(root / "unet" / "diffusion_pytorch_model.safetensors"))
    json.dump({"block_out_channels": [64, 64, 128, 128], "scaling_factor": 0.18215, "norm_num_groups": 32, "latent_channels": 4},
              open(root / "vae" / "config.json", "w"))
    legacy = {"to_q": "query", "to_k": "key", "to_v": "value", "to_out.0": "proj_attn"}
    vsd = {}
    for k, v in vae.state_dict().items():                       # pre-refactor attention names + tensors this path never reads
        for new, old in legacy.items():
            k = k.replace(f"attentions.0.{new}.", f"attentions.0.{old}.")
        vsd[k] = v.contiguous()
    vsd["encoder.conv_in.weight"] = torch.zeros(8, 3, 3, 3)
    vsd["quant_conv.weight"] = torch.zeros(8, 8, 1, 1)
    torch.save(vsd, str(root / "vae" / "diffusion_pytorch_model.bin"))
    clip.save_pretrained(str(root / "text_encoder"))
    json.dump({"_class_name": "StableDiffusionPipeline"}, open(root / "model_index.json", "w"))


@pytest.mark.parametrize("fp16_variant", [False, True])
def test_from_pretrained_equals_in_memory_pipeline(tmp_path, monkeypatch, fp16_variant):
    EMU.install_blip(monkeypatch)
    from comat_b200 import containers as Cn, synthetic
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import TrainableSDPipeline
    from comat_b200.text_encoder import EngineCLIPText
    torch.manual_seed(0)
    unet = Cn.UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=128).requires_grad_(False)
    vae = Cn.AutoencoderKL(block_out_channels=(64, 64, 128, 128)).requires_grad_(False)
    clip = R.make_clip_text("clip_l", tiny=True, seed=21)
    if fp16_variant:
        unet.half().float()                                      # what survives the fp16 file
        for p in unet.parameters():
            p.data = p.data.half().float()
    _write_checkpoint(tmp_path, unet, vae, clip, fp16_variant)
    tok = synthetic.SyntheticClipTokenizer()
    loaded = TrainableSDPipeline.from_pretrained(str(tmp_path), revision=None, torch_type=torch.float16, dtype=torch.float32, device="cpu",
                                                 tokenizer=tok, variant="fp16" if fp16_variant else None, lora_rank=4)
    assert len(loaded.unet.lora_parameters()) == 256 and all(p.dtype == torch.float32 for p in loaded.unet.lora_parameters())
    assert not any(k.startswith("encoder") for k in loaded.vae.ref.state_dict())
    unet.install_lora(4)                                         # up = 0: the LoRA branch contributes nothing in both pipelines
    direct = TrainableSDPipeline(EngineVAE(vae, torch.float32), EngineUNet(unet, torch.float32), text_encoder=EngineCLIPText(clip, torch.float32),
                                 tokenizer=tok)
    kw = dict(height=64, width=64, num_inference_steps=2, output_type="pt")
    a = loaded(["a red apple"], generator=torch.Generator().manual_seed(4), **kw).images
    b = direct(["a red apple"], generator=torch.Generator().manual_seed(4), **kw).images
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    with pytest.raises(FileNotFoundError):
        TrainableSDPipeline.from_pretrained(str(tmp_path / "nowhere"))
    # strict: a checkpoint with a missing tensor is refused, not silently half-loaded
    os.remove(tmp_path / "vae" / "diffusion_pytorch_model.bin")
    torch.save({"post_quant_conv.weight": torch.zeros(4, 4, 1, 1)}, str(tmp_path / "vae" / "diffusion_pytorch_model.bin"))
    with pytest.raises(RuntimeError):
        TrainableSDPipeline.from_pretrained(str(tmp_path), dtype=torch.float32, device="cpu", tokenizer=tok, variant="fp16" if fp16_variant else None)


def _write_clip_tokenizer(folder):
    """a real ``transformers.CLIPTokenizer`` over a fabricated vocabulary (letters, digits, one merge) - exercises the HF tokenizer
    protocol (``padding='max_length'``, ``</w>`` word pieces, BOS / EOS strings) that the stand-in tokenizers imitate."""
    import string
    os.makedirs(folder)
    chars = list(string.ascii_lowercase + string.digits)
    vocab = {c: i for i, c in enumerate(chars)}
    vocab.update({c + "</w>": len(chars) + i for i, c in enumerate(chars)})
    for t in ("ab</w>", "<|startoftext|>", "<|endoftext|>"):
        vocab[t] = len(vocab)
    json.dump(vocab, open(os.path.join(folder, "vocab.json"), "w"))
    open(os.path.join(folder, "merges.txt"), "w").write("#version: 0.2\na b</w>\n")
    json.dump({"model_max_length": 77}, open(os.path.join(folder, "tokenizer_config.json"), "w"))


def test_from_pretrained_with_a_real_clip_tokenizer(tmp_path, monkeypatch):
    EMU.install_blip(monkeypatch)
    from transformers import CLIPTextConfig, CLIPTextModel
    from comat_b200 import attr_align as AA, containers as Cn
    from comat_b200.pipelines import TrainableSDPipeline
    torch.manual_seed(0)
    unet = Cn.UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=128).requires_grad_(False)
    vae = Cn.AutoencoderKL(block_out_channels=(64, 64, 128, 128)).requires_grad_(False)
    torch.manual_seed(1)
    clip = CLIPTextModel(CLIPTextConfig(vocab_size=80, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                                        max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2, bos_token_id=0, pad_token_id=1)).eval()
    _write_checkpoint(tmp_path, unet, vae, clip, False)
    _write_clip_tokenizer(str(tmp_path / "tokenizer"))
    pipe = TrainableSDPipeline.from_pretrained(str(tmp_path), dtype=torch.float32, device="cpu")
    assert type(pipe.tokenizer).__name__ in ("CLIPTokenizer", "CLIPTokenizerFast") and pipe.tokenizer.model_max_length == 77
    prompts = ["a cab", "zz 9"]
    pe, npe = pipe.encode_prompt(prompts, torch.device("cpu"), 1, True)
    pe_ref, npe_ref, ids = R.encode_prompt_sd(clip.requires_grad_(False), pipe.tokenizer, prompts, 1, True)
    assert ids.shape == (2, 77) and int(ids[0, 0]) == pipe.tokenizer.bos_token_id
    torch.testing.assert_close(pe, pe_ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(npe, npe_ref, rtol=1e-4, atol=1e-5)
    # the alignment helpers see real CLIP word pieces: 'cab' is split into 'c' + 'ab</w>'
    assert AA.get_attention_map_index_to_wordpiece(pipe.tokenizer, "a cab") == {1: "a", 2: "c", 3: "ab"}
    assert AA.align_wordpieces_indices(AA.get_indices(pipe.tokenizer, "a cab"), 2, "cab") == [2, 3]


def test_training_entry_point_on_local_checkpoints(tmp_path, monkeypatch):
    """``--weights pretrained``: SD pipeline, discriminator and BLIP captioner all read from local directories (diffusers layout /
    transformers ``save_pretrained``), real HF tokenizers (CLIP BPE, BERT word pieces) on the prompt strings; one GAN step runs."""
    from tests.test_trainer_logic_cpu import _emulate_cuda_only
    _emulate_cuda_only(monkeypatch)
    EMU.install_blip(monkeypatch)
    from transformers import CLIPTextConfig, CLIPTextModel
    from comat_b200 import containers as Cn, gan_data as GD, synthetic
    from comat_b200.train import Trainer
    sd = tmp_path / "sd15"
    os.makedirs(sd)
    torch.manual_seed(0)
    unet = Cn.UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=128).requires_grad_(False)
    vae = Cn.AutoencoderKL(block_out_channels=(64, 64, 128, 128)).requires_grad_(False)
    torch.manual_seed(1)
    clip = CLIPTextModel(CLIPTextConfig(vocab_size=80, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                                        max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2, bos_token_id=0, pad_token_id=1)).eval()
    _write_checkpoint(sd, unet, vae, clip, False)
    _write_clip_tokenizer(str(sd / "tokenizer"))
    blip_dir = tmp_path / "blip"
    R.make_blip(large=False).save_pretrained(str(blip_dir))
    words = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "a", "photography", "of", "cab", "zz", "9", "##b"]
    open(blip_dir / "vocab.txt", "w").write("\n".join(words) + "\n")
    json.dump({"tokenizer_class": "BertTokenizer", "do_lower_case": True}, open(blip_dir / "tokenizer_config.json", "w"))
    idx = tmp_path / "train_data" / "gan.jsonl"
    os.makedirs(tmp_path / "train_data" / "latents")
    with open(idx, "w") as f:
        for p in ["a cab", "zz 9"]:
            path = str(tmp_path / "train_data" / "latents" / f"{GD.short_uid()}.pt")
            torch.save(torch.randn(4, 8, 8), path)
            f.write(json.dumps({"prompt": p, "file_path": path}) + "\n")
    a = synthetic.default_args(pretrain_model=str(sd), pretrain_model_name="sd_1_5", train_batch_size=2, K=1, total_step=2, resolution=64,
                               gan_loss=True, gan_model_arch="gansd_1_5", training_prompts=str(idx), output_dir=str(tmp_path / "run"),
                               max_train_steps=1, validation_steps=100, resume_from_checkpoint=None, seed=5, lora_rank=4,
                               gradient_accumulation_steps=1)
    tr = Trainer(a, None, torch.device("cpu"), weights="pretrained", dtype=torch.float32, blip_path=str(blip_dir))
    assert tr.caption_model.blip_model.prompt_length == 4            # len(bert('a photography of').input_ids) - 1 (caption_blip.py:38-39)
    assert tr.train() == 1
    log = json.loads(open(os.path.join(a.output_dir, "train_log.jsonl")).readline())
    assert {"Blip", "G_loss", "D_loss", "step_loss"} <= set(log)
    assert len(tr.core.G_parameters) == 256 and tr.core.gan_null_embed.shape == (2, 77, 128)
