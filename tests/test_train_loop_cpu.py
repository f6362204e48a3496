"""CPU: the training entry point (comat_b200/train.py - the loop of training_script.py:496-735): epochs / max_train_steps,
checkpoint cadence, resume arithmetic (:287-288, :546-549), per-step jsonl logs, the lr schedule, and the data-parallel prompt
sharding of SURVEY 8e.  CUDA ops emulated in torch (test infrastructure)."""
import json
import os

import pytest
import torch


def test_sharded_batches_partition_the_epoch():
    from comat_b200.data import PromptDataset, ShardedBatches
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "prompts.txt")
        open(p, "w").write("\n".join("prompt %d" % i for i in range(23)) + "\n")
        ds = PromptDataset(p)
        assert len(ds) == 23 and ds[3] == {"text": "prompt 3"}
        loaders = [ShardedBatches(ds, 2, r, 4, seed=5) for r in range(4)]
        assert {len(l) for l in loaders} == {2}                                    # 23 // (2*4): same count on every rank
        seen = [t for l in loaders for b in l for t in b["text"]]
        assert len(seen) == 16 and len(set(seen)) == 16                            # disjoint across ranks within an epoch
        e0 = [b["text"] for b in loaders[1]]
        loaders[1].set_epoch(1)
        assert [b["text"] for b in loaders[1]] != e0                               # reshuffled per epoch
        loaders[1].set_epoch(0)
        assert [b["text"] for b in loaders[1]] == e0                               # and reproducible
        js = os.path.join(td, "prompts.json")
        json.dump(["a", "b", "c"], open(js, "w"))
        assert PromptDataset(js, max_train_samples=2)[1] == {"text": "b"}


def test_lr_schedules():
    from comat_b200 import synthetic
    from comat_b200.train import lr_at
    a = synthetic.default_args(lr_scheduler="constant", lr_warmup_steps=0, max_train_steps=10)
    assert [lr_at(a, s) for s in (0, 5, 9)] == [1.0, 1.0, 1.0]
    a = synthetic.default_args(lr_scheduler="constant_with_warmup", lr_warmup_steps=4, max_train_steps=10)
    assert [lr_at(a, s) for s in (0, 2, 4, 9)] == [0.0, 0.5, 1.0, 1.0]
    a = synthetic.default_args(lr_scheduler="linear", lr_warmup_steps=2, max_train_steps=10)
    assert lr_at(a, 1) == 0.5 and lr_at(a, 2) == 1.0 and abs(lr_at(a, 6) - 0.5) < 1e-12 and lr_at(a, 10) == 0.0
    a = synthetic.default_args(lr_scheduler="cosine", lr_warmup_steps=0, max_train_steps=10)
    assert lr_at(a, 0) == 1.0 and abs(lr_at(a, 5) - 0.5) < 1e-12
    # accelerate's prepared scheduler advances num_processes times per optimiser step (training_script.py:324-330)
    assert lr_at(a, 2, world=2) == lr_at(a, 4) and lr_at(a, 3, world=8) == lr_at(a, 24)


def test_training_loop_checkpoints_logs_and_resume(tmp_path, monkeypatch):
    from tests.test_trainer_logic_cpu import _emulate_cuda_only
    from tests import cpu_ops_emulation as EMU
    _emulate_cuda_only(monkeypatch)
    EMU.install_blip(monkeypatch)
    from comat_b200 import checkpoint as CK, synthetic
    from comat_b200.train import Trainer
    prompts = tmp_path / "prompts.txt"
    prompts.write_text("\n".join(["a red apple", "two dogs on a sofa", "a blue car", "snow on a hill", "a green bench"]) + "\n")
    out = str(tmp_path / "run")

    def mk(max_steps, resume):
        a = synthetic.default_args(pretrain_model_name="sd_1_5_attrcon", train_batch_size=2, K=1, total_step=2, resolution=64,
                                   training_prompts=str(prompts), output_dir=out, max_train_steps=max_steps, validation_steps=2,
                                   resume_from_checkpoint=resume, seed=3, attrcon_train_steps=1, lr_scheduler="constant_with_warmup", gradient_accumulation_steps=1,
                                   lr_warmup_steps=2, validation_prompts=["a cat on a mat", "two birds"], num_validation_images=2)
        # 8x8 latent: the only captured cross-attention place of the tiny UNet within reses (64, 32, 16, 8) is up_8
        return Trainer(a, None, torch.device("cpu"), weights="synthetic_tiny", dtype=torch.float32, train_layer_ls=["up_8"])
    tr = mk(3, None)
    assert tr.steps_per_epoch == 2 and tr.args.num_train_epochs == 2
    assert tr.train() == 3
    assert sorted(d for d in os.listdir(out) if d.startswith("checkpoint")) == ["checkpoint-2", "checkpoint-3"]
    logs = [json.loads(l) for l in open(os.path.join(out, "train_log.jsonl"))]
    assert [l["step"] for l in logs] == [1, 2, 3] and all("step_loss" in l and "Blip" in l and "token_loss" in l and "pixel_loss" in l for l in logs)
    assert [l["lr"] for l in logs] == [0.0, tr.args.learning_rate * 0.5, tr.args.learning_rate]
    val = sorted(os.listdir(os.path.join(out, "validation", "step-2")))          # :456-489 at the validation_steps cadence
    assert val == ["test_0_0.png", "test_0_1.png", "test_1_0.png", "test_1_1.png"]
    from PIL import Image
    assert Image.open(os.path.join(out, "validation", "step-2", val[0])).size == (64, 64)
    tb_dir = os.path.join(out, "logs", "comat")                                    # --report_to tensorboard (default): :105-107, :359
    assert any(f.startswith("events.out.tfevents") for f in os.listdir(tb_dir))
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(tb_dir, size_guidance={"scalars": 0, "images": 0})
    acc.Reload()
    assert {"step_loss", "Blip", "lr", "train_loss"} <= set(acc.Tags()["scalars"]) and [e.step for e in acc.Scalars("step_loss")] == [1, 2, 3]
    assert {"test_0", "test_1"} <= set(acc.Tags()["images"])
    flat3 = tr.core.optimizer.flat.clone()
    # resume: picks checkpoint-3, skips the first batch of epoch 1, runs exactly one more step
    tr2 = mk(4, "latest")
    assert tr2.global_step == 3 and tr2.first_epoch == 1 and tr2.resume_step == 1
    assert torch.equal(tr2.core.optimizer.flat, flat3) and tr2.core.optimizer.step_count == 3
    assert tr2.train() == 4 and tr2.core.optimizer.step_count == 4
    assert CK.latest_checkpoint(out).endswith("checkpoint-4")
    assert [json.loads(l)["step"] for l in open(os.path.join(out, "train_log.jsonl"))] == [1, 2, 3, 4]
    with pytest.raises(NotImplementedError):
        Trainer(synthetic.default_args(full_finetuning=True, training_prompts=str(prompts)), {}, torch.device("cpu"))
    too_big = synthetic.default_args(pretrain_model_name="sd_1_5", train_batch_size=8, training_prompts=str(prompts), output_dir=out,
                                     resume_from_checkpoint=None, seed=3)
    with pytest.raises(ValueError):                     # 5 prompts cannot fill a batch of 8: refuse instead of "training" 0 steps
        Trainer(too_big, None, torch.device("cpu"), weights="synthetic_tiny", dtype=torch.float32)


def test_sdxl_entry_trains_from_prompt_strings(tmp_path, monkeypatch):
    """configs[3] shape of the entry point at tiny geometry: SDXL UNet with pooled-text / time-id conditioning fed by BOTH text
    encoders from prompt strings, SD1.5 discriminator with its own CLIP-L null embedding (scripts/sdxl.sh:15), attrcon losses."""
    from tests.test_trainer_logic_cpu import _emulate_cuda_only
    from tests import cpu_ops_emulation as EMU
    _emulate_cuda_only(monkeypatch)
    EMU.install_blip(monkeypatch)
    from comat_b200 import gan_data as GD, synthetic
    from comat_b200.train import Trainer
    # GAN data on disk in the reference's format (gan_dataset.py): jsonl index + latents/*.pt
    idx = tmp_path / "train_data" / "gan_train_data.jsonl"
    os.makedirs(tmp_path / "train_data" / "latents")
    g = torch.Generator().manual_seed(1)
    with open(idx, "w") as f:
        for i, p in enumerate(["a red apple", "two dogs on a sofa", "a blue car"]):
            path = str(tmp_path / "train_data" / "latents" / f"{GD.short_uid()}.pt")
            torch.save(torch.randn(4, 16, 16, generator=g), path)
            f.write(json.dumps({"prompt": p, "file_path": path}) + "\n")
    a = synthetic.default_args(pretrain_model_name="sdxl_attrcon", train_batch_size=1, K=1, total_step=2, resolution=128, gan_loss=True,   # 16x16 latent: up_8 captured
                               gan_model_arch="gansd_1_5", training_prompts=str(idx), output_dir=str(tmp_path / "run"), max_train_steps=2,
                               validation_steps=100, resume_from_checkpoint=None, seed=5, attrcon_train_steps=1, gradient_accumulation_steps=1)
    tr = Trainer(a, None, torch.device("cpu"), weights="synthetic_tiny", dtype=torch.float32, train_layer_ls=["up_8"])
    assert tr.pipeline.is_sdxl and tr.D is not None and len(tr.dataset) == 3
    assert tr.train() == 2
    assert tr.core.null_embed.shape == (1, 77, 64) and tr.core.pooled_null_embed.shape == (1, 16)        # 32 | 32 context, 16 pooled
    assert tr.core.gan_null_embed.shape == (1, 77, 128) and tr.D.D_sd_pipeline.text_encoder is None       # D's own CLIP-L, released
    logs = [json.loads(l) for l in open(os.path.join(a.output_dir, "train_log.jsonl"))]
    assert len(logs) == 2 and all(k in logs[0] for k in ("Blip", "G_loss", "D_loss", "step_loss"))
    assert sorted(os.listdir(os.path.join(a.output_dir, "checkpoint-2"))) == ["D_sd", "pytorch_lora_weights.safetensors", "trainer_state.pt"]


def test_gradient_accumulation_sums_scaled_micro_batch_gradients(tmp_path, monkeypatch):
    """accelerator.accumulate semantics (training_script.py:556,658-664): two micro-batches with accumulation 2 leave
    (g1 + g2) / 2 in the flat gradient buffer and take ONE optimiser step; the loop counts optimiser steps."""
    from tests.test_trainer_logic_cpu import _emulate_cuda_only
    from tests import cpu_ops_emulation as EMU
    _emulate_cuda_only(monkeypatch)
    EMU.install_blip(monkeypatch)
    from comat_b200 import synthetic
    from comat_b200.train import Trainer
    prompts = tmp_path / "prompts.txt"
    prompts.write_text("\n".join(["a red apple", "two dogs on a sofa", "a blue car", "snow on a hill", "a green bench"]) + "\n")

    def mk(accum, out):
        a = synthetic.default_args(pretrain_model_name="sd_1_5", train_batch_size=1, K=1, total_step=2, resolution=64,
                                   training_prompts=str(prompts), output_dir=str(tmp_path / out), max_train_steps=2, validation_steps=100,
                                   resume_from_checkpoint=None, seed=3, gradient_accumulation_steps=accum)
        return Trainer(a, None, torch.device("cpu"), weights="synthetic_tiny", dtype=torch.float32)
    tr = mk(2, "acc")
    assert tr.steps_per_epoch == 3                     # ceil(5 / 2): the dataloader's tail batch syncs on its own
    g = torch.Generator().manual_seed(1)
    mkb = lambda t: dict(text=[t], init_latents=torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(len(t))),
                         noises=[torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(7 + i)) for i in range(2)],
                         training_steps=[1], attrcon_steps=None, crop=(0, 0))
    b1, b2 = mkb("a red apple"), mkb("two dogs on a sofa")
    core = tr.core
    core.overlap_updates = False
    singles = []
    for b in (b1, b2):
        core.optimizer.zero_grad()
        core.g_losses(b)["loss"].backward()
        core.pipeline.unet.finalize_lora_grads()
        singles.append(core.optimizer.grad.clone())
    steps = []
    monkeypatch.setattr(type(core), "_g_update", lambda self: (self.pipeline.unet.finalize_lora_grads(), steps.append(self.optimizer.grad.clone())))
    core.train_step(b1, accum_steps=2, first=True, last=False)
    assert steps == [] and core.global_step == 0
    core.train_step(b2, accum_steps=2, first=False, last=True)
    assert len(steps) == 1 and core.global_step == 1
    want = 0.5 * (singles[0] + singles[1])
    assert float((steps[0] - want).abs().max()) <= 1e-5 * float(want.abs().max())
    monkeypatch.undo()
    _emulate_cuda_only(monkeypatch)
    EMU.install_blip(monkeypatch)
    tr2 = mk(2, "acc2")
    assert tr2.train() == 2 and tr2.core.optimizer.step_count == 2


@pytest.mark.parametrize("script,name,batch,lr_d", [("sd15.sh", "sd_1_5_attrcon", 4, 2e-5), ("sdxl.sh", "sdxl_attrcon_unet", 6, 5e-5)])
def test_canonical_scripts_parse_with_the_entry_point(script, name, batch, lr_d):
    """scripts/*.sh carry the reference's canonical hyper-parameters (its scripts/sd15.sh, scripts/sdxl.sh): every flag parses."""
    import argparse
    import shlex
    from tests.conftest import ROOT
    from comat_b200.arguments import parse_args
    text = open(os.path.join(ROOT, "scripts", script)).read().replace("\\\n", " ")
    cmd = next(l for l in text.splitlines() if l.startswith("torchrun"))
    argv = shlex.split(cmd.replace('"$@"', "").replace('"${WEIGHTS:-synthetic}"', "synthetic"))
    argv = argv[argv.index("comat_b200.train") + 1:]
    pre = argparse.ArgumentParser(add_help=False)
    pre.add_argument("--weights")
    _, rest = pre.parse_known_args(argv)
    a = parse_args(rest)
    assert a.pretrain_model_name == name and a.train_batch_size == batch and a.learning_rate_D == lr_d
    assert a.K == 5 and a.total_step == 50 and a.lora_rank == 128 and a.gan_loss and a.gan_model_arch == "gansd_1_5"
    assert a.gradient_accumulation_steps == 1 and a.attrcon_train_steps == 2 and a.mixed_precision == "fp16" and a.seed == 42


def test_compute_dtype_follows_mixed_precision_flag():
    from comat_b200 import synthetic
    from comat_b200.train import compute_dtype
    assert compute_dtype(synthetic.default_args(mixed_precision="fp16")) == torch.float16
    assert compute_dtype(synthetic.default_args(mixed_precision="bf16")) == torch.bfloat16
    assert compute_dtype(synthetic.default_args()) == torch.float16                  # unset: node8.yaml's fp16
    for bad in (dict(mixed_precision="no"), dict(use_8bit_adam=True), dict(optimizer_class="Lion")):
        with pytest.raises(NotImplementedError):
            compute_dtype(synthetic.default_args(**bad))


def test_entry_point_with_text_encoder_lora(tmp_path, monkeypatch):
    """--train_text_encoder_lora through the entry point: adapters installed before packing, trained with the UNet LoRA, written
    to / restored from the shared safetensors file."""
    from tests.test_trainer_logic_cpu import _emulate_cuda_only
    from tests import cpu_ops_emulation as EMU
    _emulate_cuda_only(monkeypatch)
    EMU.install_blip(monkeypatch)
    from safetensors import safe_open
    from comat_b200 import synthetic
    from comat_b200.train import Trainer
    prompts = tmp_path / "prompts.txt"
    prompts.write_text("a red apple\ntwo dogs on a sofa\na blue car\n")
    out = str(tmp_path / "run")

    def mk(steps, resume):
        a = synthetic.default_args(pretrain_model_name="sd_1_5", train_batch_size=1, K=1, total_step=2, resolution=64, training_prompts=str(prompts),
                                   output_dir=out, max_train_steps=steps, validation_steps=100, resume_from_checkpoint=resume, seed=3,
                                   gradient_accumulation_steps=1, train_text_encoder_lora=True)
        return Trainer(a, None, torch.device("cpu"), weights="synthetic_tiny", dtype=torch.float32)
    tr = mk(2, None)
    text = tr.core.text_parameters
    assert len(text) == 2 * 4 * 2 and tr.core.optimizer.n == sum(p.numel() for p in tr.core.G_parameters)
    up0 = [p.detach().clone() for p in text[1::2]]               # the zero-initialised `up` factors
    assert tr.train() == 2
    assert any(float((a - b).abs().max()) > 0 for a, b in zip(up0, text[1::2]))
    with safe_open(os.path.join(out, "checkpoint-2", "pytorch_lora_weights.safetensors"), "pt") as f:
        assert sum(k.startswith("text_encoder.") for k in f.keys()) == 16
    tr2 = mk(3, "latest")
    assert tr2.global_step == 2 and all(torch.equal(a, b) for a, b in zip(text, tr2.core.text_parameters))
