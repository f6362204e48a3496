"""Training entry point - the loop of ``Trainer.train`` (training_script.py:496-735) around ``CoMatTrainer.train_step``:
epochs over the prompt data, resume (``--resume_from_checkpoint``, :156-205), per-step logs, ``checkpoint-<n>`` every
``--validation_steps`` (:711-717), ``--max_train_steps``; one more checkpoint is written when training ends (an addition: the
reference only saves on the ``--validation_steps`` cadence, :720-722).

    torchrun --nproc-per-node 8 -m comat_b200.train --pretrain_model_name sd_1_5_attrcon --gan_loss ... [reference flags]

takes the reference's 72 flags unchanged (``comat_b200.arguments``) plus ``--weights synthetic|synthetic_tiny`` - this image has
no Hub access, diffusers or tokenizer vocabularies, so the entry point builds random-init networks at the real geometry and the
stand-in tokenizers of ``comat_b200.synthetic``; a deployment passes real modules through ``Trainer(args, components=...)``
(see INTEGRATION.md).  What the reference obtains from Grounded-SAM + spaCy per batch (noun / attribute token lists, masks:
training_script.py:627-637) comes from ``components['attr_provider'](prompts, images) -> (words, masks)``, called by the step on
the image it has just generated; the synthetic entry uses the SURVEY 8d generator.  Logs go to ``<output_dir>/train_log.jsonl`` (one JSON object per optimiser step) and, with
``--report_to tensorboard`` (the default), to ``<output_dir>/<logging_dir>/<tracker_project_name>`` like accelerate's tracker; scalars are read back once per ``--log_every`` steps, not 7 ``.item()`` syncs per step.
"""
from __future__ import annotations

import os as _os
# multi-rank runs: eager CUDA module loading (lazy loading of a kernel variant while a peer-waiting NCCL kernel is in flight dead-locks;
# see bench.py).  Effective only if the CUDA context does not exist yet.
if int(_os.environ.get("WORLD_SIZE", "1")) > 1:
    _os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import argparse
import json
import math
import os
import random
import sys
import time
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import checkpoint as CK
from . import synthetic
from .arguments import parse_args
from .data import ShardedBatches, get_dataset


def lr_at(args, step: int, world: int = 1) -> float:
    """multiplier schedule of diffusers' ``get_scheduler(args.lr_scheduler, num_warmup_steps, num_training_steps)``
    (training_script.py:289-294; un-vendored, the published definitions) after ``step`` optimiser steps.  The reference passes the
    scheduler through ``accelerator.prepare`` (:324-330): accelerate's AcceleratedScheduler then advances it ``num_processes`` times
    per optimiser step (split_batches = False), so on N GPUs warm-up and decay run N x faster - mirrored by ``world``.  Both shipped
    scripts use ``constant``, where this makes no difference."""
    name, warm, total = args.lr_scheduler, args.lr_warmup_steps, max(1, args.max_train_steps)
    step = step * max(1, int(world))
    if name == "constant":
        return 1.0
    if step < warm:
        return step / max(1, warm)
    if name == "constant_with_warmup":
        return 1.0
    prog = (step - warm) / max(1, total - warm)
    if name == "linear":
        return max(0.0, 1.0 - prog)
    if name == "cosine":
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))
    raise NotImplementedError(f"lr_scheduler {name!r}")


def compute_dtype(args) -> torch.dtype:
    """``--mixed_precision`` -> the 16-bit type of the frozen weights / activations (training_script.py:122-127; node8.yaml runs fp16).
    The flag defaults to None (accelerate then reads its own config, fp16 in node8.yaml:8); 'no' = an fp32 network has no tensor-core
    path in this package and is refused, as are optimisers other than AdamW (:216-226)."""
    if args.mixed_precision == "no":
        raise NotImplementedError("--mixed_precision no: the executors run 16-bit operands on the tensor cores (fp16 or bf16)")
    if getattr(args, "use_8bit_adam", False) or getattr(args, "optimizer_class", "AdamW") != "AdamW":
        raise NotImplementedError("only AdamW (the fused clip + AdamW kernel) is implemented; --use_8bit_adam / other --optimizer_class refused")
    return torch.bfloat16 if args.mixed_precision == "bf16" else torch.float16


def synthetic_components(args, device, dtype=torch.float16, tiny: bool = False, with_caption: bool = True) -> Dict:
    """random-init networks at the real (or tiny) geometry + stand-in tokenizers (no Hub access in this image)."""
    from . import containers as Cn
    from .blip_engine import BlipEngine
    from .caption import Blip, CaptionModelWrapper
    from .gan import load_discriminator
    from .modules import EngineUNet, EngineVAE
    from . import pipelines as P
    from .text_encoder import EngineCLIPText
    name = args.pretrain_model_name
    sdxl = "sdxl" in name
    seed = args.seed if args.seed is not None else 42
    rank = 8 if tiny else args.lora_rank

    def tiny_unet(sd):            # tiny SD1.5 geometry whose text-context width matches the tiny CLIP tower (128)
        torch.manual_seed(sd)
        with torch.device(device):
            u = Cn.UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=128)
        u.requires_grad_(False)
        u.install_lora(rank)
        return u
    torch.manual_seed(seed + 6)
    with torch.device(device):
        vae = Cn.AutoencoderKL(**(dict(block_out_channels=(64, 64, 128, 128)) if tiny else {}),
                               scaling_factor=0.13025 if sdxl else 0.18215)          # sdxl-vae / sd-vae scaling factors
    vae.requires_grad_(False)
    if sdxl:
        # SDXL: UNet 2.57 B, text context = CLIP-L penultimate (768) | OpenCLIP-bigG penultimate (1280), pooled = bigG projection
        unet = synthetic.build_sdxl_unet(device, rank=rank, seed=seed, tiny=tiny)
        small = dict(hidden_size=32, intermediate_size=64, projection_dim=16) if tiny else {}     # 32 + 32 = tiny context 64
        clip = synthetic.build_clip_text(device, torch.float32, seed=seed + 1, which="clip_l", tiny=tiny, **small)
        clip2 = synthetic.build_clip_text(device, torch.float32, seed=seed + 4, which="bigg", tiny=tiny, **small)
        cls = P.AttrConcenTrainableSDXLPipeline if "attrcon" in name else P.TrainableSDXLPipeline
        pipe = cls(EngineVAE(vae, dtype), EngineUNet(unet, dtype), text_encoder=EngineCLIPText(clip, dtype),
                   tokenizer=synthetic.SyntheticClipTokenizer(), text_encoder_2=EngineCLIPText(clip2, dtype),
                   tokenizer_2=synthetic.SyntheticClipTokenizer(pad_token_id=0))
    else:
        unet = tiny_unet(seed) if tiny else synthetic.build_sd15(device, dtype, rank=rank, seed=seed)[0]
        clip = synthetic.build_clip_text(device, torch.float32, seed=seed + 1, which="clip_l", tiny=tiny)
        if getattr(args, "train_text_encoder_lora", False):                           # pipeline.py:117-119 _modify_text_encoder(rank)
            from .text_encoder import install_text_lora
            install_text_lora(clip, rank)
        cls = P.AttrConcenTrainableSDPipeline if "attrcon" in name else P.TrainableSDPipeline
        pipe = cls(EngineVAE(vae, dtype), EngineUNet(unet, dtype), text_encoder=EngineCLIPText(clip, dtype),
                   tokenizer=synthetic.SyntheticClipTokenizer())
    comp = {"pipeline": pipe, "caption_model": None, "D": None}
    if with_caption:
        blip = Blip(BlipEngine(synthetic.build_blip(device, dtype, large=not tiny), dtype), tokenizer=synthetic.SyntheticBertTokenizer())
        comp["caption_model"] = CaptionModelWrapper(list(args.caption_model), list(args.reward_weights), blip)
    if args.gan_loss:
        # the discriminator is an SD1.5 UNet with its own CLIP-L for the '' embedding, also under SDXL (scripts/sdxl.sh:15)
        d_unet = tiny_unet(seed + 2) if tiny else synthetic.build_sd15(device, dtype, rank=rank, seed=seed + 2)[0]
        d_clip = EngineCLIPText(synthetic.build_clip_text(device, torch.float32, seed=seed + 5, which="clip_l", tiny=tiny), dtype) if sdxl else pipe.text_encoder
        d_pipe = P.TrainableSDPipeline(None, None, text_encoder=d_clip, tokenizer=synthetic.SyntheticClipTokenizer())
        D = load_discriminator(args, EngineUNet(d_unet, dtype))
        if D is None:
            raise NotImplementedError(f"--gan_model_arch {args.gan_model_arch!r} resolves to no discriminator (gan_sd_model.py:8-14 "
                                      "strips 'gan' and knows 'sd_1_5' only; the shipped scripts pass gansd_1_5)")
        D.D_sd_pipeline = d_pipe
        comp["D"] = D
    if "attrcon" in name:
        g = torch.Generator().manual_seed(seed + 3)
        rr = random.Random(seed + 3)

        def attr_provider(prompts, images):
            """SURVEY 8d stand-in for Grounded-SAM + spaCy: 1-3 words per prompt with 1-3 token positions each inside the
            prompt's token span, rectangle masks, 10 % empty ('not detected', gsam_interface.py:132-133)."""
            res = images.shape[-1]
            words, masks = [], []
            for _ in prompts:
                nw = rr.randint(1, 3)
                pos = rr.sample(range(1, 40), 9)
                words.append([[pos.pop() for _ in range(rr.randint(1, 3))] for _ in range(nw)])
                m = torch.cat([synthetic.random_mask(g, res, empty=(rr.random() < 0.1)) for _ in range(nw)]).to(images.device)
                masks.append([m[i:i + 1] for i in range(nw)])
            return words, masks
        comp["attr_provider"] = attr_provider
    return comp


def pretrained_components(args, device, dtype=torch.float16, blip_path: Optional[str] = None, d_model_path: Optional[str] = None,
                          attr_provider=None) -> Dict:
    """real checkpoints from LOCAL diffusers / transformers directories (``load_pipeline``, training_utils/pipeline.py:19-82;
    ``load_model``, concept_mat_utils/load_captionmodel.py:3-8; ``D_sd.__init__``, gan_sdxl.py:7-35 - the reference pulls the same
    three from the Hub): ``--pretrain_model`` must be a directory, ``blip_path`` one holding Salesforce/blip-image-captioning-large,
    ``d_model_path`` one holding runwayml/stable-diffusion-v1-5 (defaults to ``--pretrain_model`` for SD1.5 runs)."""
    from transformers import AutoTokenizer, BlipForConditionalGeneration
    from . import pipelines as P
    from .blip_engine import BlipEngine
    from .caption import Blip, CaptionModelWrapper
    from .gan import load_discriminator
    from .loading import load_unet
    name = args.pretrain_model_name
    sdxl = "sdxl" in name
    cls = {(False, False): P.TrainableSDPipeline, (False, True): P.AttrConcenTrainableSDPipeline,
           (True, False): P.TrainableSDXLPipeline, (True, True): P.AttrConcenTrainableSDXLPipeline}[(sdxl, "attrcon" in name)]
    unet = load_unet(args.sdxl_unet_path, device) if (name.endswith("_unet") and args.sdxl_unet_path) else None     # pipeline.py:30-36
    pipe = cls.from_pretrained(args.pretrain_model, revision=args.revision, unet=unet, dtype=dtype, device=device, lora_rank=args.lora_rank)
    if blip_path is None:
        raise ValueError("--blip_path: a local directory with the BLIP captioner (the reference downloads "
                         "Salesforce/blip-image-captioning-large, load_captionmodel.py:5)")
    blip_hf = BlipForConditionalGeneration.from_pretrained(blip_path, local_files_only=True).to(device=device, dtype=dtype).eval()
    blip = Blip(BlipEngine(blip_hf.requires_grad_(False), dtype), tokenizer=AutoTokenizer.from_pretrained(blip_path, local_files_only=True))
    comp = {"pipeline": pipe, "caption_model": CaptionModelWrapper(list(args.caption_model), list(args.reward_weights), blip), "D": None,
            "attr_provider": attr_provider}
    if args.gan_loss:
        d_pipe = P.TrainableSDPipeline.from_pretrained(d_model_path or args.pretrain_model, dtype=dtype, device=device, lora_rank=args.lora_rank)
        D = load_discriminator(args, d_pipe.unet)
        if D is None:
            raise NotImplementedError(f"--gan_model_arch {args.gan_model_arch!r} resolves to no discriminator (gan_sd_model.py:8-14)")
        d_pipe.unet = d_pipe.vae = None                 # D keeps its UNet; the D pipeline only lends its text encoder (gan_sdxl.py:134-155)
        D.D_sd_pipeline = d_pipe
        comp["D"] = D
    return comp


class Trainer:
    """training_script.py:100-735 around the B200 step."""

    def __init__(self, args, components: Optional[Dict] = None, device=None, rank: int = 0, world: int = 1, process_group=None,
                 weights: str = "synthetic", log_every: int = 1, dtype=torch.float16, train_layer_ls=None, blip_path: Optional[str] = None,
                 d_model_path: Optional[str] = None):
        from .pipelines import AttentionStore, register_attention_control
        from .trainer import CoMatTrainer
        self.accum = max(1, int(args.gradient_accumulation_steps))
        if args.full_finetuning or args.tune_vae:
            raise NotImplementedError("--full_finetuning / --tune_vae: this path trains the LoRA factors only")
        self.args, self.rank, self.world, self.log_every = args, rank, world, max(1, log_every)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        if args.seed is not None:                                                    # :117-118 set_seed(args.seed): the SAME seed on every rank
            random.seed(args.seed)                                                   # (step selection and sampler noise streams coincide across
            torch.manual_seed(args.seed)                                             # ranks in the reference; only the prompts differ)
            if self.device.type == "cuda":
                torch.cuda.manual_seed_all(args.seed)
        if components is None and weights == "pretrained":
            components = pretrained_components(args, self.device, dtype, blip_path, d_model_path)
        comp = components or synthetic_components(args, self.device, dtype, tiny=(weights == "synthetic_tiny"))
        self.pipeline, self.caption_model, self.D = comp["pipeline"], comp["caption_model"], comp.get("D")
        self.attr_provider = comp.get("attr_provider")
        if "attrcon" in args.pretrain_model_name:                                     # :307-320
            layers = train_layer_ls or comp.get("train_layer_ls") or (["mid_16", "up_16", "up_32"] if "sdxl" in args.pretrain_model_name
                                                    else ["mid_8", "up_16", "up_32", "up_64"])
            args.train_layer_ls = layers
            register_attention_control(self.pipeline, AttentionStore(layers))
            if self.attr_provider is None:
                raise NotImplementedError("attrcon models need components['attr_provider'] (Grounded-SAM + spaCy are external)")
        self.pipeline.unet.use_graphs = True
        self.core = CoMatTrainer(args, self.pipeline, self.caption_model, self.D, rng=random.Random(args.seed or 0),
                                 process_group=process_group, attr_provider=self.attr_provider)
        self.dataset = get_dataset(args)
        self.loader = ShardedBatches(self.dataset, args.train_batch_size, rank, world, seed=args.seed or 0)
        if len(self.loader) == 0:
            raise ValueError(f"{len(self.dataset)} training prompts do not fill one batch of {args.train_batch_size} on each of "
                             f"{world} rank(s)")
        self.steps_per_epoch = math.ceil(len(self.loader) / self.accum)               # :282 optimiser steps per epoch
        if args.max_train_steps is None:
            args.max_train_steps = args.num_train_epochs * self.steps_per_epoch
        args.num_train_epochs = math.ceil(args.max_train_steps / self.steps_per_epoch)   # :326
        self.global_step = 0
        if args.resume_from_checkpoint:                                               # :156-205
            if args.resume_from_checkpoint == "latest":
                step = CK.load_checkpoint(self.core, args.output_dir, "latest")
            else:
                step = CK.load_checkpoint(self.core, args.resume_from_checkpoint, "explicit")
            if step is None:
                self._print(f"Checkpoint '{args.resume_from_checkpoint}' does not exist. Starting a new training run.")
            else:
                self._print(f"Resuming from checkpoint-{step}")
                self.global_step = step
        self.first_epoch = self.global_step // self.steps_per_epoch                   # :287-288 (resume_step counts micro-batches)
        self.resume_step = (self.global_step * self.accum) % (self.steps_per_epoch * self.accum)
        self._pending = []
        self._micro = 0
        self._log_file = None
        self._tb = None
        if rank == 0:
            os.makedirs(args.output_dir, exist_ok=True)
            self._log_file = open(os.path.join(args.output_dir, "train_log.jsonl"), "a")
            if getattr(args, "report_to", None) == "tensorboard":
                # :105-107, :359: accelerate's tensorboard tracker writes under <output_dir>/<logging_dir>/<tracker_project_name>
                from torch.utils.tensorboard import SummaryWriter
                self._tb = SummaryWriter(os.path.join(args.output_dir, args.logging_dir, args.tracker_project_name))

    def _print(self, *a):
        if self.rank == 0:
            print(*a, file=sys.stderr, flush=True)

    def _barrier(self):
        if self.world > 1 and dist.is_initialized():
            dist.barrier()

    def save(self):
        """:711-717: rank 0 writes ``checkpoint-<global_step>`` (all ranks hold identical parameters under DP)."""
        path = None
        if self.rank == 0:
            path = CK.save_checkpoint(self.core, self.args.output_dir, self.global_step)
        self._barrier()
        return path

    @torch.no_grad()
    def validate(self):
        """the evaluate half of ``save_and_evaluate`` (training_script.py:456-489): ``num_validation_images`` samples per
        validation prompt, one prompt per call, ``num_inference_steps = total_step``, one seeded generator for the whole sweep;
        PNGs under ``<output_dir>/validation/step-<n>/test_<prompt>_<k>.png`` stand in for TensorBoard's ``test_<i>`` image rows."""
        a = self.args
        if self.rank != 0 or not a.validation_prompts or a.num_validation_images <= 0:
            return []
        prompts = list(a.validation_prompts)
        if a.validation_prompts_file is not None:
            with open(a.validation_prompts_file, "r") as f:
                prompts += f.readlines()
        prompts = [p.strip() for p in prompts]
        gen = torch.Generator(device=self.device).manual_seed(a.seed) if a.seed else None
        self.core.sync()
        out_dir = os.path.join(a.output_dir, "validation", f"step-{self.global_step}")
        os.makedirs(out_dir, exist_ok=True)
        paths = []
        for i, p in enumerate(prompts):
            for k in range(a.num_validation_images):
                img = self.pipeline([p], height=a.resolution, width=a.resolution, num_inference_steps=a.total_step, generator=gen,
                                    guidance_scale=a.cfg_scale, guidance_rescale=a.cfg_rescale, output_type="pil").images[0]
                paths.append(os.path.join(out_dir, f"test_{i}_{k}.png"))
                img.save(paths[-1])
            if self._tb is not None:                                                  # :485-489 add_images('test_<i>', NHWC)
                import numpy as np
                from PIL import Image
                row = np.stack([np.asarray(Image.open(q)) for q in paths[-a.num_validation_images:]])
                self._tb.add_images(f"test_{i}", row, self.global_step, dataformats="NHWC")
        return paths

    def _flush_logs(self):
        """ONE device->host transfer for all scalars queued since the last flush."""
        if not self._pending:
            return
        keys = sorted({k for _, d in self._pending for k in d})
        mat = torch.stack([torch.stack([d[k].float().reshape(()) if k in d else torch.full((), float("nan"), device=self.device)
                                        for k in keys]) for _, d in self._pending])
        if self.world > 1 and dist.is_initialized():                                  # :673-683 accelerator.gather(...).mean()
            dist.all_reduce(mat, op=dist.ReduceOp.SUM)
            mat /= self.world
        host = mat.cpu()
        if self._log_file is not None:
            for (step, _), row in zip(self._pending, host):
                rec = {"step": step, "lr": self.args.learning_rate * lr_at(self.args, step - 1, self.world)}
                rec.update({k: float(v) for k, v in zip(keys, row) if not math.isnan(float(v))})
                self._log_file.write(json.dumps(rec) + "\n")
                if self._tb is not None:                                              # :702-703 accelerator.log(..., step=global_step)
                    for k, v in rec.items():
                        if k != "step":
                            self._tb.add_scalar(k, v, step)
                    self._tb.add_scalar("train_loss", rec.get("loss", rec.get("step_loss", 0.0)), step)
            self._log_file.flush()
        self._pending = []

    def _make_batch(self, raw: Dict) -> Dict:
        a = self.args
        batch = dict(raw)
        if a.batch_repeat > 1:                                                        # :552-553
            batch["text"] = batch["text"] * a.batch_repeat
        if "real_latents" in batch:
            batch["real_latents"] = batch["real_latents"].to(self.device, non_blocking=True)
        return batch

    def train(self) -> int:
        a = self.args
        self._print(f"***** Running training *****  examples {len(self.dataset)}  epochs {a.num_train_epochs}  "
                    f"batch/device {a.train_batch_size}  world {self.world}  optimisation steps {a.max_train_steps}")
        t0 = time.time()
        for epoch in range(self.first_epoch, a.num_train_epochs):
            self.loader.set_epoch(epoch)
            for step, raw in enumerate(self.loader):
                if a.resume_from_checkpoint and epoch == self.first_epoch and step < self.resume_step:   # :546-549
                    continue
                if self.global_step >= a.max_train_steps:
                    break
                batch = self._make_batch(raw)
                self.core.optimizer.lr = a.learning_rate * lr_at(a, self.global_step, self.world)  # :663 lr_scheduler.step()
                # accelerator.accumulate (:556): gradients sync every accum-th batch and at the end of the dataloader
                first = self._micro == 0
                last = self._micro == self.accum - 1 or step == len(self.loader) - 1
                logs = self.core.train_step(batch, accum_steps=self.accum, first=first, last=last)
                self._micro = 0 if last else self._micro + 1
                if not last:
                    continue
                self.global_step += 1
                self._pending.append((self.global_step, {k: v.detach() for k, v in logs.items() if torch.is_tensor(v) and v.numel() == 1}))
                if self.global_step % self.log_every == 0:
                    self._flush_logs()
                if self.global_step % a.validation_steps == 0:                        # :711-717
                    self._flush_logs()
                    self.save()
                    self.validate()
            if self.global_step >= a.max_train_steps:
                break
        self.core.close() if hasattr(self.core, "close") else self.core.sync()       # joins the side-stream updates, restores the GC
        self._flush_logs()
        self.save()                                                                   # addition: the reference ends without saving (:720-722)
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        self._print(f"done: {self.global_step} steps in {time.time() - t0:.1f} s")
        if self._log_file is not None:
            self._log_file.close()
        if self._tb is not None:
            self._tb.close()
        return self.global_step


def main(argv=None) -> int:
    pre = argparse.ArgumentParser(add_help=False)
    pre.add_argument("--weights", choices=["synthetic", "synthetic_tiny", "pretrained"], default="synthetic")
    pre.add_argument("--blip_path", type=str, default=None)
    pre.add_argument("--d_model_path", type=str, default=None)
    pre.add_argument("--log_every", type=int, default=10)
    extra, rest = pre.parse_known_args(argv)
    if any(h in rest for h in ("-h", "--help")):
        print("entry-point options (on top of the reference flags below):\n"
              "  --weights {synthetic,synthetic_tiny,pretrained}  random-init networks at the real / a tiny geometry, or local checkpoints:\n"
              "                                         --pretrain_model DIR (diffusers layout), --blip_path DIR, --d_model_path DIR\n"
              "  --log_every N                          read the step scalars back every N steps\n")
    args = parse_args(rest)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("comat_b200.train needs a GPU: the package has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tr = Trainer(args, None, dev, rank, world, weights=extra.weights, log_every=extra.log_every, blip_path=extra.blip_path,
                 d_model_path=extra.d_model_path, dtype=compute_dtype(args))
    tr.train()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
