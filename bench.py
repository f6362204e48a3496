#!/usr/bin/env python
"""bench.py — SD1.5 512^2 CoMat train-steps/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W             # product arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the oracle's CPU path on the host cores

Workload (config.workload): BASELINE.json configs[1] — SD1.5 geometry, full CoMat step (BLIP concept-matching reward +
attention-map token/pixel loss on 2 attrcon steps + GAN fidelity G and D updates), S=20 DDPM steps, K=5 back-propagated,
per-GPU batch 4, LoRA rank 128, cfg 7.5; random-init weights (seed 42) and synthetic prompt embeddings / token ids /
masks / real latents (no Hub or dataset access).  One "step" = one G optimiser step + one D optimiser step
(training_script.py:543-719 at gradient_accumulation_steps=1).

value : device-resident inputs, CUDA-event timed, K steps bracketed by barrier+synchronize, max over ranks.
e2e   : the same K steps through the public trainer API with HOST (pinned) inputs: H2D of the step's batch and a D2H
        read of the step loss inside the timed region.
roofline: the tcgen05 GEMM/conv kernel family (dominant), algorithmic FLOPs / per-launch CUDA-event time, from one
        extra instrumented step after the timed region.
"""
from __future__ import annotations

import os as _os
# Multi-rank runs: load every CUDA module eagerly.  With the default lazy loading, the first use of a kernel variant inside an optimiser
# tail (first launches of a GEMM flavour, on the side stream, while the gradient all-reduce of the OTHER optimiser is in flight and
# waits for the peer) blocked the host in the module load on both ranks and the job dead-locked in its first step
# (gpurun_out/r02_bench_2gpu_dbg.err: both ranks parked inside comat_gemm).  Must be set before the CUDA context exists.
if int(_os.environ.get("WORLD_SIZE", "1")) > 1:
    _os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sd15_512_comat_train_steps_per_sec"
UNIT = "train-steps/s"
TFLOP_PER_STEP_CFG2 = 203.0          # SURVEY 8d / BASELINE.md section 3 (per GPU, B=4, S=20, K=5, GAN on)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="comat_b200", choices=["comat_b200", "reference", "gpu_reference"])
    p.add_argument("--no_gpu_reference", action="store_true", help="skip the eager-oracle-on-this-GPU denominator (gpu_reference key)")
    p.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                   help="BASELINE.json configs[N-1]: 2 = SD1.5 full CoMat B=4 (the headline, default); 3 = the same at B=8/GPU; 4 = SDXL 512^2 "
                        "attrcon + GAN (SD1.5 discriminator) B=1/GPU; 5 = SDXL UNet-only forward+backward microbench (--latent, B=2/GPU)")
    p.add_argument("--batch", type=int, default=0, help="per-GPU batch (0 = the config's: 4 / 8 / 1 / 2)")
    p.add_argument("--latent", type=int, default=128, help="config 5: latent side (64 = 512^2, 96 = 768^2, 128 = 1024^2)")
    p.add_argument("--total_step", type=int, default=20)
    p.add_argument("--K", type=int, default=5)
    p.add_argument("--rank_lora", type=int, default=128)
    p.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    p.add_argument("--tiny", action="store_true", help="reduced geometry (debug only; never a reported number)")
    p.add_argument("--no_cpu_baseline", action="store_true")
    p.add_argument("--no_gan", action="store_true")
    p.add_argument("--no_attrcon", action="store_true")
    p.add_argument("--no_graphs", action="store_true", help="disable CUDA-graph replay of the no-grad UNet forwards")
    p.add_argument("--graph_taped", default="auto", choices=["auto", "on", "off"],
                   help="CUDA-graph the back-propagated UNet calls too (forward + backward graph pairs); auto = on where a call is "
                        "configs 2 and 4 (measured: SDXL batch 1 772 -> 481 ms/step, SD1.5 batch 4 486 -> 477 ms/step at +70 GB of resident "
                        "activations), off for config 3 (batch 8: the resident activation sets do not fit)")
    p.add_argument("--kineto_steps", type=int, default=1, help="consecutive steps inside the --kineto_step window")
    p.add_argument("--sync_debug", action="store_true", help="run ONE step with torch.cuda.set_sync_debug_mode('warn') and list the host-sync call sites")
    p.add_argument("--gemm_shapes", default="", help="write the per-shape GEMM table of the instrumented step to this path")
    p.add_argument("--kineto_step", default="", help="profile ONE step with torch.profiler (CUPTI) and write a per-kernel table to this path")
    p.add_argument("--profile_step", action="store_true",
                   help="run warm-up then ONE step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`) and exit")
    return p.parse_args()


# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def load_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return d["bf16_tflops_sustained"], d["bf16_tflops"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1400.0, 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------------
def cpu_oracle_sample(threads=None, steps=1, warmup=0, tiny=False):
    """Times the oracle (plain PyTorch, fp32, eager) on the host cores on a BOUNDED sample of the workload: SD1.5 at full geometry
    (859.5 M-param UNet, VAE decoder, BLIP-large), B=1, S=2 DDPM steps, K=1 back-propagated step, cfg 7.5, concept-matching loss,
    64x64 latent (512^2 image) = BASELINE.json configs[0] exactly, the reference's own CPU-runnable case (~10 TFLOP, several seconds
    per step).  The sample's algorithmic FLOPs are counted with torch.utils.flop_counter and the result is scaled to config-2
    train-steps by FLOPs (203 TFLOP per config-2 step)."""
    import torch
    from torch.utils.flop_counter import FlopCounterMode
    from oracle import comat_ref as R
    from oracle import sd_modules as sdm
    # the GPU box reports 128 cores; aten/oneDNN scale to ~32 threads there and collapse beyond (profiles/r01_host_cpu_thread_probe.log)
    threads = threads or min(os.cpu_count() or 8, int(os.environ.get("COMAT_CPU_THREADS", "32")))
    torch.set_num_threads(threads)
    torch.manual_seed(42)
    if tiny:
        unet = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(width=64, cross_attention_dim=64))
        vae = sdm.AutoencoderKL(block_out_channels=(64, 64, 128, 128))
        blip = R.make_blip(large=False)
        ctx, res = 64, 128
    else:
        unet, vae, blip = sdm.UNet2DConditionModel(**sdm.SD15_UNET_CONFIG), sdm.AutoencoderKL(), R.make_blip(large=True)
        ctx, res = 768, 512
    unet.requires_grad_(False)
    vae.requires_grad_(False)
    params = sdm.install_lora(unet, 128 if not tiny else 8)
    g = torch.Generator().manual_seed(1)
    lat = res // 8
    batch = dict(prompt_embeds=torch.randn(1, 77, ctx, generator=g), null_embeds=torch.randn(1, 77, ctx, generator=g),
                 latents=torch.randn(1, 4, lat, lat, generator=g), noises=[torch.randn(1, 4, lat, lat, generator=g) for _ in range(2)],
                 training_steps=[1], crop=(0, 0), blip_ids=torch.tensor([[101, 1037, 5855, 1997] + list(range(2000, 2014)) + [102]]),
                 blip_mask=torch.ones(1, 19, dtype=torch.long))
    cfg = dict(S=2, resolution=res)
    opt = torch.optim.AdamW(params, lr=5e-5)
    times, flops = [], None
    first = 0 if warmup >= 1 else -1              # the FLOP-counting pass is the first warm-up step (an extra untimed one if W = 0)
    for i in range(first, warmup + steps):
        t0 = time.perf_counter()
        counter = FlopCounterMode(display=False) if i == first else None
        if counter is not None:
            counter.__enter__()
        out = R.g_step_loss(unet, vae, sdm.DDPMScheduler(), blip, batch, cfg)
        opt.zero_grad()
        out["loss"].backward()
        if counter is not None:
            counter.__exit__(None, None, None)
            flops = float(counter.get_total_flops())
        torch.nn.utils.clip_grad_norm_(params, 0.1)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, threads, flops


SAMPLE_DESC = ("BASELINE configs[0]: SD1.5 full geometry (UNet 859.5 M, VAE decoder, BLIP-large), B=1, S=2, K=1, cfg 7.5, concept-match loss, "
               "64x64 latent (512^2 image), fp32 eager oracle")


def run_reference(a):
    """Reference arm: the reference's CPU path == the oracle restatement (the reference is pure Python on top of
    diffusers, which cannot be installed offline; oracle/pin_against_reference.py pins the restatement against the
    reference's own modules).  Each step is a bounded sample (config-1 geometry) scaled to config-2 units."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    times, threads, flops = cpu_oracle_sample(steps=a.steps, warmup=a.warmup, tiny=a.tiny)
    total = sum(times)
    sample_sps = len(times) / total
    value = sample_sps * (flops / (TFLOP_PER_STEP_CFG2 * 1e12))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SD1.5 512^2 full CoMat, S=20, K=5, B=4 (BASELINE configs[1]); CPU value = bounded sample scaled by "
                                   "algorithmic FLOPs (%.2f TFLOP counted / 203 TFLOP per config-2 step)" % (flops / 1e12),
                       "sample": ("TINY-DEBUG " if a.tiny else "") + SAMPLE_DESC},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{len(times)} sample steps of {total / len(times):.1f} s each ({flops / 1e12:.2f} TFLOP per sample step, "
                                       f"{flops / 1e12 / (total / len(times)):.2f} TFLOP/s on {threads} threads)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)
    return 0


# --------------------------------------------------------------------------------------------------------------
GPU_REF_DESC = ("the oracle restatement of the reference's diffusers / HF path run eagerly on this GPU (SURVEY 8d, BASELINE.md section 4: "
                "diffusers cannot be installed offline): 16-bit weights + fp32 LoRA under autocast (training_utils/pipeline.py:60-65, accelerate "
                "mixed_precision), the attrcon pipeline's hooked explicit attention on the generator UNet (tc_attn_utils.py:104-161: baddbmm + "
                "softmax + bmm, P materialised), SDPA on the discriminator UNet / VAE (diffusers' default processor), gradient checkpointing of "
                "every ResnetBlock2D / Transformer2DModel of a training-mode UNet (scripts/sd15.sh --gradient_checkpointing), cuDNN / cuBLAS, "
                "GradScaler, torch AdamW + clip_grad_norm_, the Python-loop mask loss and the reference's ~7 .item() syncs per step")


def gpu_reference_sample(a, dev, steps, warmup):
    """BASELINE's denominator ("the reference's own 1xB200 diffusers path"): times `steps` full config-2 train steps of the
    reference's algorithm on stock PyTorch kernels.  None of the product's kernels, executors or graphs are on this path; it is
    a reported baseline (like cpu_baseline), never the thing shipped."""
    import random
    import torch
    import torch.nn.functional as F
    from torch.utils.checkpoint import checkpoint
    from comat_b200 import synthetic
    from oracle import comat_ref as R
    from oracle import sd_modules as sdm
    dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    B, S, K = a.batch, a.total_step, a.K
    tiny = a.tiny
    res, ctx_dim = (256, 64) if tiny else (512, 768)
    rank = 8 if tiny else a.rank_lora
    torch.manual_seed(42)

    def make_unet():
        cfg = sdm.tiny_unet_config(width=64, cross_attention_dim=64) if tiny else sdm.SD15_UNET_CONFIG
        with torch.device(dev):
            u = sdm.UNet2DConditionModel(**cfg)
        u.requires_grad_(False)
        u.to(dt)                                                    # pipeline.unet.to(weight_dtype)
        params = sdm.install_lora(u, rank)                          # fp32 LoRA factors (training_utils/pipeline.py:94-115)
        for m in u.modules():                                       # --gradient_checkpointing (diffusers: use_reentrant=False)
            if isinstance(m, (sdm.ResnetBlock2D, sdm.Transformer2DModel)):
                def wrapped(*args, _f=m.forward, _u=u, **kw):
                    if _u.training and torch.is_grad_enabled():
                        return checkpoint(_f, *args, use_reentrant=False, **kw)
                    return _f(*args, **kw)
                m.forward = wrapped
        return u, params

    class Sdpa:                                                     # diffusers AttnProcessor2_0
        def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
            residual, nd = hidden_states, hidden_states.ndim
            if nd == 4:
                b, c, h, w = hidden_states.shape
                hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
            if attn.group_norm is not None:
                hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
            ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
            q, k, v = attn.to_q(hidden_states), attn.to_k(ctx), attn.to_v(ctx)
            n, L, C = q.shape
            sp = lambda x: x.view(n, -1, attn.heads, C // attn.heads).transpose(1, 2)
            o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(n, L, C).to(q.dtype)
            o = attn.to_out[0](o)
            if nd == 4:
                o = o.transpose(-1, -2).reshape(b, c, h, w)
            if attn.residual_connection:
                o = o + residual
            return o / attn.rescale_output_factor

    class Prepared:
        """accelerator.prepare(model) under mixed precision: forward inside autocast, outputs converted to fp32"""
        def __init__(self, m):
            self.m = m

        def __call__(self, *args, **kw):
            with torch.autocast("cuda", dtype=dt):
                out = self.m(*args, **kw)
            return tuple(o.float() for o in out)

        def __getattr__(self, n):
            return getattr(self.m, n)

    unet, g_params = make_unet()
    d_unet, d_params = make_unet()
    with torch.device(dev):
        vae = sdm.AutoencoderKL(block_out_channels=(64, 64, 128, 128)) if tiny else sdm.AutoencoderKL()
    vae.requires_grad_(False).to(dt)
    for u in (d_unet, vae):
        for m in u.modules():
            if m.__class__.__name__ == "Attention":
                m.processor = Sdpa()
    layers = ["up_8", "up_16", "up_32"] if tiny else ["mid_8", "up_16", "up_32", "up_64"]
    ctrl = R.AttentionStore(layers)
    R.register_attention_control(unet, ctrl)                        # training_script.py:318-320: every Attention of the G UNet is hooked
    blip = R.make_blip(large=not tiny, dtype=dt).to(dev)            # from_pretrained(torch_dtype=float16), caption_blip.py:18
    head = torch.nn.Sequential(torch.nn.Linear(4, 1)).to(dev)
    d_params = d_params + list(head.parameters())
    opt_g = torch.optim.AdamW(g_params, lr=5e-5, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    opt_d = torch.optim.AdamW(d_params, lr=2e-5, betas=(0.0, 0.999), weight_decay=1e-2, eps=1e-8)
    scaler = torch.amp.GradScaler("cuda", enabled=dt == torch.float16)
    sched = sdm.DDPMScheduler()
    rng = random.Random(42)
    batches = [synthetic.batch_to_device(synthetic.synthetic_batch(B, 7000 + i, ctx_dim, res, True, True), dev)[0] for i in range(2)]
    pu, pd = Prepared(unet), Prepared(d_unet)
    lat = res // 8

    def step(b):
        unet.train()
        T, A = R.select_training_steps(S, K, rng, 2)                # training_script.py:563-566, :589-590
        z0 = torch.randn(B, 4, lat, lat, device=dev, dtype=dt)
        noises = [torch.randn(B, 4, lat, lat, device=dev, dtype=dt) for _ in range(S)]
        image, z, attn = R.rollout(pu, vae, sched, b["prompt_embeds"].to(dt), b["null_embeds"].to(dt), z0, noises, S, T, 7.5, 0.0, A, ctrl,
                                   return_latents=True)
        off = res // 224
        ox, oy, size = rng.randint(0, off), rng.randint(0, off), res - off
        with torch.autocast("cuda", dtype=dt):                      # caption_blip.py:56-58
            reward = R.blip_score(blip, image[:, :, ox:ox + size, oy:oy + size].to(dt), b["blip"]["input_ids"], b["blip"]["attention_mask"], 4)
        loss = -reward.float().mean()
        d_unet.eval()                                               # gan_sdxl.py:55
        g_loss = R.d_forward(pd, head, sched, z, b["gan_null_embeds"].to(dt), S, "G")
        loss = loss + g_loss
        tok, pix = R.mask_loss(attn, b["words"], b["masks"], layers, image.detach().float())
        loss = loss + 1e-3 * tok + 5e-5 * pix
        image.register_hook(lambda g: (g.norm(2).item(), g)[1])    # record_grad: a host sync inside backward (:644-651)
        loss.item()                                                 # avg_loss gather (:653-654)
        opt_g.zero_grad()
        scaler.scale(loss).backward()
        scaler.unscale_(opt_g)
        torch.nn.utils.clip_grad_norm_(g_params, 0.1)
        scaler.step(opt_g)
        logs = [loss.item(), reward.item(), g_loss.item(), tok.item(), pix.item()]          # :667-676
        d_unet.train()                                              # gan_sdxl.py:94
        d_loss = R.d_forward(pd, head, sched, z.detach(), b["gan_null_embeds"].to(dt), S, "D", b["real_latents"].to(dt))
        logs.append(d_loss.item())
        opt_d.zero_grad()
        scaler.scale(d_loss).backward()
        scaler.unscale_(opt_d)
        torch.nn.utils.clip_grad_norm_(d_params, 1.0)
        scaler.step(opt_d)
        scaler.update()
        return logs

    for i in range(warmup):
        step(batches[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(steps):
        logs = step(batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    peak = torch.cuda.max_memory_allocated() / 2**30
    del unet, d_unet, vae, blip, opt_g, opt_d, pu, pd, ctrl
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return {"value": steps / t, "unit": UNIT, "ms_per_step": 1e3 * t / steps, "steps": steps, "warmup": warmup, "dtype": a.dtype,
            "config": ("TINY-DEBUG " if tiny else "") + "same workload as the product line (SD1.5 512^2 full CoMat, S=%d, K=%d, batch %d, LoRA r=%d)" % (S, K, B, rank),
            "what": GPU_REF_DESC, "losses_last_step": logs, "peak_allocated_gb": peak}


def run_gpu_reference(a):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    clocks = ClockSampler(dev.index)
    clocks.start()
    r = gpu_reference_sample(a, dev, a.steps, a.warmup)
    line = {"impl": "gpu_reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype,
            "data": "synthetic", "config": {"workload": r["config"], "what": r["what"]}, "clocks": clocks.stop(),
            "gpu_launches": 0, "peak_allocated_gb": r["peak_allocated_gb"]}
    _emit(line)
    return 0


# --------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _claim_stdout():
    """the driver parses ONE JSON line from stdout: route everything libraries print to fd 1 during the run (NCCL's version
    banner, warnings) to stderr and keep the original stdout for that line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


CONFIGS = {2: dict(batch=4, metric=METRIC, tflop=TFLOP_PER_STEP_CFG2, name="SD1.5 512^2 full CoMat", tag="BASELINE configs[1]"),
           3: dict(batch=8, metric=METRIC, tflop=TFLOP_PER_STEP_CFG2 * 2, name="SD1.5 512^2 full CoMat", tag="BASELINE configs[2] (masks precomputed; GAN on)"),
           4: dict(batch=1, metric="sdxl_512_comat_train_steps_per_sec", tflop=90.0, name="SDXL 512^2 full CoMat, SD1.5 discriminator",
                   tag="BASELINE configs[3]"),
           5: dict(batch=2, metric="sdxl_unet_fwd_bwd_iters_per_sec", tflop=None, name="SDXL UNet forward + backward", tag="BASELINE configs[4]")}


def run_config5(a):
    """BASELINE configs[4]: SDXL UNet-only denoise forward + backward (data gradient + all 1 120 LoRA weight gradients) on the explicit
    executors, per-GPU batch --batch at a --latent^2 latent.  One "step" = one fwd+bwd; value = iterations/s summed over ranks."""
    import torch
    import torch.distributed as dist
    from comat_b200 import _lib, ops, synthetic
    from comat_b200.modules import EngineUNet
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))      # torchrun pins OMP_NUM_THREADS=1: model construction crawls
    dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    lat, n = (32 if a.tiny else a.latent), a.batch
    unet_p = synthetic.build_sdxl_unet(dev, rank=8 if a.tiny else a.rank_lora, seed=42, tiny=a.tiny)
    mod = EngineUNet(unet_p, dt)
    for p_ in mod.lora_parameters():
        p_.grad = torch.zeros_like(p_)
    mod.direct_lora_grads = True
    # the taped call replayed from a (forward, backward) CUDA-graph pair: at batch 1-2 the eager executor is host-bound
    # (~6 000 launches per iteration, profiles/r02_config5_sweep.jsonl: 118-130 ms whatever the shape)
    mod.graph_taped = (not a.no_graphs) and a.graph_taped != "off"
    g = torch.Generator(device="cuda").manual_seed(rank)
    x = torch.randn(n, 4, lat, lat, device=dev, generator=g)
    ctx = torch.randn(n, 77, 64 if a.tiny else 2048, device=dev, generator=g)
    added = dict(text_embeds=torch.randn(n, 16 if a.tiny else 1280, device=dev, generator=g),
                 time_ids=torch.tensor([[8.0 * lat, 8.0 * lat, 0, 0, 8.0 * lat, 8.0 * lat]] * n, device=dev))
    t = torch.tensor(500, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def it():
        xr = x.clone().requires_grad_(True)
        eps = mod(xr, t, encoder_hidden_states=ctx, added_cond_kwargs=added)[0]
        eps.float().pow(2).mean().backward()
        mod.finalize_lora_grads()
        return eps
    for _ in range(max(a.warmup, 3)):          # eager pass (marks the signature warm), capture pass, first replay
        it()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    l0 = _lib.LAUNCH_COUNT
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(a.steps):
        eps = it()
    e1.record()
    torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev = float(tt)
    launches = _lib.LAUNCH_COUNT - l0
    F_fwd = {32: None, 64: 1.589, 96: 1.589 * 2.25 * 1.05, 128: 6.76}.get(lat)          # TFLOP per sample forward (96: interpolated)
    sustained, burst, hbm, peak_src = load_peaks()
    if rank == 0:
        value = world * a.steps / t_dev
        tf = None if F_fwd is None else 3.0 * n * F_fwd * a.steps / t_dev              # fwd + dgrad + LoRA wgrad ~ 3 x forward FLOPs... dgrad only: 2x
        line = {"metric": CONFIGS[5]["metric"], "value": value, "unit": "iters/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype,
                "data": "synthetic",
                "config": {"workload": ("TINY-DEBUG " if a.tiny else "") + "SDXL UNet (2.567 B) forward + backward incl. LoRA r=%d weight gradients, latent %dx%d, "
                           "batch %d/GPU (BASELINE configs[4])" % (a.rank_lora, lat, lat, n), "parallelism": f"dp{world}", "global_batch": n * world, "cuda_graphs_taped_calls": bool(mod.graph_taped),
                           "l2_policy": "activations of one pass (> 10 GB) exceed the 126 MB L2"},
                "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "microbench: inputs are device tensors by definition of this config"},
                "gpu_launches": launches, "clocks": clocks.stop() if clocks else None,
                "roofline": {"bound": "tensor", "kernel": "whole UNet fwd+bwd (tcgen05 GEMM / conv / attention)", "achieved": tf, "peak": sustained,
                             "unit": "TFLOP/s", "frac": None if tf is None else tf / sustained, "traffic": None, "peak_source": peak_src + ", sustained",
                             "flops_definition": "2 x forward FLOPs (forward + data gradient; LoRA weight gradients and recompute not counted)"},
                "eps_abs_mean": float(eps.abs().mean())}
        if tf is not None:
            line["roofline"]["achieved"] = 2.0 * n * F_fwd * a.steps / t_dev
            line["roofline"]["frac"] = line["roofline"]["achieved"] / sustained
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse()
    if a.batch == 0:
        a.batch = CONFIGS[a.config]["batch"]
    _claim_stdout()
    # watchdog: dump every thread's Python stack to stderr and exit if the run is still going after N seconds.  On by default (30 min)
    # for multi-rank product runs, where a dead-lock between ranks would otherwise sit until the launcher's own limit
    # (COMAT_BENCH_WATCHDOG=<seconds> sets it for any run, 0 disables).
    wd = os.environ.get("COMAT_BENCH_WATCHDOG")
    if wd is None and a.impl == "comat_b200" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        wd = "1800"
    if wd and int(wd) > 0:
        import faulthandler
        faulthandler.dump_traceback_later(int(wd), exit=True)
    if os.environ.get("COMAT_HOST_ONLY_TIMING"):
        raise SystemExit("bench.py: COMAT_HOST_ONLY_TIMING is set - refusing to emit a bench line (tools/host_issue_time.py is the host-only probe)")
    if a.impl == "reference":
        return run_reference(a)
    if a.impl == "gpu_reference":
        return run_gpu_reference(a)
    if a.config == 5:
        return run_config5(a)
    import torch
    import torch.distributed as dist
    from comat_b200 import _lib, attention, caption, image_ops, ops, synthetic
    from comat_b200.caption import Blip, CaptionModelWrapper
    from comat_b200.gan import D_sd
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import AttentionStore, AttrConcenTrainableSDPipeline, register_attention_control
    from comat_b200.trainer import CoMatTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (product arm) needs a GPU: comat_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))      # torchrun pins OMP_NUM_THREADS=1: model construction crawls
    _lib.lib()
    dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    gan, attrcon = not a.no_gan, not a.no_attrcon
    gpu_ref = None
    if rank == 0 and not a.no_gpu_reference and a.config == 2 and not (a.kineto_step or a.profile_step or a.sync_debug):
        # the denominator of BASELINE's ">= 10x the reference's own 1xB200 path" target, measured on this same box BEFORE the product
        # is built (its 85 GB peak is released again; the product then holds up to ~150 GB of weights, CUDA graphs and activations)
        try:
            gpu_ref = gpu_reference_sample(a, dev, steps=2, warmup=1)
        except Exception as e:  # a baseline must never take the bench line down
            gpu_ref = {"value": None, "unit": UNIT, "error": repr(e)[:300]}
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    if world > 1:
        dist.barrier()

    sdxl = a.config == 4
    base = ("sdxl" if sdxl else "sd_1_5") + ("_attrcon" if attrcon else "")
    args = synthetic.default_args(pretrain_model_name=base, train_batch_size=a.batch,
                                  gradient_accumulation_steps=1, learning_rate=5e-5, learning_rate_D=2e-5, max_grad_norm=0.1,
                                  max_grad_norm_D=1.0, adam_beta1_D=0.0, lora_rank=a.rank_lora, K=a.K, total_step=a.total_step,
                                  gan_loss=gan, gan_model_arch="gansd_1_5", gan_loss_weight=1.0, attrcon_train_steps=2, seed=42,
                                  resolution=256 if a.tiny else 512)
    rank_lora = 8 if a.tiny else a.rank_lora
    if sdxl:
        from comat_b200 import containers as Cn
        from comat_b200.pipelines import AttrConcenTrainableSDXLPipeline
        unet_p = synthetic.build_sdxl_unet(dev, rank=rank_lora, seed=42, tiny=a.tiny)
        torch.manual_seed(48)
        with torch.device(dev):
            vae_p = Cn.AutoencoderKL(**(dict(block_out_channels=(64, 64, 128, 128)) if a.tiny else {}), scaling_factor=0.13025)
        vae_p.requires_grad_(False)
        pipe = AttrConcenTrainableSDXLPipeline(EngineVAE(vae_p, dt), EngineUNet(unet_p, dt))
        layers = ["up_8", "up_16"] if a.tiny else ["mid_16", "up_16", "up_32"]              # training_script.py:312
    else:
        unet_p, vae_p = synthetic.build_sd15(dev, dt, rank=rank_lora, seed=42, tiny=a.tiny)
        pipe = AttrConcenTrainableSDPipeline(EngineVAE(vae_p, dt), EngineUNet(unet_p, dt))
        layers = ["up_8", "up_16", "up_32"] if a.tiny else ["mid_8", "up_16", "up_32", "up_64"]
    if attrcon:
        args.train_layer_ls = layers
        register_attention_control(pipe, AttentionStore(layers))
    from comat_b200.blip_engine import BlipEngine
    blip = Blip(BlipEngine(synthetic.build_blip(dev, dt, large=not a.tiny), dt))
    cap = CaptionModelWrapper(["Blip"], [1.0], blip)
    D = None
    if gan:
        d_unet, _ = synthetic.build_sd15(dev, dt, rank=rank_lora, seed=43, tiny=a.tiny)     # SD1.5 discriminator, also under SDXL (scripts/sdxl.sh:15)
        D = D_sd(EngineUNet(d_unet, dt))
    pipe.unet.use_graphs = not a.no_graphs
    taped = (not a.no_graphs) and (a.graph_taped == "on" or (a.graph_taped == "auto" and a.config in (2, 4)))
    pipe.unet.graph_taped = taped
    if D is not None:
        D.unet.graph_taped = taped
    trainer = CoMatTrainer(args, pipe, cap, D, process_group=None,
                           manual_gc_interval=0 if os.environ.get("COMAT_MANUAL_GC") == "0" else 25)
    if world > 1 and a.config == 4:
        # configs[3] at 2 ranks never finished its first step with the optimiser tails on the side stream (open):
        # run them on the main stream there (configs[1] at N = 2 is verified with the side-stream tails and keeps them)
        trainer.overlap_updates = False
    ctx_dim = 64 if a.tiny else (2048 if sdxl else 768)
    sd15_ctx = 64 if a.tiny else 768
    pooled = 0 if not sdxl else (64 - 6 * 8 if a.tiny else 1280)
    host_batches = [synthetic.synthetic_batch(a.batch, 1000 * rank + i, ctx_dim, args.resolution, attrcon, gan, pinned=True,
                                              pooled_dim=pooled, gan_ctx_dim=sd15_ctx)
                    for i in range(4)]
    dev_batches = [synthetic.batch_to_device(b, dev)[0] for b in host_batches]
    h2d_bytes = synthetic.batch_to_device(host_batches[0], dev)[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    loss_pinned = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    host_losses = []

    def timed(n_steps, host_inputs):
        barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        d2h = 0
        th0 = time.perf_counter()
        e0.record()
        pending = None
        for i in range(n_steps):
            if host_inputs:
                b, _ = synthetic.batch_to_device(host_batches[i % len(host_batches)], dev)
            else:
                b = dev_batches[i % len(dev_batches)]
            logs = trainer.train_step(b)
            if host_inputs:
                # D2H read of EVERY step's result, pipelined the way a training loop logs: the copy into pinned memory is queued
                # behind the step, the host reads the value of step i-1 while step i is being enqueued (a .cpu() here would stall
                # the host until the whole step has run and leave the GPU idle while the next step is issued)
                slot = loss_pinned[i % 2]
                slot.copy_(logs["step_loss"].detach().float().reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if pending is not None:
                    pending[0].synchronize()
                    host_losses.append(float(pending[1][0]))
                pending = (ev, slot)
                d2h = 4
        if pending is not None:
            pending[0].synchronize()
            host_losses.append(float(pending[1][0]))
        trainer.sync()                                            # side-stream optimiser tails of the last step join the timed stream
        e1.record()
        timed.host_issue_s = (time.perf_counter() - th0) / n_steps    # time the host needed to enqueue a step (no sync inside)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t), d2h, logs

    # set-up, not measurement: two priming steps build the CUDA graphs of the no-grad forwards, the folded-weight buffers and
    # grow the caching allocator to its steady-state footprint (a cudaMalloc inside a timed step stalls the queue for tens of
    # ms: profiles/r01_bench_v8.json shows one such step); the W warm-up steps the contract asks for follow.
    for i in range(2):
        trainer.train_step(dev_batches[(i + 1) % len(dev_batches)])
    for i in range(a.warmup):
        trainer.train_step(dev_batches[i % len(dev_batches)])
    if a.sync_debug:
        import traceback, warnings
        torch.cuda.synchronize()
        seen = {}
        def _show(message, category, filename, lineno, file=None, line=None):
            st = [f for f in traceback.extract_stack() if "/comat_b200/" in f.filename or f.filename.endswith("bench.py")]
            key = " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in reversed(st[-4:]))
            seen[key] = seen.get(key, 0) + 1
        warnings.showwarning = _show
        warnings.simplefilter("always")
        torch.cuda.set_sync_debug_mode("warn")
        trainer.train_step(dev_batches[0])
        torch.cuda.set_sync_debug_mode("default")
        torch.cuda.synchronize()
        for k, v in sorted(seen.items(), key=lambda kv: -kv[1]):
            print(f"SYNC x{v}: {k}")
        return 0
    if a.kineto_step:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        if a.gemm_shapes:
            ops.PROFILE = gp = {"flops": 0.0, "events": [], "keys_only": True}
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(a.kineto_steps):
                trainer.train_step(dev_batches[i % len(dev_batches)])
            torch.cuda.synchronize()
        ops.PROFILE = None
        if a.gemm_shapes:
            # kernel durations (CUPTI) of the GEMM launches in issue order <-> the shapes ops.gemm recorded in the same order
            # (only valid with --no_graphs: graph replays do not pass through ops.gemm)
            kev = sorted((e for e in prof.events() if "gemm_tc_" in e.name), key=lambda e: e.time_range.start)
            by = {}
            if len(kev) == len(gp["events"]):
                for e, (_, _, key, fl) in zip(kev, gp["events"]):
                    r = by.setdefault(key, [0, 0.0, 0.0])
                    r[0] += 1
                    r[1] += e.device_time_total * 1e-3
                    r[2] += fl
            with open(a.gemm_shapes, "w") as f:
                f.write(f"GEMM launches of one train step by shape, CUPTI kernel durations ({len(kev)} kernels, {len(gp['events'])} calls)\n")
                f.write("| M | N | K segs | taps | split_k | fp32 out | calls | total ms | us/call | TFLOP/s |\n|---|---|---|---|---|---|---:|---:|---:|---:|\n")
                for key, r in sorted(by.items(), key=lambda kv: -kv[1][1]):
                    f.write("| %d | %d | %s | %d | %d | %s | %d | %.2f | %.1f | %.0f |\n" % (
                        key[0], key[1], "+".join(map(str, key[2])), key[3], key[4], key[5], r[0], r[1], 1e3 * r[1] / r[0],
                        r[2] / max(r[1], 1e-9) / 1e9))
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in rows)
        with open(a.kineto_step, "w") as fh:
            fh.write(f"one train step, torch.profiler (CUPTI) device time per kernel; total {tot / 1e3:.1f} ms\n")
            fh.write("| kernel | calls | total ms | share |\n|---|---:|---:|---:|\n")
            for e in rows[:45]:
                fh.write(f"| `{e.key[:100]}` | {e.count} | {e.device_time_total / 1e3:.2f} | {100 * e.device_time_total / tot:.1f} % |\n")
            # idle time between consecutive kernels on the device timeline: ~1-2 us gaps are back-to-back launches, long
            # gaps mean the GPU waited for the host
            try:
                from torch.autograd import DeviceType
                dev_ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                                 if e.device_type == DeviceType.CUDA and e.time_range.end > e.time_range.start), key=lambda t: t[0])
                edges = [2, 5, 20, 100, 1000, 1e9]
                cnt, tot_gap = [0] * len(edges), [0.0] * len(edges)
                end, prev_name, big = dev_ev[0][1], dev_ev[0][2], []
                for st, en, nm in dev_ev[1:]:
                    gap = st - end
                    if gap > 0:
                        k = next(i for i, e_ in enumerate(edges) if gap < e_)
                        cnt[k] += 1
                        tot_gap[k] += gap
                        if gap > 200:
                            big.append((gap, (end - dev_ev[0][0]) / 1e3, prev_name[:60], nm[:60]))
                    if en >= end:
                        end, prev_name = en, nm
                span = (dev_ev[-1][1] - dev_ev[0][0]) / 1e3
                fh.write(f"\ndevice timeline: {len(dev_ev)} kernels over {span:.1f} ms; idle gaps between consecutive kernels:\n")
                fh.write("| gap (us) | count | total ms |\n|---|---:|---:|\n")
                lo = 0
                for e_, c_, t_ in zip(edges, cnt, tot_gap):
                    fh.write(f"| {lo}-{e_ if e_ < 1e9 else 'inf'} | {c_} | {t_ / 1e3:.2f} |\n")
                    lo = e_
                fh.write("\nlargest idle gaps (us, at ms from the first kernel, kernel before -> kernel after):\n")
                for g_, at_, pn_, nn_ in sorted(big, reverse=True)[:16]:
                    fh.write(f"- {g_:.0f} us at {at_:.1f} ms: `{pn_}` -> `{nn_}`\n")
            except Exception as ex:  # profiler internals differ between torch versions
                fh.write(f"\n(gap analysis unavailable: {ex!r})\n")
        return 0
    if a.profile_step:
        trainer.sync()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        trainer.train_step(dev_batches[0])
        trainer.sync()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0
    clocks = ClockSampler(local) if rank == 0 else None
    ms0 = torch.cuda.memory_stats()
    if clocks:
        clocks.start()
    l0 = _lib.LAUNCH_COUNT
    lib0 = attention.LIBRARY_CALLS + caption.LIBRARY_CALLS + image_ops.LIBRARY_CALLS
    t_dev, _, logs = timed(a.steps, host_inputs=False)
    host_issue_ms = 1e3 * timed.host_issue_s
    ms1 = torch.cuda.memory_stats()
    mem_note = {"peak_allocated_gb": ms1.get("allocated_bytes.all.peak", 0) / 2**30, "peak_reserved_gb": ms1.get("reserved_bytes.all.peak", 0) / 2**30,
                "cudaMalloc_calls_in_timed_steps": ms1.get("num_device_alloc", 0) - ms0.get("num_device_alloc", 0),
                "alloc_retries_in_timed_steps": ms1.get("num_alloc_retries", 0) - ms0.get("num_alloc_retries", 0)}
    launches = _lib.LAUNCH_COUNT - l0
    lib_calls = attention.LIBRARY_CALLS + caption.LIBRARY_CALLS + image_ops.LIBRARY_CALLS - lib0
    # one untimed step through the host-input path first: its H2D staging tensors come out of the caching allocator for the first time
    # (a cudaMalloc inside a 3-step timed region cost ~20 ms per step in profiles/r02_bench_v15_default.json: e2e 2.11 vs value 2.22)
    n_host_seen = len(host_losses)
    timed(1, host_inputs=True)
    del host_losses[n_host_seen:]
    t_e2e, d2h, _ = timed(a.steps, host_inputs=True)
    clk = clocks.stop() if clocks else None

    # ---- roofline of the dominant kernel family: one instrumented step (per-launch CUDA events)
    # (eager launches for this one step: kernels replayed from a CUDA graph do not pass through ops.gemm's event pair)
    graphs_were, taped_were = pipe.unet.use_graphs, pipe.unet.graph_taped
    pipe.unet.use_graphs = pipe.unet.graph_taped = False
    if D is not None:
        D.unet.graph_taped = False
    ops.PROFILE = prof = {"flops": 0.0, "events": []}
    trainer.train_step(dev_batches[0])
    torch.cuda.synchronize()
    ops.PROFILE = None
    pipe.unet.use_graphs, pipe.unet.graph_taped = graphs_were, taped_were
    gemm_s = sum(ev[0].elapsed_time(ev[1]) for ev in prof["events"]) * 1e-3
    sustained, burst, hbm, peak_src = load_peaks()
    step_s = t_dev / a.steps
    roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 GEMM / implicit-GEMM conv)",
            "achieved": prof["flops"] / max(gemm_s, 1e-9) / 1e12, "peak": sustained, "unit": "TFLOP/s",
            "frac": prof["flops"] / max(gemm_s, 1e-9) / 1e12 / sustained, "peak_source": peak_src + ", sustained",
            "launches_per_step": len(prof["events"]), "kernel_seconds_per_step": gemm_s,
            "share_of_step": gemm_s / step_s, "algorithmic_tflop_per_step": prof["flops"] / 1e12,
            "timing": "one CUDA-event pair per launch on the launching stream, eager instrumented step after the timed region",
            "note": "the family's launches also carry work that used to be separate passes: bias / time-embedding / residual adds, "
                    "torch.cat as K segments, fused GEGLU (no-grad passes) and, since r02, the GroupNorm statistics of their outputs "
                    "(routed for 49 of 61 GroupNorms per UNet call; 59 % of a step's GroupNorm forwards) - their time counts against the GEMM FLOPs here.  An event pair around every "
                    "launch also measures the ~5-7 us record gap: the same launches' CUPTI kernel durations (bench.py --kineto_step / "
                    "--gemm_shapes, profiles/r02_gemm_shapes_v16.md) sum to 0.238 s per step = 756 TFLOP/s = 0.55 of the peak",
            # DRAM bytes of ONE launch of the family's largest in-step shape from the committed ncu --set full capture
            "traffic": 22.86e6, "traffic_note": "dram__bytes_read + write of one conv3x3 320->320 launch at 64x64, n=8 (gemm_tc_kernel<160,3,LEAN>): 22.85 MB "
                       "read + 0.006 MB written back at capture time; algorithmic bytes 21.0 MB activations in + 1.8 MB weights + 21.0 MB out (the "
                       "output is still L2-resident when the kernel ends); profiles/r02_ncu_gemm_final.md"}

    if rank == 0:
        # whole-job aggregate (weak scaling): every rank runs K optimiser steps on its own batch of `--batch` prompts, so the job
        # processes world x K per-GPU-batch steps in t_dev; the data-parallel optimiser advances value / world global steps/s
        value = world * a.steps / t_dev
        cfgd = CONFIGS[a.config]
        line = {"metric": cfgd["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * t_dev / a.steps, "host_issue_ms_per_step": host_issue_ms, "memory": mem_note, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": a.dtype, "data": "synthetic",
                "config": {"workload": ("TINY-DEBUG " if a.tiny else "") + cfgd["name"] + " (BLIP concept-match + attention-map "
                           "token/pixel loss on 2 attrcon steps + GAN G/D), S=%d DDPM steps, K=%d, batch %d/GPU, LoRA r=%d, cfg 7.5 "
                           "(%s)" % (a.total_step, a.K, a.batch, rank_lora, cfgd["tag"]),
                           "bench_config": a.config, "algorithmic_tflop_per_step_nominal": cfgd["tflop"],
                           "global_batch": a.batch * world, "parallelism": f"dp{world}", "cuda_graphs": not a.no_graphs,
                           "cuda_graphs_taped_calls": bool(taped),
                           "l2_policy": "inputs rotate over 4 batches; per-step working set (weights 7 GB + activations > 50 GB) exceeds the 126 MB L2",
                           "samples_per_sec": value * a.batch, "global_optimizer_steps_per_sec": value / world,
                           "value_definition": "per-GPU-batch train-steps per second summed over ranks (= n_gpus x global optimiser steps/s)",
                           "library_calls_per_step": lib_calls / a.steps,
                           # second half of BASELINE's metric ("UNet attn tensor-pipe %"): not measurable without a profiler, so the
                           # committed ncu --set full capture is cited, never a number taken in this (unprofiled) run
                           "unet_attn_tensor_pipe_pct": {"attn_fwd_long_kernel<40>": 26.2, "attn_bwd_kernel<40,dQ>": 23.2,
                                                         "attn_bwd_kernel<40,dKdV>": 21.1, "attn_fwd_kernel<40> cross-attention + P export (77 keys)": 5.3,
                                                         "shape": "SD1.5 64^2-latent self-attention, n=8, 8 heads, d=40",
                                                         "ceiling": "37 % at d = 40 while every exponential goes through MUFU (512 vs 192 clk per tile; profiles/r02_attn_fwd_analysis.md)",
                                                         "source": "profiles/r02_ncu_v3.md (sm__pipe_tensor cycles active, ncu --set full --clock-control none)"},
                           "library_note": ("0 = every conv / linear / attention (fwd+bwd) / norm / loss / resize / optimiser launch of the step "
                                            "is a comat_b200 kernel; torch supplies memory, the fp32 latent-chain glue, "
                                            "embedding gathers and layout permutes") if lib_calls == 0 else
                                           "calls that fell back to aten/HF kernels (shapes the native kernels do not cover)"},
                "e2e": {"value": world * a.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h,
                        "losses_read_on_host": host_losses[-a.steps:],
                        "note": "pinned-host batch copied in and the step loss copied out every step inside the timed region; the loss of "
                                "step i-1 is read on the host while step i is enqueued"},
                "gpu_launches": launches, "clocks": clk, "roofline": roof,
                "losses": {k: float(v.detach()) for k, v in logs.items() if hasattr(v, "numel") and v.numel() == 1}}
        if gpu_ref is not None:
            if gpu_ref.get("value"):
                gpu_ref["product_over_gpu_reference"] = {"n_gpus": world, "value_ratio": value / gpu_ref["value"],
                                                         "e2e_ratio": line["e2e"]["value"] / gpu_ref["value"]}
            line["gpu_reference"] = gpu_ref
        if not a.no_cpu_baseline and a.config == 2:
            try:
                times, threads, flops = cpu_oracle_sample(steps=2, warmup=0, tiny=a.tiny)
                v = (len(times) / sum(times)) * (flops / (TFLOP_PER_STEP_CFG2 * 1e12))
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                        "sample": f"{len(times)} steps of [{SAMPLE_DESC}] = {sum(times):.1f} s, {flops / 1e12:.2f} TFLOP counted per step "
                                                  f"({flops / 1e12 * len(times) / sum(times):.2f} TFLOP/s), scaled by FLOPs to config-2 steps (203 TFLOP)"}
            except Exception as e:  # the baseline must never take the bench line down
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
