"""Measure the SURVEY 8f rows on one B200 at full geometry (CUDA events on the launching stream, warm-up first):

  * f-1  ``encode_prompt``: SD1.5 (CLIP-L) and SDXL (CLIP-L + OpenCLIP-bigG, 32 layers) on B prompt strings, with the CPU
         timing of the HF fp32 modules next to it (the reference's path for this row, host cores);
  * f-2  GAN ground-truth producer: full SD1.5 UNet, 50 DDPM steps, cfg 7.5, batch 8, CUDA-graphed forwards -> samples/s
         and TFLOP/s (2 * 0.803 TFLOP per sample-step);
  * f-3  checkpoint save / load of the r=128 LoRA set (25.5 M fp32 parameters + optimiser state).

Prints one JSON object.  ``python tools/bench_8f.py > gpurun_out/bench_8f.json``"""
import json
import os
import random
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ev_time(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from comat_b200 import _lib, checkpoint as CK, gan_data as GD, synthetic
    from comat_b200.gan import D_sd
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import TrainableSDPipeline, TrainableSDXLPipeline
    from comat_b200.text_encoder import EngineCLIPText
    from comat_b200.trainer import CoMatTrainer
    dev = torch.device("cuda")
    dt = torch.float16
    out = {"device": torch.cuda.get_device_name(0)}
    B = 4
    prompts = ["a red apple on a wooden table next to a blue cup", "two dogs playing on a green sofa", "a yellow bus under a bridge",
               "three white birds above a dark lake at night"][:B]

    # ---- f-1
    clip_l = synthetic.build_clip_text(dev, torch.float32, seed=7, which="clip_l")
    bigg = synthetic.build_clip_text(dev, torch.float32, seed=8, which="bigg")
    e1, e2 = EngineCLIPText(clip_l, dt), EngineCLIPText(bigg, dt)
    sd = TrainableSDPipeline.__new__(TrainableSDPipeline)
    TrainableSDPipeline.__init__(sd, vae=None, unet=None, text_encoder=e1, tokenizer=synthetic.SyntheticClipTokenizer())
    xl = TrainableSDXLPipeline.__new__(TrainableSDXLPipeline)
    TrainableSDXLPipeline.__init__(xl, vae=None, unet=None, text_encoder=e1, tokenizer=synthetic.SyntheticClipTokenizer(), text_encoder_2=e2,
                                   tokenizer_2=synthetic.SyntheticClipTokenizer(pad_token_id=0), force_zeros_for_empty_prompt=True)
    e1.use_graphs = e2.use_graphs = False
    l0 = _lib.LAUNCH_COUNT
    sd.encode_prompt(prompts, dev, 1, False)
    launches_sd = _lib.LAUNCH_COUNT - l0
    ms_sd_eager = ev_time(lambda: sd.encode_prompt(prompts, dev, 1, False), 20)
    ms_xl_eager = ev_time(lambda: xl.encode_prompt(prompts, device=dev, num_images_per_prompt=1, do_classifier_free_guidance=True), 10)
    e1.use_graphs = e2.use_graphs = True
    ms_sd = ev_time(lambda: sd.encode_prompt(prompts, dev, 1, False), 20)
    ms_xl = ev_time(lambda: xl.encode_prompt(prompts, device=dev, num_images_per_prompt=1, do_classifier_free_guidance=True), 10)
    ids = synthetic.SyntheticClipTokenizer()(prompts).input_ids.to(dev)
    ms_l_only = ev_time(lambda: e1(ids), 20)
    ms_g_only = ev_time(lambda: e2(ids, output_hidden_states=True), 10)
    # algorithmic FLOPs of one encoder pass: per layer 2*T*(4 C^2 + 2 C F) + 4 T^2 C (causal counted dense)
    def enc_flops(C, F, L, T=77):
        return L * (2 * T * (4 * C * C + 2 * C * F) + 4 * T * T * C)
    fl_l, fl_g = B * enc_flops(768, 3072, 12), B * enc_flops(1280, 5120, 32)
    # reference path for this row on the host cores: HF fp32 modules (what encode_prompt calls), same prompts
    cl_cpu, bg_cpu = clip_l.cpu(), bigg.cpu()
    ids_cpu = ids.cpu()
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    with torch.no_grad():
        cl_cpu(ids_cpu); t0 = time.perf_counter(); [cl_cpu(ids_cpu) for _ in range(3)]; cpu_l = (time.perf_counter() - t0) / 3 * 1e3
        bg_cpu(ids_cpu, output_hidden_states=True); t0 = time.perf_counter(); bg_cpu(ids_cpu, output_hidden_states=True); cpu_g = (time.perf_counter() - t0) * 1e3
    out["f1_encode_prompt"] = {
        "batch": B, "sd15_ms": ms_sd, "sd15_ms_eager": ms_sd_eager, "sdxl_ms_eager": ms_xl_eager, "sd15_launches": launches_sd, "sdxl_ms_both_encoders_cfg_zeroed_negative": ms_xl,
        "clip_l_forward_ms": ms_l_only, "bigg_forward_ms": ms_g_only,
        "clip_l_tflops": fl_l / (ms_l_only * 1e-3) / 1e12, "bigg_tflops": fl_g / (ms_g_only * 1e-3) / 1e12,
        "cpu_hf_fp32_clip_l_ms": cpu_l, "cpu_hf_fp32_bigg_ms": cpu_g, "cpu_threads": torch.get_num_threads(),
        "note": "sd15_ms / sdxl_ms / *_forward_ms: CUDA-graphed encoder forwards (default); *_eager: launch by launch"}
    del cl_cpu, bg_cpu, bigg, e2, xl
    torch.cuda.empty_cache()

    # ---- f-2
    unet_p, vae_p = synthetic.build_sd15(dev, dt, rank=128, seed=42)
    pipe = TrainableSDPipeline(EngineVAE(vae_p, dt), EngineUNet(unet_p, dt), text_encoder=EngineCLIPText(clip_l.to(dev), dt),
                               tokenizer=synthetic.SyntheticClipTokenizer())
    bs, S = 8, 50
    gt_prompts = ["prompt number %d with a few more words" % i for i in range(3 * bs)]
    with tempfile.TemporaryDirectory() as td:
        index = os.path.join(td, "train_data", "gan_train_data.jsonl")
        gen = torch.Generator(device="cuda").manual_seed(1)
        GD.generate_gan_ground_truth(pipe, gt_prompts[:bs], index, batch_size=bs, num_inference_steps=S, generator=gen)   # graph capture + warm-up
        torch.cuda.synchronize()
        l0 = _lib.LAUNCH_COUNT
        t0 = time.perf_counter()
        n = GD.generate_gan_ground_truth(pipe, gt_prompts[bs:], index, batch_size=bs, num_inference_steps=S, generator=gen)
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        files = len(os.listdir(os.path.join(td, "train_data", "latents")))
    out["f2_gan_gt_producer"] = {
        "model": "SD1.5 UNet 859.5 M, fp16, cuda graphs", "batch": bs, "steps": S, "cfg": 7.5, "samples": n, "seconds_wall_incl_file_io": dt_s,
        "samples_per_sec": n / dt_s, "tflops": n * S * 2 * 0.803 / dt_s, "kernel_launches": _lib.LAUNCH_COUNT - l0, "files_written": files}

    # ---- f-3
    d_p, _ = synthetic.build_sd15(dev, dt, rank=128, seed=43)
    args = synthetic.default_args(pretrain_model_name="sd_1_5", gan_loss=True, seed=1)
    tr = CoMatTrainer(args, pipe, None, D_sd(EngineUNet(d_p, dt)), rng=random.Random(1))
    with tempfile.TemporaryDirectory() as td:
        t0 = time.perf_counter(); path = CK.save_checkpoint(tr, td); t_save = time.perf_counter() - t0
        sizes = {os.path.relpath(os.path.join(r, f), path): os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(path) for f in fs}
        t0 = time.perf_counter(); step = CK.load_checkpoint(tr, td, "latest"); torch.cuda.synchronize(); t_load = time.perf_counter() - t0
    out["f3_checkpoint"] = {"save_s": t_save, "load_s": t_load, "files_bytes": sizes, "global_step": step,
                            "lora_params_G": tr.optimizer.n, "params_D_incl_head": tr.D_optimizer.n}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
