"""Pin the oracle against the reference's OWN code and write the golden fixtures.

Run in the build container (needs /root/reference):   python -m oracle.pin_against_reference
Writes tests/golden/*.pt (small) — committed; the GPU box has no /root/reference, so ``-m gpu`` tests
compare the CUDA path with the oracle *and* with these goldens.

Each case: run the reference's unmodified function (through oracle/ref_shim.py), run the oracle
restatement on the same seeded inputs, assert agreement, store outputs.  Inputs are regenerated from
seeds by oracle/fixtures.py (the same generator the tests use), so fixtures hold outputs only.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import torch

from . import comat_ref as R
from . import fixtures as FX
from . import ref_shim
from . import sd_modules as sdm

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _close(a, b, tol=1e-6, what=""):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    err = (a - b).abs().max().item() / max(1e-12, b.abs().max().item())
    assert err <= tol, f"{what}: rel err {err:.3e} > {tol}"
    return err


def pin_layer_loss():
    ref = ref_shim.import_reference("attn_utils.tc_loss_utils")
    out = {}
    for case in FX.LAYER_LOSS_CASES:
        maps, masks, words, res = FX.layer_loss_inputs(**case)
        m_ref = [m.clone().requires_grad_(True) for m in maps]
        d_ref = ref.get_grounding_loss_by_layer(masks, words, res, m_ref, False)
        m_or = [m.clone().requires_grad_(True) for m in maps]
        d_or = R.grounding_loss_by_layer(masks, words, res, m_or)
        rec = {"case": case}
        for k in ("token_loss", "pixel_loss"):
            _close(d_or[k], d_ref[k], 1e-6, f"layer_loss {case} {k}")
            rec[k] = float(d_ref[k])
        if len(words):
            (d_ref["token_loss"] * 1.0 + 0.5 * d_ref["pixel_loss"]).backward()
            (d_or["token_loss"] * 1.0 + 0.5 * d_or["pixel_loss"]).backward()
            for a, b in zip(m_or, m_ref):
                _close(a.grad, b.grad, 1e-5, f"layer_loss grad {case}")
            rec["grad_l2"] = [float(b.grad.double().norm()) for b in m_ref]
            rec["grad_probe"] = [b.grad.flatten()[:: max(1, b.grad.numel() // 64)][:64].clone() for b in m_ref]
        rec["mask_resized_sum"] = [float(R.resize_mask(m, res).sum()) for m in masks]
        out[FX.case_key(case)] = rec
    return out


def pin_mask_loss():
    gs = ref_shim.import_reference("attr_concen_utils.gsam_interface")
    out = {}
    for case in FX.MASK_LOSS_CASES:
        attn_dict, subtrees, idx2wp, masks_by_sample, layers, B = FX.mask_loss_inputs(**case)
        calls = {"i": 0}

        def get_mask(image, nouns):
            i = calls["i"]
            calls["i"] += 1
            return masks_by_sample[get_mask.order[i]]

        # the reference calls get_mask only for samples that survive the noun filter, in sample order
        self_ns = SimpleNamespace(train_layer_ls=layers)
        self_ns.update_nouns_attributes = lambda n, a: gs.GsamSegModel.update_nouns_attributes(self_ns, n, a)
        surviving, words = [], []
        for b in range(B):
            nouns, attrs = R.words_from_subtrees(subtrees[b], idx2wp[b], self_ns.update_nouns_attributes)
            words.append(attrs)
            if len(nouns):
                surviving.append(b)
        get_mask.order = surviving
        self_ns.get_mask = get_mask
        images = torch.zeros(B, 3, 64, 64)
        ad_ref = {t: {k: [m.clone().requires_grad_(True) for m in v] for k, v in d.items()} for t, d in attn_dict.items()}
        tok, pix, _ = gs.GsamSegModel.get_mask_loss(self_ns, images, [""] * B, subtrees, idx2wp, ad_ref)
        ad_or = {t: {k: [m.clone().requires_grad_(True) for m in v] for k, v in d.items()} for t, d in attn_dict.items()}
        masks_in = [masks_by_sample[b] if b in surviving else None for b in range(B)]
        tok2, pix2 = R.mask_loss(ad_or, words, masks_in, layers, images)
        _close(tok2, tok, 1e-6, f"mask_loss token {case}")
        _close(pix2, pix, 1e-6, f"mask_loss pixel {case}")
        (1e-3 * tok + 5e-5 * pix).backward()
        (1e-3 * tok2 + 5e-5 * pix2).backward()
        gl2 = {}
        for t in ad_ref:
            for k in ad_ref[t]:
                for a, b in zip(ad_or[t][k], ad_ref[t][k]):
                    if b.grad is not None:
                        _close(a.grad, b.grad, 1e-5, f"mask_loss grad {case} {t} {k}")
                gl2[f"{t}/{k}"] = [float(b.grad.double().norm()) if b.grad is not None else 0.0 for b in ad_ref[t][k]]
        out[FX.case_key(case)] = {"case": case, "token_loss": float(tok), "pixel_loss": float(pix), "grad_l2": gl2,
                                  "words": words, "surviving": surviving}
    return out


def pin_attention_store_and_pipeline():
    """Reference AttrConcenTrainableSDPipeline.forward + register_attention_control + AttentionStore, verbatim,
    on the restated (tiny-geometry) UNet/VAE/scheduler; vs oracle rollout()."""
    ref_shim.install()
    tca = ref_shim.import_reference("attn_utils.tc_attn_utils")
    pl = ref_shim.import_reference("AttrConcenTrainableSDPipeline")
    out = {}
    for case in FX.PIPELINE_CASES:
        w = FX.pipeline_world(**case)
        S, T, A = case["S"], w["training_steps"], w["attrcon_steps"]
        # --- reference run
        unet = w["make_unet"]()
        ctrl = tca.AttentionStore(w["train_layer_ls"])
        n = tca.register_attention_control(unet, ctrl)
        n_layers = ctrl.num_att_layers
        pipe = pl.AttrConcenTrainableSDPipeline.__new__(pl.AttrConcenTrainableSDPipeline)
        ref_shim._PipelineBase.__init__(pipe, vae=w["vae"], text_encoder=None, tokenizer=None, unet=unet,
                                        scheduler=sdm.DDPMScheduler())
        pipe.parser = lambda p: p
        pipe.attn_dict = {}
        pipe.controller = ctrl
        gen = torch.Generator().manual_seed(case["seed"] + 77)
        prompts = ["p%d" % i for i in range(case["B"])]
        image, lat = pipe.forward(prompt=prompts, height=case["hw"] * 8, width=case["hw"] * 8, training_timesteps=T,
                                  detach_gradient=True, train_text_encoder=False, num_inference_steps=S,
                                  guidance_scale=7.5, guidance_rescale=case.get("rescale", 0.0),
                                  negative_prompt_embeds=w["null_embeds"], prompt_embeds=w["prompt_embeds"],
                                  latents=w["latents"].clone(), generator=gen, early_exit=False, return_latents=True,
                                  bp_on_trained=True, double_laststep=False, fast_training=False,
                                  attrcon_train_steps=A)
        torch.set_grad_enabled(True)
        ref_attn = pipe.attn_dict
        params_ref = [p for p in unet.parameters() if p.requires_grad]
        loss_ref = (image.float() ** 2).mean() + sum((m.float() ** 2).sum() for d in ref_attn.values() for v in d.values() for m in v) * 1e-3
        g_ref = torch.autograd.grad(loss_ref, params_ref, allow_unused=True)
        # --- oracle run (noise pre-drawn from the same generator stream)
        unet2 = w["make_unet"]()
        ctrl2 = R.AttentionStore(w["train_layer_ls"])
        assert R.register_attention_control(unet2, ctrl2) == n_layers
        gen2 = torch.Generator().manual_seed(case["seed"] + 77)
        noises = [torch.randn(w["latents"].shape, generator=gen2) for _ in range(S)]
        image2, lat2, attn2 = R.rollout(unet2, w["vae"], sdm.DDPMScheduler(), w["prompt_embeds"], w["null_embeds"],
                                        w["latents"].clone(), noises, S, T, 7.5, case.get("rescale", 0.0), A, ctrl2,
                                        return_latents=True)
        _close(image2, image, 1e-5, f"pipeline image {case}")
        _close(lat2, lat, 1e-5, f"pipeline latents {case}")
        assert set(attn2.keys()) == set(ref_attn.keys()), (attn2.keys(), ref_attn.keys())
        keyset = {}
        for t in ref_attn:
            assert set(attn2[t].keys()) == set(ref_attn[t].keys())
            for k in ref_attn[t]:
                assert len(attn2[t][k]) == len(ref_attn[t][k])
                keyset[k] = len(ref_attn[t][k])
                for a, b in zip(attn2[t][k], ref_attn[t][k]):
                    _close(a, b, 1e-5, f"pipeline attn {t} {k}")
        params2 = [p for p in unet2.parameters() if p.requires_grad]
        loss2 = (image2.float() ** 2).mean() + sum((m.float() ** 2).sum() for d in attn2.values() for v in d.values() for m in v) * 1e-3
        g2 = torch.autograd.grad(loss2, params2, allow_unused=True)
        for a, b in zip(g2, g_ref):
            if b is not None:
                _close(a, b, 2e-4, f"pipeline lora grad {case}")
        out[FX.case_key(case)] = {
            "case": case, "num_att_layers": n_layers, "keyset": keyset, "timesteps": sorted(ref_attn.keys()),
            "image_mean": float(image.double().mean()), "image_l2": float(image.double().norm()),
            "latents": lat.detach().clone(), "loss": float(loss_ref),
            "grad_l2": [float(g.double().norm()) if g is not None else 0.0 for g in g_ref],
        }
    return out


def pin_blip_score():
    cb = ref_shim.import_reference("concept_mat_utils.caption_blip")
    out = {}
    for case in FX.BLIP_CASES:
        model, images, ids, mask = FX.blip_inputs(**case)

        class _Tok:
            pad_token_id = 0

        class _Proc:
            tokenizer = _Tok()

            def __call__(self, images=None, text=None, return_tensors=None, padding=None):
                return {"pixel_values": images, "input_ids": ids.clone(), "attention_mask": mask.clone()}

        blip = object.__new__(cb.Blip)
        torch.nn.Module.__init__(blip)
        blip.processor, blip.model = _Proc(), model
        blip.transforms = cb.Compose([cb.Resize(size=(384, 384), interpolation=cb.InterpolationMode.BICUBIC, antialias=True),
                                      cb.Normalize(mean=list(R.CLIP_MEAN), std=list(R.CLIP_STD))])
        blip.prompt, blip.prompt_length = "a photography of", 4
        im_ref = images.clone().requires_grad_(True)
        r_ref = blip.score(im_ref, ["x"] * images.shape[0])
        im_or = images.clone().requires_grad_(True)
        r_or = R.blip_score(model, im_or, ids, mask, 4)
        _close(r_or, r_ref, 1e-6, f"blip reward {case}")
        (-r_ref).backward()
        (-r_or).backward()
        _close(im_or.grad, im_ref.grad, 1e-5, f"blip image grad {case}")
        out[FX.case_key(case)] = {"case": case, "reward": float(r_ref), "grad_l2": float(im_ref.grad.double().norm()),
                                  "grad_probe": im_ref.grad.flatten()[::4099][:64].clone()}
    return out


def pin_gan():
    gan = ref_shim.import_reference("training_utils.gan_sdxl")
    out = {}
    for case in FX.GAN_CASES:
        w = FX.gan_world(**case)
        self_ns = SimpleNamespace(unet=w["d_unet"], mlp=w["head"], ori_scheduler=sdm.DDPMScheduler(),
                                  cls_loss_fn=torch.nn.BCEWithLogitsLoss(), weight_dtype=torch.float32,
                                  D_args=SimpleNamespace(condition_discriminator=False, gan_unet_lastlayer_cls=False),
                                  set_D_sd_pipeline_lora=lambda requires_grad=True: None)
        self_ns.get_D_gt_noise = lambda device, **kw: kw["batch"]["latents"]
        zf = w["fake"].clone().requires_grad_(True)
        g_ref = gan.D_sd.D_sd_pipeline_forward(self_ns, zf, side="G", negative_prompt_embeds=w["null"],
                                               num_inference_steps=case["S"])
        d_ref = gan.D_sd.D_sd_pipeline_forward(self_ns, w["fake"].clone(), side="D", negative_prompt_embeds=w["null"],
                                               num_inference_steps=case["S"], batch={"latents": w["real"]})
        w["d_unet"].eval()
        zf2 = w["fake"].clone().requires_grad_(True)
        g_or = R.d_forward(w["d_unet"], w["head"], sdm.DDPMScheduler(), zf2, w["null"], case["S"], "G")
        d_or = R.d_forward(w["d_unet"], w["head"], sdm.DDPMScheduler(), w["fake"], w["null"], case["S"], "D", w["real"])
        _close(g_or, g_ref, 1e-6, "gan G")
        _close(d_or, d_ref, 1e-6, "gan D")
        out[FX.case_key(case)] = {"case": case, "G_loss": float(g_ref), "D_loss": float(d_ref)}
    return out


def pin_encode_prompt():
    """Reference TrainableSDPipeline.encode_prompt (TrainableSDPipeline.py:227-424), verbatim, over an HF CLIPTextModel (tiny
    geometry, seeded) and the framing-exact tokenizer stub; vs oracle encode_prompt_sd.  The SDXL twin lives in diffusers
    (un-vendored) and stays unpinned."""
    ref_shim.install()
    pl = ref_shim.import_reference("TrainableSDPipeline")
    out = {}
    for case in FX.ENCODE_PROMPT_CASES:
        enc = R.make_clip_text("clip_l", tiny=True, seed=case["seed"])
        tok = FX.ClipTokenizerStub()
        pipe = pl.TrainableSDPipeline.__new__(pl.TrainableSDPipeline)
        ref_shim._PipelineBase.__init__(pipe, vae=None, text_encoder=enc, tokenizer=tok, unet=None, scheduler=None)
        with torch.no_grad():
            pe_ref, npe_ref = pipe.encode_prompt(case["prompts"], torch.device("cpu"), case["n_per"], case["cfg"],
                                                 negative_prompt=None if case["negative"] is None else [case["negative"]] * len(case["prompts"]),
                                                 clip_skip=case["clip_skip"])
        neg = case["negative"]
        pe, npe, ids = R.encode_prompt_sd(enc, tok, case["prompts"], case["n_per"], case["cfg"],
                                          negative_prompt=None if neg is None else [neg] * len(case["prompts"]),
                                          clip_skip=case["clip_skip"])
        _close(pe, pe_ref, 1e-6, f"encode_prompt embeds {case['seed']}")
        if case["cfg"]:
            _close(npe, npe_ref, 1e-6, f"encode_prompt negative embeds {case['seed']}")
        else:
            assert npe_ref is None and npe is None
        out["seed%d" % case["seed"]] = {"case": case, "input_ids": ids.clone(), "prompt_embeds": pe_ref.clone(),
                                        "negative_prompt_embeds": None if npe_ref is None else npe_ref.clone()}
    return out


def reference_function(rel_path: str, name: str, namespace: dict):
    """compile ONE top-level function (or class) of a reference file that cannot be imported whole (training_script.py imports accelerate)
    from its own source text - nothing is copied into the repo; the function body runs verbatim."""
    import ast
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, rel_path)).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == name)
    code = compile(ast.Module(body=[fn], type_ignores=[]), os.path.join(ref_shim.REFERENCE_ROOT, rel_path), "exec")
    exec(code, namespace)
    return namespace[name]


def pin_lora_state_dict():
    """Reference ``unet_lora_state_dict`` (training_script.py:50-66), verbatim, on the restated diffusers-shaped UNets with LoRA
    installed (meta device: names and shapes only) -> the tensor names / shapes of pytorch_lora_weights.safetensors before
    diffusers' own ``unet.`` prefixing.  Written as JSON (tests/golden/lora_state_dict_keys.json)."""
    import json
    fn = reference_function("training_script.py", "unet_lora_state_dict", {"UNet2DConditionModel": sdm.UNet2DConditionModel, "torch": torch})
    out = {}
    for name, cfg, rank in (("sd15_r128", {}, 128), ("sdxl_r128", sdm.SDXL_UNET_CONFIG, 128), ("tiny_r4", sdm.tiny_unet_config(width=64, cross_attention_dim=64), 4)):
        with torch.device("meta"):
            unet = sdm.UNet2DConditionModel(**cfg)
            sdm.install_lora(unet, rank)
        sd = fn(unet)
        out[name] = {"n": len(sd), "numel": int(sum(v.numel() for v in sd.values())), "keys": [[k, list(v.shape)] for k, v in sd.items()]}
    with open(os.path.join(GOLDEN_DIR, "lora_state_dict_keys.json"), "w") as f:
        json.dump(out, f)
    print({k: (v["n"], v["numel"]) for k, v in out.items()})
    return None


def pin_attr_align():
    """Reference ``attribute_concen_utils`` (imported as is) and ``AttrConcenTrainableSDPipeline._extract_attribution_indices`` /
    ``_align_indices`` / ``unify_lists`` (verbatim through the shim) over hand-written dependency parses and a CLIP-convention
    word-piece stub -> tests/golden/attr_align.json (token texts per extractor, aligned CLIP positions, index -> word piece)."""
    import json
    ref_shim.install()
    acu = ref_shim.import_reference("attribute_concen_utils")
    pl = ref_shim.import_reference("AttrConcenTrainableSDPipeline")
    out = {}
    for prompt, spec in FX.ATTR_ALIGN_CASES.items():
        doc = FX.fake_doc(spec)
        tok = FX.BpeStub(FX.ATTR_ALIGN_SPLITS)
        pipe = pl.AttrConcenTrainableSDPipeline.__new__(pl.AttrConcenTrainableSDPipeline)
        pipe.tokenizer, pipe.doc = tok, {prompt: doc}
        names = lambda groups: None if groups is None else [[t.i for t in g] for g in groups]
        out[prompt] = {
            "plain": names(acu.extract_attribution_indices(doc)),
            "with_verbs": names(acu.extract_attribution_indices_with_verbs(doc)),
            "verb_root": names(acu.extract_attribution_indices_with_verb_root(doc)),
            "aligned": pipe._extract_attribution_indices(prompt),
            "idx_to_wp": {str(k): v for k, v in acu.get_attention_map_index_to_wordpiece(tok, prompt).items()},
        }
    with open(os.path.join(GOLDEN_DIR, "attr_align.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, "->", v["aligned"])
    return None


def pin_sdxl_pipeline():
    """Reference AttrConcenTrainableSDXLPipeline.forward + _attrcon_forward (AttrConcenTrainableSDXLPipeline.py:234-496) with the
    SDXL attention hook (attn_utils/tc_sdxl_attn_utils.py), verbatim, on the restated tiny SDXL-geometry UNet / VAE (VAE in
    fp16 as the reference decodes ``latents.half()``, :438-444); vs oracle rollout(sdxl=True).  Pins rows a3 / a5: always-detached
    UNet input, pooled-text + time-id conditioning split per CFG half, un-rescaled image with ``return_latents``."""
    from types import SimpleNamespace
    ref_shim.install()
    tca = ref_shim.import_reference("attn_utils.tc_sdxl_attn_utils")
    pl = ref_shim.import_reference("AttrConcenTrainableSDXLPipeline")
    out = {}
    for case in FX.SDXL_PIPELINE_CASES:
        w = FX.pipeline_world(**case, sdxl=True)
        S, T, A, B, hw = case["S"], w["training_steps"], w["attrcon_steps"], case["B"], case["hw"]
        g = torch.Generator().manual_seed(case["seed"] + 9)
        pooled, npooled = torch.randn(B, 16, generator=g), torch.randn(B, 16, generator=g)
        vae = w["vae"].half()
        layers = ["up_8", "up_16"]
        # --- reference run
        unet = w["make_unet"]()
        ctrl = tca.AttentionStore(layers)
        tca.register_attention_control(unet, ctrl)
        n_layers = ctrl.num_att_layers
        pipe = pl.AttrConcenTrainableSDXLPipeline.__new__(pl.AttrConcenTrainableSDXLPipeline)
        sys.modules["diffusers"].StableDiffusionXLPipeline.__init__(
            pipe, vae=vae, text_encoder=None, text_encoder_2=SimpleNamespace(config=SimpleNamespace(projection_dim=16)),
            tokenizer=None, tokenizer_2=None, unet=unet, scheduler=sdm.DDPMScheduler())
        pipe.parser = lambda p: p
        pipe.attn_dict = {}
        pipe.controller = ctrl
        gen = torch.Generator().manual_seed(case["seed"] + 77)
        prompts = ["p%d" % i for i in range(B)]
        image, lat = pipe.forward(prompt=prompts, height=hw * 8, width=hw * 8, training_timesteps=T, detach_gradient=True,
                                  train_text_encoder=False, num_inference_steps=S, guidance_scale=7.5,
                                  guidance_rescale=case.get("rescale", 0.0), prompt_embeds=w["prompt_embeds"],
                                  negative_prompt_embeds=w["null_embeds"], pooled_prompt_embeds=pooled,
                                  negative_pooled_prompt_embeds=npooled, latents=w["latents"].clone(), generator=gen, early_exit=False,
                                  return_latents=True, attrcon_train_steps=A)
        torch.set_grad_enabled(True)
        ref_attn = pipe.attn_dict
        params_ref = [p for p in unet.parameters() if p.requires_grad]
        loss_ref = (image.float() ** 2).mean() + sum((m.float() ** 2).sum() for d in ref_attn.values() for v in d.values() for m in v) * 1e-3
        g_ref = torch.autograd.grad(loss_ref, params_ref, allow_unused=True)
        # --- oracle run (noise pre-drawn from the same generator stream)
        unet2 = w["make_unet"]()
        ctrl2 = R.AttentionStore(layers)
        assert R.register_attention_control(unet2, ctrl2) == n_layers
        gen2 = torch.Generator().manual_seed(case["seed"] + 77)
        noises = [torch.randn(w["latents"].shape, generator=gen2) for _ in range(S)]
        ids = torch.tensor([[hw * 8., hw * 8, 0, 0, hw * 8, hw * 8]]).repeat(B, 1)
        added = {"text_embeds": torch.cat([npooled, pooled]), "time_ids": torch.cat([ids, ids])}
        image2, lat2, attn2 = R.rollout(unet2, vae, sdm.DDPMScheduler(), w["prompt_embeds"], w["null_embeds"], w["latents"].clone(),
                                        noises, S, T, 7.5, case.get("rescale", 0.0), A, ctrl2, added_cond_kwargs=added, sdxl=True,
                                        return_latents=True)
        _close(lat2.half().float(), lat.float(), 1e-5, f"sdxl pipeline latents {case}")
        _close(image2.float(), image.float(), 2e-3, f"sdxl pipeline image (fp16 VAE) {case}")
        assert set(attn2.keys()) == set(ref_attn.keys()), (attn2.keys(), ref_attn.keys())
        keyset = {}
        for t in ref_attn:
            assert set(attn2[t].keys()) == set(ref_attn[t].keys())
            for k in ref_attn[t]:
                assert len(attn2[t][k]) == len(ref_attn[t][k])
                keyset[k] = len(ref_attn[t][k])
                for a, b in zip(attn2[t][k], ref_attn[t][k]):
                    _close(a, b, 1e-5, f"sdxl pipeline attn {t} {k}")
        params2 = [p for p in unet2.parameters() if p.requires_grad]
        loss2 = (image2.float() ** 2).mean() + sum((m.float() ** 2).sum() for d in attn2.values() for v in d.values() for m in v) * 1e-3
        g2 = torch.autograd.grad(loss2, params2, allow_unused=True)
        n_checked = 0
        for a, b in zip(g2, g_ref):
            if b is not None and float(b.abs().max()) > 0:
                _close(a, b, 5e-3, f"sdxl pipeline lora grad {case}")
                n_checked += 1
        assert n_checked > 0
        out[FX.case_key(case)] = {
            "case": case, "num_att_layers": n_layers, "keyset": keyset, "timesteps": sorted(ref_attn.keys()),
            "image_mean": float(image.double().mean()), "image_l2": float(image.double().norm()),
            "latents": lat.detach().clone(), "pooled": pooled, "npooled": npooled, "loss": float(loss_ref),
            "grad_l2": [float(g.double().norm()) if g is not None else 0.0 for g in g_ref],
        }
    return out


def pin_lora_install():
    """Reference ``set_pipeline_trainable_module`` + ``get_trainable_parameters`` (training_utils/pipeline.py:84-121, :123-187; row
    a14), verbatim through the shim, on the restated UNets (meta device): which projections carry LoRA, rank, dtype, and the order
    of the trainable-parameter list -> tests/golden/lora_trainable_params.json."""
    import argparse
    import json
    from types import SimpleNamespace
    ref_shim.install()
    tp = ref_shim.import_reference("training_utils.pipeline")
    out = {}
    for name, cfg, rank in (("sd15_r128", {}, 128), ("sdxl_r128", sdm.SDXL_UNET_CONFIG, 128), ("tiny_r4", sdm.tiny_unet_config(width=64, cross_attention_dim=64), 4)):
        args = argparse.Namespace(full_finetuning=False, lora_rank=rank, train_text_encoder_lora=False, tune_vae=False, tune_text_encoder=False)
        with torch.device("meta"):
            unet = sdm.UNet2DConditionModel(**cfg)
            pipe = SimpleNamespace(unet=unet, vae=None, text_encoder=None)
            tp.set_pipeline_trainable_module(args, pipe)
            G, text = tp.get_trainable_parameters(args, pipe)
        names = {id(p): n for n, p in unet.named_parameters()}
        assert text == [] and all(p.dtype == torch.float32 for p in G)
        out[name] = {"n": len(G), "numel": int(sum(p.numel() for p in G)), "params": [[names[id(p)], list(p.shape)] for p in G]}
    with open(os.path.join(GOLDEN_DIR, "lora_trainable_params.json"), "w") as f:
        json.dump(out, f)
    print({k: (v["n"], v["numel"], v["params"][0]) for k, v in out.items()})
    return None


PINS = [("layer_loss", pin_layer_loss), ("mask_loss", pin_mask_loss), ("blip_score", pin_blip_score),
        ("gan", pin_gan), ("pipeline", pin_attention_store_and_pipeline), ("encode_prompt", pin_encode_prompt),
        ("lora_state_dict", pin_lora_state_dict), ("attr_align", pin_attr_align),
        ("sdxl_pipeline", pin_sdxl_pipeline), ("lora_install", pin_lora_install)]


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(8)
    only = set(sys.argv[1:])            # e.g. `python -m oracle.pin_against_reference encode_prompt` re-pins one golden file
    for name, fn in PINS:
        if only and name not in only:
            continue
        res = fn()
        if res is None:                   # the pin wrote its own (JSON) golden
            continue
        path = os.path.join(GOLDEN_DIR, f"{name}.pt")
        torch.save(res, path)
        print(f"pinned {name}: {len(res)} cases -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    sys.exit(main())
