"""CPU: host-side logic of the product package + C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import pytest
import torch

from oracle import comat_ref as R
from oracle import fixtures as FX

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
    from comat_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    hdr = open(os.path.join(ROOT, "include", "comat_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(comat_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 7
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/comat_b200.h but not exported"
    lib.comat_version.restype = ctypes.c_int
    assert lib.comat_version() >= 100
    lib.comat_strerror.restype = ctypes.c_char_p
    assert lib.comat_strerror(-1) == b"invalid argument"


def test_product_has_no_cpu_fallback():
    from comat_b200 import _lib, attn_loss
    with pytest.raises(_lib.ComatError):
        attn_loss.mask_resize_any(torch.zeros(1, 8, 8, dtype=torch.uint8), 4)


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; import comat_b200, comat_b200.attn_loss, comat_b200._lib; "
            "bad=[m for m in sys.modules if m=='oracle' or m.startswith('oracle.')]; assert not bad, bad")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "comat_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_noun_filter_matches_oracle():
    from comat_b200 import attn_loss
    nouns = ["dog", "cat", "dog", "sky", "waves", "skys", "apple"]
    attrs = [[1], [2, 3], [4], [5], [6], [7], [8, 9]]
    assert attn_loss.update_nouns_attributes(nouns, attrs) == FX.update_nouns_attributes(nouns, attrs) == (["cat", "apple"], [[2, 3], [8, 9]])
    st = [[3, [4, 5]], [7], []]
    wp = {3: "big", 4: "ap", 5: "ple", 7: "sea"}
    assert attn_loss.words_from_subtrees(st, wp) == (["apple"], [[3, 4, 5]])
    assert R.words_from_subtrees(st, wp, FX.update_nouns_attributes) == (["apple"], [[3, 4, 5]])


def _aa_any_emulation(mask, res):
    """python emulation of comat_mask_resize_any (csrc/attnmap_loss.cu) — validates the window rule against torchvision."""
    import math
    H, W = mask.shape
    out = torch.zeros(res, res)

    def win(o, scale, n):
        support = scale if scale >= 1 else 1.0
        inv = 1.0 / scale if scale >= 1 else 1.0
        c = scale * (o + 0.5)
        lo, hi = max(int(c - support + 0.5), 0), min(int(c + support + 0.5), n)
        return [j for j in range(lo, hi) if 1.0 - abs((j - c + 0.5) * inv) > 0]
    sy, sx = H / res, W / res
    for y in range(res):
        ys = win(y, sy, H)
        for x in range(res):
            xs = win(x, sx, W)
            out[y, x] = float(mask[ys][:, xs].any())
    return out


@pytest.mark.parametrize("size,res", [(512, 8), (512, 64), (64, 8), (96, 16), (100, 16)])
def test_mask_resize_rule_matches_torchvision(size, res):
    g = torch.Generator().manual_seed(size + res)
    m = FX.random_mask(g, size)
    m[0, 0, 3, 5] = True
    ref = R.resize_mask(m, res)[0]
    assert torch.equal(_aa_any_emulation(m[0, 0], res), ref)
    single = torch.zeros(1, 1, size, size, dtype=torch.bool)
    single[0, 0, size // 2, size // 2] = True
    assert torch.equal(_aa_any_emulation(single[0, 0], res), R.resize_mask(single, res)[0])


def test_select_training_steps():
    import random
    steps, attr = R.select_training_steps(20, 5, random.Random(0), 2)
    assert len(steps) == 5 and steps[1] - steps[0] == 4 and all(a in steps for a in attr)
    steps, _ = R.select_training_steps(50, 5, random.Random(1), 2)
    assert steps[-1] <= 49 and steps[1] - steps[0] == 10


@pytest.mark.needs_reference
def test_caption_model_wrapper_vs_reference_class():
    """row a12: the reference's own ``CaptionModelWrapper`` (training_script.py:69-97; the file imports accelerate, so the class is
    compiled from its source text, nothing copied) vs the product's, over one fake scorer: weights * reward and 'total'."""
    import torch
    from oracle.pin_against_reference import reference_function
    from comat_b200.caption import CaptionModelWrapper

    class Scorer:
        def score(self, images, prompts, **kw):
            return images.mean() * len(prompts)

    def load_model(self, caption_model, device, args):            # concept_mat_utils/load_captionmodel.py:3-8 needs Hub weights
        self.blip_model = Scorer()
    Ref = reference_function("training_script.py", "CaptionModelWrapper", {"torch": torch, "load_model": load_model})
    ref = Ref(["Blip"], [0.75], "cpu", None, torch.float32)
    ours = CaptionModelWrapper(["Blip"], [0.75], Scorer())
    x = torch.rand(2, 3, 8, 8)
    a, b = ref(x, ["p", "q"], batch={}), ours(x, ["p", "q"], batch={})
    assert set(a) == set(b) == {"Blip", "total"}
    assert torch.equal(a["Blip"], b["Blip"]) and torch.equal(a["total"], b["total"])


@pytest.mark.needs_reference
@pytest.mark.parametrize("arch,kind", [("gansd_1_5", "D_sd"), ("gan_sd_1_5", None), ("sd_1_5", "D_sd"), ("ganother", None)])
def test_load_discriminator_arch_strings_vs_reference(arch, kind, monkeypatch):
    """gan_sd_model.py:8-14 run verbatim (D_sd patched to a marker): which ``--gan_model_arch`` strings resolve to a discriminator."""
    import argparse
    from oracle import ref_shim
    from comat_b200 import gan as G
    ref = ref_shim.import_reference("training_utils.gan_sd_model")
    monkeypatch.setattr(ref, "D_sd", lambda *a: "D_sd")
    got_ref = ref.load_discriminator(argparse.Namespace(gan_model_arch=arch), None, None)
    monkeypatch.setattr(G, "D_sd", lambda unet: "D_sd")
    got = G.load_discriminator(argparse.Namespace(gan_model_arch=arch, gan_unet_lastlayer_cls=False, condition_discriminator=False), None)
    assert got_ref == got == kind


@pytest.mark.needs_reference
@pytest.mark.parametrize("S,K,n_attr", [(20, 5, 2), (50, 5, 2), (50, 1, 5), (7, 3, 2), (4, 4, 9)])
def test_step_selection_vs_reference_statements(S, K, n_attr):
    """row a1, step selection: the reference's own statements (training_script.py:563-566 ``interval / max_start / start /
    training_steps`` and :589-590 ``random.choices``) are lifted from the AST of ``Trainer.train`` and executed on a seeded
    ``random.Random``; the product's ``CoMatTrainer.select_steps`` must draw the same steps from the same stream."""
    import ast
    import argparse
    import random
    from oracle import ref_shim
    from comat_b200.trainer import CoMatTrainer
    tree = ast.parse(open(os.path.join(ref_shim.REFERENCE_ROOT, "training_script.py")).read())
    train = next(n for c in tree.body if isinstance(c, ast.ClassDef) and c.name == "Trainer"
                 for n in c.body if isinstance(n, ast.FunctionDef) and n.name == "train")
    want = ["interval", "max_start", "start", "training_steps"]
    picked = []
    for node in ast.walk(train):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id in want:
            picked.append(node)
        if (isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Subscript) and
                getattr(node.targets[0].slice, "value", None) == "attrcon_train_steps"):
            picked.append(node)
    picked.sort(key=lambda n: n.lineno)
    assert [getattr(n.targets[0], "id", "kwargs") for n in picked] == want + ["kwargs"]
    rng = random.Random(1234)
    ns = {"random": rng, "total_step": S, "args": argparse.Namespace(K=K, attrcon_train_steps=n_attr), "kwargs": {}}
    exec(compile(ast.Module(body=picked, type_ignores=[]), "training_script.py", "exec"), ns)
    tr = CoMatTrainer.__new__(CoMatTrainer)
    tr.args = argparse.Namespace(K=K, total_step=S, attrcon_train_steps=n_attr)
    tr.rng, tr.attrcon = random.Random(1234), True
    steps, attr = tr.select_steps()
    assert steps == ns["training_steps"] and attr == ns["kwargs"]["attrcon_train_steps"]
    assert len(steps) >= K and all(0 <= s < S for s in steps)


@pytest.mark.needs_reference
@pytest.mark.parametrize("res", [512, 256, 768, 1024])
def test_reward_crop_vs_reference_statements(res):
    """row a1, reward crop (training_script.py:606-611): offsets and size of the crop handed to the captioner, lifted from the
    reference's AST, vs the arithmetic in ``CoMatTrainer.g_losses`` (same RNG stream: x offset first, then y)."""
    import ast
    import argparse
    import random
    import re
    from oracle import ref_shim
    tree = ast.parse(open(os.path.join(ref_shim.REFERENCE_ROOT, "training_script.py")).read())
    want = ["offset_range", "random_offset_x", "random_offset_y", "size"]
    picked = sorted((n for n in ast.walk(tree) if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Name) and n.targets[0].id in want),
                    key=lambda n: n.lineno)
    assert [n.targets[0].id for n in picked] == want
    ns = {"random": random.Random(7), "args": argparse.Namespace(resolution=res)}
    exec(compile(ast.Module(body=picked, type_ignores=[]), "training_script.py", "exec"), ns)
    # the product's statements, taken from its own source so the test follows the code
    src = open(os.path.join(ROOT, "comat_b200", "trainer.py")).read()
    m = re.search(r"off = a\.resolution // 224.*?\n\s+ox, oy = batch\.get\(\"crop\"\) or \(self\.rng\.randint\(0, off\), self\.rng\.randint\(0, off\)\)\n\s+size = a\.resolution - off", src)
    assert m, "trainer.py crop statements changed: update this test"
    rng = random.Random(7)
    off = res // 224
    ox, oy = rng.randint(0, off), rng.randint(0, off)
    assert (off, ox, oy, res - off) == (ns["offset_range"], ns["random_offset_x"], ns["random_offset_y"], ns["size"])


@pytest.mark.needs_reference
@pytest.mark.parametrize("norm_grad", [False, True])
def test_record_grad_hook_vs_reference(norm_grad):
    """row a19: the reference's own ``record_grad`` closure (training_script.py:644-651, lifted from the AST) vs the hook
    ``CoMatTrainer.g_losses`` registers on the image (same rescaling; the norm stays a device scalar instead of ``.item()``)."""
    import ast
    import argparse
    from oracle import ref_shim
    tree = ast.parse(open(os.path.join(ref_shim.REFERENCE_ROOT, "training_script.py")).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "record_grad")
    ns = {"norm": {}, "args": argparse.Namespace(norm_grad=norm_grad)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "training_script.py", "exec"), ns)
    g = torch.randn(2, 3, 16, 16, generator=torch.Generator().manual_seed(1)) * 3e-4
    want = ns["record_grad"](g.clone())
    # the product's hook, exercised through a real backward on a stand-in image
    from comat_b200.trainer import CoMatTrainer
    tr = CoMatTrainer.__new__(CoMatTrainer)
    tr.args = argparse.Namespace(norm_grad=norm_grad, resolution=16, gan_loss=False, cfg_scale=7.5, cfg_rescale=0.0, total_step=1,
                                 do_classifier_free_guidance=False)
    tr.attrcon, tr.D, tr.rng, tr.train_text = False, None, None, False
    image = torch.zeros(2, 3, 16, 16, requires_grad=True)
    leaf = image * 1.0
    tr.pipeline = type("P", (), {"is_sdxl": False, "attn_dict": {}, "forward": lambda self, **kw: leaf})()
    tr.caption_model = lambda img, text, batch=None: {"total": img.sum() * 0.0, "Blip": img.sum() * 0.0}
    logs = tr.g_losses({"prompt_embeds": None, "training_steps": [0], "attrcon_steps": None, "crop": (0, 0)})
    leaf.backward(g.clone())
    assert abs(float(logs["_norm_holder"]["reward_norm"]) - ns["norm"]["reward_norm"]) <= 1e-6 * ns["norm"]["reward_norm"]
    torch.testing.assert_close(image.grad, want, rtol=1e-6, atol=0)


@pytest.mark.needs_reference
def test_optimizer_hyper_parameter_wiring_vs_reference():
    """row a20: which ``args.*`` feed AdamW / clipping of the generator and the discriminator - read off the reference's own
    ``optimizer_cls(...)`` and ``clip_grad_norm_(...)`` calls (AST of training_script.py:215-275, :661, :692) - vs the values
    ``CoMatTrainer`` hands its two ``FlatAdamW`` instances when every flag carries a distinct value."""
    import ast
    import argparse
    from oracle import ref_shim
    tree = ast.parse(open(os.path.join(ref_shim.REFERENCE_ROOT, "training_script.py")).read())

    def arg_name(node):
        return node.attr if isinstance(node, ast.Attribute) and getattr(node.value, "id", "") == "args" else None
    wiring = {}
    for call in (n for n in ast.walk(tree) if isinstance(n, ast.Call) and getattr(n.func, "id", "") == "optimizer_cls"):
        params = ast.unparse(call.args[0])
        if params not in ("self.G_parameters", "self.D_parameters"):
            continue                                               # the text-encoder-LoRA param-group variant (:239-252)
        kw = {k.arg: k.value for k in call.keywords}
        wiring[params] = {"lr": arg_name(kw["lr"]), "betas": tuple(arg_name(e) for e in kw["betas"].elts),
                          "wd": arg_name(kw["weight_decay"]), "eps": arg_name(kw["eps"])}
    for call in (n for n in ast.walk(tree) if isinstance(n, ast.Call) and getattr(n.func, "attr", "") == "clip_grad_norm_"):
        wiring[ast.unparse(call.args[0])]["clip"] = arg_name(call.args[1])
    assert set(wiring) == {"self.G_parameters", "self.D_parameters"}
    vals = dict(learning_rate=1e-3, learning_rate_D=2e-3, adam_beta1=0.11, adam_beta2=0.22, adam_beta1_D=0.33, adam_beta2_D=0.44,
                adam_weight_decay=0.55, adam_epsilon=6e-7, max_grad_norm=0.77, max_grad_norm_D=0.88)
    from comat_b200.trainer import CoMatTrainer
    mk = lambda: [torch.nn.Parameter(torch.zeros(3))]
    pipe = type("P", (), {"unet": type("U", (), {"lora_parameters": lambda self: mk(), "refresh_lora": lambda self, **k: None})()})()
    D = type("D", (), {"get_trainable_parameters": lambda self: mk(), "unet": pipe.unet})()
    tr = CoMatTrainer(argparse.Namespace(seed=0, pretrain_model_name="sd_1_5", **vals), pipe, None, D)
    for key, opt in (("self.G_parameters", tr.optimizer), ("self.D_parameters", tr.D_optimizer)):
        w = wiring[key]
        assert opt.lr == vals[w["lr"]] and tuple(opt.betas) == tuple(vals[b] for b in w["betas"])
        assert opt.wd == vals[w["wd"]] and opt.eps == vals[w["eps"]] and opt.max_norm == vals[w["clip"]]


def test_seg_model_adapter_has_the_reference_call_signature(monkeypatch):
    """SegModel.get_mask_loss(images, prompt, all_subtree_indices, attn_map_idx_to_wp_all, attn_map) -> 3 values, with the reference's
    noun / attribute bookkeeping (gsam_interface.py:163-202) in front of the fused loss."""
    import torch
    from comat_b200 import attn_loss
    seen = {}

    def fake_loss(attn_map, words, masks, layers, tokens=77):
        seen.update(words=words, masks=masks, layers=list(layers))
        return torch.tensor(2.0), torch.tensor(3.0)
    monkeypatch.setattr(attn_loss, "get_mask_loss", fake_loss)
    calls = []

    def mask_fn(image, nouns):
        calls.append(list(nouns))
        return None if nouns == ["dog"] else [torch.ones(1, 1, 8, 8, dtype=torch.bool) for _ in nouns]
    seg = attn_loss.SegModel(["up_16"], mask_fn)
    wp = {1: "a", 2: "red", 3: "app", 4: "le", 5: "dog", 6: "street"}
    subtrees = [[[2, [3, 4]]], [[5]], [[6]], []]              # red apple | dog (no mask found) | street (stop-listed) | nothing
    tok, pix, d = seg.get_mask_loss(torch.zeros(4, 3, 8, 8), ["p"] * 4, subtrees, [wp] * 4, {"1": {"up_16": []}})
    assert (float(tok), float(pix)) == (2.0, 3.0) and set(d) == {"token/total", "pixel/total"}
    assert calls == [["apple"], ["dog"]]
    assert seen["words"] == [[[2, 3, 4]], [], [], []] and [m is None for m in seen["masks"]] == [False, True, True, True]


def test_load_discriminator_accepts_the_reference_signature(monkeypatch):
    import torch
    from types import SimpleNamespace
    from comat_b200 import gan, pipelines
    made = {}

    class FakeUNet:
        device = "cpu"

        def lora_parameters(self):
            return []

    def fake_from_pretrained(path, **kw):
        made.update(path=path, **kw)
        return SimpleNamespace(unet=FakeUNet(), vae=object(), text_encoder=object(), is_sdxl=False)
    monkeypatch.setattr(pipelines.TrainableSDPipeline, "from_pretrained", staticmethod(fake_from_pretrained))
    args = SimpleNamespace(gan_model_arch="gansd_1_5", pretrain_model="/models/sd15", lora_rank=128, gan_unet_lastlayer_cls=False,
                           condition_discriminator=False)
    D = gan.load_discriminator(args, torch.float16, "cpu")        # training_script.py:290 form
    assert isinstance(D, gan.D_sd) and made["path"] == "/models/sd15" and made["dtype"] == torch.float16 and made["lora_rank"] == 128
    assert D.D_sd_pipeline.unet is None and D.D_sd_pipeline.text_encoder is not None
    args.gan_model_arch = "gan_sd_1_5"                            # the default string resolves to nothing in the reference too
    assert gan.load_discriminator(args, torch.float16, "cpu") is None
