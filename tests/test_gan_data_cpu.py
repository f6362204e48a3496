"""CPU: GAN ground-truth producer / reader (tools/gan_gt_generate.py:109-193, training_utils/gan_dataset.py:28-74):
file formats, cache semantics, and the sampled latents vs the oracle's rollout on the same weights and noise stream (CUDA ops
emulated in torch - test infrastructure)."""
import json
import os
import random
import uuid
from types import SimpleNamespace

import torch

from oracle import comat_ref as R
from oracle import fixtures as FX
from oracle import sd_modules as sdm
from tests import cpu_ops_emulation as EMU


def test_short_uid_known_answers():
    from comat_b200.gan_data import _ALPHABET, short_uid
    # shortuuid's documented example: shortuuid.uuid(name="example.com") == 'exu3DTbj2ncsn9tLdLWspw'
    assert short_uid(uuid.uuid5(uuid.NAMESPACE_DNS, "example.com")) == "exu3DTbj2ncsn9tLdLWspw"
    assert short_uid(uuid.UUID(int=0)) == "2" * 22 and short_uid(uuid.UUID(int=57)) == "2" * 20 + "32"
    ids = {short_uid() for _ in range(200)}
    assert len(ids) == 200 and all(len(i) == 22 and set(i) <= set(_ALPHABET) for i in ids)


def _world(monkeypatch):
    EMU.install_blip(monkeypatch)
    from comat_b200 import containers as Cn, synthetic
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import TrainableSDPipeline
    from comat_b200.text_encoder import EngineCLIPText
    torch.manual_seed(0)
    cfg = dict(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=128)
    unet = Cn.UNet2DConditionModel(**cfg)
    vae = Cn.AutoencoderKL(block_out_channels=(64, 64, 128, 128))
    unet.requires_grad_(False); vae.requires_grad_(False)
    unet.install_lora(4, up_std=0.05)
    clip = R.make_clip_text("clip_l", tiny=True, seed=21)
    o_unet = sdm.UNet2DConditionModel(**sdm.tiny_unet_config(width=64, cross_attention_dim=128))
    o_unet.requires_grad_(False)
    sdm.install_lora(o_unet, 4)
    o_unet.load_state_dict(unet.state_dict())
    pipe = TrainableSDPipeline(EngineVAE(vae, torch.float32), EngineUNet(unet, torch.float32), text_encoder=EngineCLIPText(clip, torch.float32),
                               tokenizer=synthetic.SyntheticClipTokenizer())
    return pipe, o_unet, clip


def test_producer_files_cache_and_latents_vs_oracle(tmp_path, monkeypatch):
    from comat_b200 import gan_data as GD
    pipe, o_unet, clip = _world(monkeypatch)
    prompts = ["a red apple", "two dogs on a sofa", "a blue car", "snow on a hill", "a green bench"]
    index = str(tmp_path / "train_data" / "gan_train_data.jsonl")
    S, hw = 3, 128
    n = GD.generate_gan_ground_truth(pipe, prompts, index, batch_size=2, num_inference_steps=S, guidance_scale=7.5, height=hw, width=hw,
                                     generator=torch.Generator().manual_seed(11))
    assert n == 5 and pipe.unet.use_graphs is False                 # switch restored
    recs = GD.read_jsonl(index)
    assert [r["prompt"] for r in recs] == prompts and all(set(r) == {"prompt", "file_path"} for r in recs)
    lat_dir = os.path.join(os.path.dirname(index), "latents")
    assert sorted(os.listdir(lat_dir)) == sorted(os.path.basename(r["file_path"]) for r in recs)
    assert all(len(os.path.basename(r["file_path"])) == 22 + 3 for r in recs)
    got = [torch.load(r["file_path"]) for r in recs]
    assert all(t.shape == (4, hw // 8, hw // 8) and t.dtype == torch.float32 and t.device.type == "cpu" for t in got)
    # oracle: same generator stream (initial latents, then one variance noise per step, batch by batch)
    gen = torch.Generator().manual_seed(11)
    tok = FX.ClipTokenizerStub()
    k = 0
    for i in range(0, 5, 2):
        chunk = prompts[i:i + 2]
        pe, npe, _ = R.encode_prompt_sd(clip, tok, chunk, 1, True)
        z = torch.randn(len(chunk), 4, hw // 8, hw // 8, generator=gen)
        noises = [torch.randn(z.shape, generator=gen) for _ in range(S)]
        _, lat, _ = R.rollout(o_unet, None, sdm.DDPMScheduler(), pe, npe, z, noises, S, [], 7.5, decode=False)
        for j in range(len(chunk)):
            torch.testing.assert_close(got[k], lat[j], rtol=2e-4, atol=2e-4)
            k += 1
    # use_cache: only prompts not in the index are generated, appended to the same index
    n2 = GD.generate_gan_ground_truth(pipe, prompts + ["a new prompt", "a new prompt"], index, batch_size=4, num_inference_steps=1,
                                      height=hw, width=hw, use_cache=True, generator=torch.Generator().manual_seed(1))
    assert n2 == 1 and len(GD.read_jsonl(index)) == 6 and GD.read_jsonl(index)[-1]["prompt"] == "a new prompt"


def test_call_output_types(monkeypatch):
    pipe, _, _ = _world(monkeypatch)
    g = torch.Generator().manual_seed(3)
    r = pipe(["a cat"], height=64, width=64, num_inference_steps=2, generator=g, output_type="pt")
    assert r.images.shape == (1, 3, 64, 64) and float(r.images.min()) >= 0.0 and float(r.images.max()) <= 1.0
    pil = pipe(["a cat"], height=64, width=64, num_inference_steps=1, generator=g, output_type="pil").images
    assert len(pil) == 1 and pil[0].size == (64, 64)
    lat = pipe(["a cat"], height=64, width=64, num_inference_steps=1, generator=g, output_type="latent", return_dict=False)[0]
    assert lat.shape == (1, 4, 8, 8) and not lat.requires_grad


def test_reader_formats(tmp_path):
    from comat_b200 import gan_data as GD
    lat = [torch.randn(4, 8, 8) for _ in range(3)]
    paths = []
    for i, t in enumerate(lat):
        p = str(tmp_path / f"l{i}.pt")
        torch.save(t, p)
        paths.append(p)
    recs = [{"prompt": "p0", "file_path": paths[0]}, {"prompt": "p1", "file_path": [paths[1], paths[2]], "image": "x.png", "source": "hrs"}]
    jl = str(tmp_path / "d.jsonl")
    with open(jl, "w") as f:
        f.write("\n".join(json.dumps(r) for r in recs) + "\n")
    ds = GD.Gan_Dataset(SimpleNamespace(training_prompts=jl), rng=random.Random(0))
    assert len(ds) == 2
    e0 = ds[0]
    assert e0["text"] == "p0" and torch.equal(e0["latents"], lat[0]) and set(e0) == {"text", "latents"}
    e1 = ds[1]
    assert e1["source"] == "hrs" and "image" not in e1 and any(torch.equal(e1["latents"], t) for t in lat[1:])
    js = str(tmp_path / "d.json")
    json.dump(recs, open(js, "w"))
    assert len(GD.Gan_Dataset(SimpleNamespace(training_prompts=js))) == 2
    opened = []
    ds2 = GD.Gan_Dataset(SimpleNamespace(training_prompts=jl), opener=lambda p: (opened.append(p), open(p, "rb").read())[1])
    b = GD.collate_gan_batch([ds2[0], ds2[0]])
    assert opened == [paths[0]] * 2 and b["real_latents"].shape == (2, 4, 8, 8) and b["text"] == ["p0", "p0"]
    txt = str(tmp_path / "p.txt")
    open(txt, "w").write("a\nb\nc\n")
    assert GD.read_prompts(txt, 1, 3) == ["b", "c"]


import pytest


@pytest.mark.needs_reference
def test_reference_reader_reads_what_our_producer_writes(tmp_path, monkeypatch):
    """drop-in check in the build container: the reference's OWN ``Gan_Dataset`` (training_utils/gan_dataset.py, imported as is;
    its Ceph client and matplotlib are stubbed) reads the jsonl + latents our producer wrote, item for item like our reader."""
    import sys
    import types
    from oracle import ref_shim
    from comat_b200 import gan_data as GD
    pipe, _, _ = _world(monkeypatch)
    index = str(tmp_path / "train_data" / "gan_train_data.jsonl")
    prompts = ["a red apple", "two dogs on a sofa", "a blue car"]
    GD.generate_gan_ground_truth(pipe, prompts, index, batch_size=2, num_inference_steps=1, height=64, width=64,
                                 generator=torch.Generator().manual_seed(2))

    class Client:                                       # aoss_client.client.Client('~/aoss.conf').get(path) -> bytes
        def __init__(self, *_):
            pass

        def get(self, path):
            return open(path, "rb").read()
    for name in ("aoss_client", "aoss_client.client", "matplotlib", "matplotlib.pyplot"):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    sys.modules["aoss_client.client"].Client = Client
    ref_shim.install()
    ref = ref_shim.import_reference("training_utils.gan_dataset")
    args = SimpleNamespace(training_prompts=index)
    theirs, ours = ref.Gan_Dataset(args), GD.Gan_Dataset(args)
    assert len(theirs) == len(ours) == 3
    for i in range(3):
        a, b = theirs[i], ours[i]
        assert a["text"] == b["text"] == prompts[i] and set(a) == set(b) == {"text", "latents"}
        assert a["latents"].dtype == torch.float32 and a["latents"].shape == (4, 8, 8) and torch.equal(a["latents"], b["latents"])
