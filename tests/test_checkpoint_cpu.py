"""CPU: checkpoint wire format - tensor names / shapes against the reference's own ``unet_lora_state_dict``
(tests/golden/lora_state_dict_keys.json, produced by running that function verbatim in oracle/pin_against_reference.py), file
layout of ``checkpoint-<n>/``, save -> load round trips, resume semantics of training_script.py:156-205."""
import json
import os
import random

import pytest
import torch

from tests.conftest import GOLDEN


def _keys(sd):
    return [[k, list(v.shape)] for k, v in sd.items()]


@pytest.mark.parametrize("name", ["sd15_r128", "sdxl_r128", "tiny_r4"])
def test_lora_state_dict_names_match_reference_function(name):
    from comat_b200 import checkpoint as CK, containers as Cn
    gold = json.load(open(os.path.join(GOLDEN, "lora_state_dict_keys.json")))[name]
    cfg = {"sd15_r128": {}, "sdxl_r128": Cn.SDXL_UNET,
           "tiny_r4": dict(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=64)}[name]
    with torch.device("meta"):
        unet = Cn.UNet2DConditionModel(**cfg)
        unet.install_lora(128 if name != "tiny_r4" else 4)
    sd = CK.unet_lora_state_dict(unet)
    assert len(sd) == gold["n"] and sum(v.numel() for v in sd.values()) == gold["numel"]
    assert _keys(sd) == gold["keys"]
    assert all(k.startswith("unet.") and (k.endswith(".lora.down.weight") or k.endswith(".lora.up.weight")) for k in sd)


@pytest.mark.parametrize("prefix", [True, False])
def test_lora_file_round_trip(tmp_path, prefix):
    from safetensors import safe_open
    from comat_b200 import checkpoint as CK, synthetic
    u1, _ = synthetic.build_sd15("cpu", torch.float32, rank=4, seed=1, tiny=True, lora_up_std=0.05)
    u2, _ = synthetic.build_sd15("cpu", torch.float32, rank=4, seed=2, tiny=True, lora_up_std=0.05)
    path = CK.save_lora_weights(str(tmp_path / "ck"), CK.unet_lora_state_dict(u1), diffusers_prefix=prefix)
    assert os.path.basename(path) == "pytorch_lora_weights.safetensors"
    with safe_open(path, "pt") as f:
        assert f.metadata() == {"format": "pt"}
        names = list(f.keys())
    lead = "unet.unet." if prefix else "unet."
    assert all(n.startswith(lead) and not n.startswith(lead + "unet.") for n in names) and len(names) == 256
    state = CK.lora_state_dict(str(tmp_path / "ck"))
    assert all(not k.startswith("unet.") for k in state)
    assert CK.load_lora_into_unet(state, u2) == 256
    for (k1, a), (k2, b) in zip(CK.unet_lora_state_dict(u1).items(), CK.unet_lora_state_dict(u2).items()):
        assert k1 == k2 and torch.equal(a, b) and b.dtype == torch.float32
    # errors: a module the UNet does not have, a wrong shape, half an adapter
    bad = dict(state); bad["nowhere.to_q.lora.down.weight"] = torch.zeros(4, 8)
    with pytest.raises(KeyError):
        CK.load_lora_into_unet(bad, u2)
    k0 = next(iter(state))
    with pytest.raises(ValueError):
        CK.load_lora_into_unet({**state, k0: torch.zeros(3, 3)}, u2)
    with pytest.raises(KeyError):
        CK.load_lora_into_unet({k: v for k, v in state.items() if k != k0}, u2)


def test_load_installs_missing_adapters_at_the_stored_rank(tmp_path):
    from comat_b200 import checkpoint as CK, containers as Cn, synthetic
    u1, _ = synthetic.build_sd15("cpu", torch.float32, rank=4, seed=1, tiny=True, lora_up_std=0.05)
    CK.save_lora_weights(str(tmp_path), CK.unet_lora_state_dict(u1))
    torch.manual_seed(0)
    bare = Cn.UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=64)
    assert len(CK.unet_lora_state_dict(bare)) == 0
    CK.load_lora_into_unet(CK.lora_state_dict(str(tmp_path)), bare)
    sd = CK.unet_lora_state_dict(bare)
    assert _keys(sd) == _keys(CK.unet_lora_state_dict(u1))
    assert all(torch.equal(a, b) for a, b in zip(sd.values(), CK.unet_lora_state_dict(u1).values()))


def _tiny_trainer(seed, monkeypatch):
    from tests.test_trainer_logic_cpu import _emulate_cuda_only
    _emulate_cuda_only(monkeypatch)
    from comat_b200 import synthetic
    from comat_b200.gan import D_sd
    from comat_b200.modules import EngineUNet, EngineVAE
    from comat_b200.pipelines import TrainableSDPipeline
    from comat_b200.trainer import CoMatTrainer
    unet, vae = synthetic.build_sd15("cpu", torch.float32, rank=4, seed=seed, tiny=True, lora_up_std=0.05)
    d, _ = synthetic.build_sd15("cpu", torch.float32, rank=4, seed=seed + 100, tiny=True, lora_up_std=0.02)
    args = synthetic.default_args(pretrain_model_name="sd_1_5", gan_loss=True, seed=seed)
    torch.manual_seed(seed)
    pipe = TrainableSDPipeline(EngineVAE(vae, torch.float32), EngineUNet(unet, torch.float32))
    return CoMatTrainer(args, pipe, None, D_sd(EngineUNet(d, torch.float32)), rng=random.Random(seed))


def test_trainer_checkpoint_layout_and_exact_resume(tmp_path, monkeypatch):
    from comat_b200 import checkpoint as CK
    a = _tiny_trainer(1, monkeypatch)
    g = torch.Generator().manual_seed(3)
    for opt in (a.optimizer, a.D_optimizer):                   # a state a few steps into training
        opt.m.copy_(torch.randn(opt.n, generator=g)); opt.v.copy_(torch.rand(opt.n, generator=g)); opt.step_count = 17
        opt.flat.add_(torch.randn(opt.n, generator=g) * 0.01)
    a.global_step = 10
    a.rng.random(); torch.rand(3)
    out = str(tmp_path / "run")
    CK.save_checkpoint(a, out, global_step=9)
    path = CK.save_checkpoint(a, out)
    assert path.endswith("checkpoint-10") and CK.latest_checkpoint(out) == path          # numeric, not lexicographic, order
    assert sorted(os.listdir(path)) == ["D_sd", "pytorch_lora_weights.safetensors", "trainer_state.pt"]
    assert sorted(os.listdir(os.path.join(path, "D_sd"))) == ["mlp.pt", "pytorch_lora_weights.safetensors"]
    assert sorted(torch.load(os.path.join(path, "D_sd", "mlp.pt")).keys()) == ["0.bias", "0.weight"]
    nxt_a = (a.rng.random(), torch.rand(2))
    # resume 'latest' into a differently-initialised trainer: generator, discriminator, head, moments, counters, RNG streams
    b = _tiny_trainer(2, monkeypatch)
    assert not torch.equal(a.optimizer.flat, b.optimizer.flat)
    assert CK.load_checkpoint(b, out, "latest") == 10 and b.global_step == 10
    for x, y in ((a.optimizer, b.optimizer), (a.D_optimizer, b.D_optimizer)):
        assert torch.equal(x.flat, y.flat) and torch.equal(x.m, y.m) and torch.equal(x.v, y.v) and y.step_count == 17
    for p, q in zip(a.D.mlp.parameters(), b.D.mlp.parameters()):
        assert torch.equal(p, q) and q.dtype == torch.float32
    nxt_b = (b.rng.random(), torch.rand(2))
    assert nxt_a[0] == nxt_b[0] and torch.equal(nxt_a[1], nxt_b[1])
    # the executors' 16-bit operand images follow the restored masters
    l = b.pipeline.unet.engine.loras[0]
    assert torch.equal(l.down16.float(), l.down.detach().to(l.down16.dtype).float())
    # an explicit checkpoint path restores the generator only (training_script.py:191: D only on 'latest')
    c = _tiny_trainer(3, monkeypatch)
    d_before = c.D_optimizer.flat.clone()
    assert CK.load_checkpoint(c, os.path.join(out, "checkpoint-9"), resume="explicit") == 9
    assert torch.equal(c.optimizer.flat, a.optimizer.flat) and torch.equal(c.D_optimizer.flat, d_before)
    assert CK.load_checkpoint(c, str(tmp_path / "empty"), "latest") is None


@pytest.mark.parametrize("name", ["sd15_r128", "sdxl_r128", "tiny_r4"])
def test_lora_install_matches_reference_functions(name):
    """row a14: the reference's own ``set_pipeline_trainable_module`` + ``get_trainable_parameters`` (training_utils/pipeline.py:84-187,
    run verbatim for tests/golden/lora_trainable_params.json) and the product's ``install_lora`` / executor agree on which projections
    carry LoRA, the rank, fp32 masters, and the ORDER of the trainable-parameter list (= layout of the optimiser's flat buffer)."""
    from comat_b200 import containers as Cn
    gold = json.load(open(os.path.join(GOLDEN, "lora_trainable_params.json")))[name]
    cfg = {"sd15_r128": {}, "sdxl_r128": Cn.SDXL_UNET,
           "tiny_r4": dict(block_out_channels=(64, 128, 256, 256), heads=4, cross_attention_dim=64)}[name]
    with torch.device("meta"):
        unet = Cn.UNet2DConditionModel(**cfg)
        params = unet.install_lora(128 if name != "tiny_r4" else 4)
    names = {id(p): n for n, p in unet.named_parameters()}
    assert [[names[id(p)], list(p.shape)] for p in params] == gold["params"]
    assert len(params) == gold["n"] and sum(p.numel() for p in params) == gold["numel"] and all(p.dtype == torch.float32 for p in params)


def test_executor_lora_parameter_order_matches_reference_order():
    from comat_b200 import synthetic
    from comat_b200.modules import EngineUNet
    gold = json.load(open(os.path.join(GOLDEN, "lora_trainable_params.json")))["tiny_r4"]
    unet, _ = synthetic.build_sd15("cpu", torch.float32, rank=4, seed=1, tiny=True)
    eng = EngineUNet(unet, torch.float32)
    names = {id(p): n for n, p in unet.named_parameters()}
    got = [[names[id(p)], list(p.shape)] for p in eng.lora_parameters()]
    assert sorted(map(str, got)) == sorted(map(str, gold["params"]))               # same set of tensors ...
    assert got == gold["params"]                                                    # ... in the reference's order
