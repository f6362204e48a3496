"""CPU: the full-size parity harness (tests/parity_world.py) itself, at the tiny geometry on emulated ops - product trainer vs
oracle for the generator step (all five losses + LoRA gradients) and the discriminator step (loss, D LoRA + head gradients)."""
import torch

from tests import cpu_ops_emulation as EMU
from tests import parity_world as PW
from tests.test_trainer_logic_cpu import _emulate_cuda_only


def test_harness_generator_and_discriminator_steps(monkeypatch):
    _emulate_cuda_only(monkeypatch)
    EMU.install_blip(monkeypatch)
    w = PW.sd15_world("cpu", torch.float32, tiny=True, B=2, S=3, K=2, res=128, rank=4, layers=["up_8", "up_16"])
    r = PW.g_step_compare(w)
    for k in ("Blip", "G_loss", "token_loss", "pixel_loss", "loss"):
        assert r[k] < 2e-4, (k, r)
    assert r["image"] < 1e-4 and r["grad_cos"] > 0.9999 and abs(r["grad_norm_ratio"] - 1) < 1e-3, r
    d = PW.d_step_compare(w)
    assert d["D_loss"] < 1e-5 and d["grad_cos"] > 0.9999 and abs(d["grad_norm_ratio"] - 1) < 1e-3 and d["head_grad_rel"] < 1e-4, d
