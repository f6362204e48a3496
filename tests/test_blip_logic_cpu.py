"""CPU: BLIP executor logic (patch embed, ViT blocks, causal / cross attention wiring, post-LN decoder, LM head, label-smoothed CE,
explicit backward) vs the HF module, with the CUDA ops emulated in torch (test infrastructure)."""
import torch

from oracle import comat_ref as R
from oracle import fixtures as FX
from tests import cpu_ops_emulation as EMU


def test_blip_executor_matches_hf(monkeypatch):
    EMU.install_blip(monkeypatch)
    from comat_b200.blip_engine import BlipEngine
    from comat_b200.caption import Blip
    model = R.make_blip(large=False, seed=3)
    g = torch.Generator().manual_seed(5)
    B = 2
    images = torch.rand(B, 3, 200, 200, generator=g)
    ids, mask = FX.blip_token_batch(g, B, 8)
    ref_img = images.clone().requires_grad_(True)
    r_ref = R.blip_score(model, ref_img, ids, mask, 4)
    (-r_ref).backward()
    img = images.clone().requires_grad_(True)
    r = Blip(BlipEngine(model, torch.float32)).score(img, None, input_ids=ids, attention_mask=mask)
    (-r).backward()
    assert abs(float(r) - float(r_ref)) < 1e-4 * abs(float(r_ref)), (float(r), float(r_ref))
    err = ((img.grad - ref_img.grad).norm() / ref_img.grad.norm()).item()
    assert err < 1e-3, err
