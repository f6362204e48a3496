"""GPU parity: native BLIP executor (ViT + text decoder + label-smoothed CE on the tcgen05 kernels) vs the HF
BlipForConditionalGeneration module (fp32) on the same random-init weights — loss and gradient w.r.t. the image."""
import pytest
import torch

from oracle import comat_ref as R
from oracle import fixtures as FX

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("large,layers", [(False, None), (True, (2, 2))])
def test_blip_engine_loss_and_image_grad(large, layers):
    from comat_b200.blip_engine import BlipEngine
    from comat_b200.caption import Blip
    model = R.make_blip(large=large, seed=3, layers=layers).cuda()
    g = torch.Generator().manual_seed(5)
    B = 3
    images = torch.rand(B, 3, 254, 254, generator=g).cuda()
    ids, mask = FX.blip_token_batch(g, B, 12)
    ids, mask = ids.cuda(), mask.cuda()
    ref_img = images.clone().requires_grad_(True)
    r_ref = R.blip_score(model, ref_img, ids, mask, 4)
    (-r_ref).backward()
    eng = BlipEngine(model, torch.float16)
    img = images.clone().requires_grad_(True)
    r = Blip(eng).score(img, None, input_ids=ids, attention_mask=mask)
    (-r).backward()
    assert abs(float(r) - float(r_ref)) < 1e-3 * abs(float(r_ref)), (float(r), float(r_ref))     # north-star tolerance
    cos = float((img.grad.double() * ref_img.grad.double()).sum() / (img.grad.double().norm() * ref_img.grad.double().norm()))
    assert cos > 0.97 and abs(float(img.grad.norm() / ref_img.grad.norm()) - 1) < 0.1, (cos, float(img.grad.norm() / ref_img.grad.norm()))
