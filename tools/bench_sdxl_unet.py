"""SDXL UNet (2.57 B parameters, BASELINE configs[4]) forward + backward on the B200 executors at full geometry:
finite outputs / gradients and ms per call.  python tools/bench_sdxl_unet.py [latent=64|128] [batch=2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import synthetic
from comat_b200.modules import EngineUNet

lat = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda")
unet_p = synthetic.build_sdxl_unet(dev, rank=128, seed=42)
mod = EngineUNet(unet_p, torch.float16)
for p in mod.lora_parameters():
    p.grad = torch.zeros_like(p)
mod.direct_lora_grads = True
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, 4, lat, lat, device=dev, generator=g)
ctx = torch.randn(n, 77, 2048, device=dev, generator=g)
added = dict(text_embeds=torch.randn(n, 1280, device=dev, generator=g), time_ids=torch.tensor([[8.0 * lat, 8.0 * lat, 0, 0, 8.0 * lat, 8.0 * lat]] * n, device=dev))
t = torch.tensor(500, device=dev)

def fwd_bwd():
    xr = x.clone().requires_grad_(True)
    eps = mod(xr, t, encoder_hidden_states=ctx, added_cond_kwargs=added)[0]
    eps.float().pow(2).mean().backward()
    mod.finalize_lora_grads()
    return eps, xr.grad

for _ in range(2):
    eps, gx = fwd_bwd()
torch.cuda.synchronize()
assert torch.isfinite(eps).all() and torch.isfinite(gx).all()
gn = sum(float(p.grad.float().pow(2).sum()) for p in mod.lora_parameters()) ** 0.5
assert gn == gn and gn > 0, gn
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(3):
    fwd_bwd()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
with torch.no_grad():
    for _ in range(2):
        mod(x, t, encoder_hidden_states=ctx, added_cond_kwargs=added)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        mod(x, t, encoder_hidden_states=ctx, added_cond_kwargs=added)
    e1.record(); torch.cuda.synchronize()
msf = e0.elapsed_time(e1) / 3
mod.use_graphs = True                       # the same no-grad forward replayed from a CUDA graph (what the rollout's no-grad steps use)
with torch.no_grad():
    for _ in range(2):
        eg = mod(x, t, encoder_hidden_states=ctx, added_cond_kwargs=added)[0]
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        mod(x, t, encoder_hidden_states=ctx, added_cond_kwargs=added)
    e1.record(); torch.cuda.synchronize()
msg = e0.elapsed_time(e1) / 3
assert torch.isfinite(eg).all()
F = {64: 1.589, 128: 6.76}.get(lat)
print(f"SDXL UNet latent {lat}x{lat} n={n}: fwd eager {msf:.1f} ms" + (f" ({n * F / msf * 1e3:.0f} TFLOP/s)" if F else "") +
      f", fwd graph {msg:.1f} ms" + (f" ({n * F / msg * 1e3:.0f} TFLOP/s)" if F else "") +
      f", fwd+bwd(+LoRA grads) {ms:.1f} ms" + (f" ({3 * n * F / ms * 1e3:.0f} TFLOP/s at 3x fwd FLOPs)" if F else "") +
      f", |eps| {float(eps.abs().mean()):.3f}, LoRA grad norm {gn:.3e}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GB")
