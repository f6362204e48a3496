"""The 64x64-latent UNet self-attention shape in a loop (for `ncu --set full`): n=8, 4096 tokens, 8 heads, d=40."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import attention as A
n, L, H, d = 8, 4096, 8, 40
q, k, v = (torch.randn(n, L, H * d, device="cuda").half() for _ in range(3))
for _ in range(6):
    o, _, lse = A.attention_fwd_native(q, k, v, H, need_lse=True)
do = torch.randn_like(o)
for _ in range(2):
    A.attention_bwd_native(q, k, v, o, lse, None, H, do, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5):
    A.attention_fwd_native(q, k, v, H)
e1.record(); torch.cuda.synchronize()
fl = 4.0 * n * H * L * L * d
print("attn fwd ms", e0.elapsed_time(e1) / 5, "TFLOP/s", fl / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12)
e0.record()
for _ in range(5):
    A.attention_bwd_native(q, k, v, o, lse, None, H, do, None)
e1.record(); torch.cuda.synchronize()
print("attn bwd ms", e0.elapsed_time(e1) / 5, "TFLOP/s", 2.5 * fl / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12)
