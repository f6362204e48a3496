"""The kernel BASELINE's north star names, in a loop (for `ncu --set full`): UNet cross-attention at the 64x64 level with the fp32
probability export the AttentionStore clones (tc_attn_utils.py:126-145) - cond half of a B=4 attrcon step: n=4, 4096 queries,
77 text tokens, 8 heads, d=40 - plus its backward with the attention-map loss's dP."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import attention as A
n, Lq, Lk, H, d = 4, 4096, 77, 8, 40
q = torch.randn(n, Lq, H * d, device="cuda").half()
k, v = (torch.randn(n, Lk, H * d, device="cuda").half() for _ in range(2))
for _ in range(6):
    o, probs, lse = A.attention_fwd_native(q, k, v, H, export_probs=True, need_lse=True)
do, dp = torch.randn_like(o), torch.randn_like(probs) * 1e-3
for _ in range(3):
    A.attention_bwd_native(q, k, v, o, lse, probs, H, do, dp)
torch.cuda.synchronize()
def t(fn, reps=20):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
tf = t(lambda: A.attention_fwd_native(q, k, v, H, export_probs=True, need_lse=True))
tb = t(lambda: A.attention_bwd_native(q, k, v, o, lse, probs, H, do, dp))
pb = probs.numel() * 4
print(f"xattn fwd+P export {tf:.1f} us: P write {pb / 1e6:.1f} MB -> {pb / tf / 1e3:.0f} GB/s; {4.0 * n * H * Lq * Lk * d / tf / 1e6:.1f} TFLOP/s")
print(f"xattn bwd (dP in)  {tb:.1f} us: P + dP read {2 * pb / 1e6:.1f} MB -> {2 * pb / tb / 1e3:.0f} GB/s")
