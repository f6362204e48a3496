"""One small-K GEMM shape in a loop (for `ncu --set full`): M=32768 K=320 N=2560, the epilogue/overhead-bound UNet FF projection."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import ops
M, K, N = 32768, 320, 2560
if len(sys.argv) > 3:
    M, K, N = map(int, sys.argv[1:4])
a = torch.randn(M, K, device="cuda").half()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
out = torch.empty(M, N, device="cuda", dtype=torch.float16)
for _ in range(40):
    ops.gemm([a], [w], out=out)
torch.cuda.synchronize()
