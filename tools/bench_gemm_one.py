"""One GEMM / conv shape in a loop (for `ncu --set full`).
  python tools/bench_gemm_one.py M K N            # linear  (default: M=32768 K=320 N=2560, the epilogue-bound UNet FF projection)
  python tools/bench_gemm_one.py conv n H C Cout   # conv3x3 on an (n, H, H, C) NHWC tensor"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import ops
kw = {}
if os.environ.get("BN"):
    kw["force_bn"] = int(os.environ["BN"])
if os.environ.get("KERNEL"):
    kw["kernel"] = os.environ["KERNEL"]
if len(sys.argv) > 1 and sys.argv[1] == "conv":
    n, H, C, Co = map(int, sys.argv[2:6])
    x = torch.randn(n, H, H, C, device="cuda").half()
    w = (torch.randn(Co, 9 * C, device="cuda") / (9 * C) ** 0.5).half()
    out = torch.empty(n * H * H, Co, device="cuda", dtype=torch.float16)
    for _ in range(40):
        ops.gemm([x], [w], conv_taps=ops.TAPS_3x3, out=out, **kw)
else:
    M, K, N = 32768, 320, 2560
    if len(sys.argv) > 3:
        M, K, N = map(int, sys.argv[1:4])
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for _ in range(40):
        ops.gemm([a], [w], out=out, **kw)
torch.cuda.synchronize()
