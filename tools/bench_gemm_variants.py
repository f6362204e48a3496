"""Kernel / tile-width variants for the K <= 640 linears (cold inputs: buffers rotate beyond L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import ops
from tools.bench_gemm_res import timeit  # noqa

ops.TUNING = {}
for M, N, K in [(32768, 320, 320), (32768, 960, 320), (32768, 2560, 320), (32768, 1280, 320), (16384, 320, 320), (32768, 320, 64), (8192, 640, 640)]:
    nb = max(3, int(3e8 // (M * (K + 2 * N) * 2)) + 1)
    a = [torch.randn(M, K, device="cuda").half() for _ in range(nb)]
    res = [torch.randn(M, N, device="cuda").half() for _ in range(nb)]
    out = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(nb)]
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    bias = torch.randn(N, device="cuda")
    line = [f"({M},{N},{K})"]
    for kern, bn in [(None, 0), ("persist", 160), ("persist", 128), ("persist", 256), ("pair", 160), ("pair", 128), ("pair", 256), ("tile", 160), ("tile", 128)]:
        try:
            t = timeit(lambda i: ops.gemm([a[i % nb]], [w], bias=bias, residual=res[i % nb], out=out[i % nb], kernel=kern, force_bn=bn), nb)
            line.append(f"{kern or 'auto'}/{bn}: {t:5.1f}")
        except Exception as e:
            line.append(f"{kern}/{bn}: err")
    print("  ".join(line), flush=True)
