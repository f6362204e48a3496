"""Quick correctness + speed check of the CTA-pair GEMM kernel (run under `timeout`: a protocol slip in a 2-CTA kernel is a hang)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import ops

def check(M, N, K, bn):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randint(-2, 3, (M, K), device="cuda", generator=g).half()
    b = torch.randint(-2, 3, (N, K), device="cuda", generator=g).half()
    out = ops.gemm([a], [b], force_bn=bn, kernel="pair")
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    bad = (out.float() - ref).abs() > 0.5
    print(f"pair M{M} N{N} K{K} bn{bn}: mismatches {int(bad.sum())} / {bad.numel()}", flush=True)
    if bad.any():
        idx = bad.nonzero()
        print("  first bad:", idx[:4].tolist(), "rows bad:", sorted(set((idx[:, 0] // 128).tolist()))[:8], "col tiles bad:", sorted(set((idx[:, 1] // 32).tolist()))[:16])
    return not bad.any()

def speed(M, N, K, bn, kern):
    a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for _ in range(3): ops.gemm([a], [w], out=out, force_bn=bn, kernel=kern)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): ops.gemm([a], [w], out=out, force_bn=bn, kernel=kern)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{kern:8s} bn{bn} M{M} N{N} K{K}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.0f} TFLOP/s", flush=True)

ok = True
for cfg in [(256, 256, 64, 256), (256, 256, 256, 256), (512, 512, 512, 256), (384, 320, 192, 160), (5120, 1024, 320, 256), (2048, 1280, 1280, 128)]:
    ok &= check(*cfg)
if ok:
    for (M, N, K) in [(8192, 8192, 8192), (8192, 5120, 640), (32768, 2560, 320), (2048, 10240, 1280), (32768, 1280, 2880)]:
        speed(M, N, K, 256, "pair"); speed(M, N, K, 256, "tile"); speed(M, N, K, 256, "persist")
print("PAIR_OK" if ok else "PAIR_FAIL")
