"""Micro-benchmark of the tcgen05 GEMM / implicit-GEMM conv kernel at UNet shapes (n = 8 = 2*B CFG batch).
CUDA events on the launching stream; each shape rotates over enough distinct buffers to exceed L2 between repeats."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import ops

def timeit(fn, iters=8, warm=3, burst=16):
    """each sample = `burst` back-to-back launches between two events (the queue stays full, so the sample measures
    GPU time, not the host's launch latency); buffers rotate inside the burst."""
    for _ in range(warm): fn(0)
    torch.cuda.synchronize()
    ts = []
    for i in range(iters):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda._sleep(2_000_000)          # ~1 ms of GPU spin so the burst is queued before it starts executing
        e0.record()
        for j in range(burst):
            fn(i * burst + j)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3 / burst)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

def main():
    dt = torch.float16
    res = []
    n = 8
    for (H, C, Co) in [(64, 320, 320), (32, 640, 640), (16, 1280, 1280), (8, 1280, 1280), (64, 640, 320), (32, 1280, 640)]:
        nb = max(2, int(2.6e8 // (n * H * H * C * 2)) + 1)
        xs = [torch.randn(n, H, H, C, device="cuda").to(dt) for _ in range(nb)]
        w = (torch.randn(Co, 9 * C, device="cuda") / (9 * C) ** 0.5).to(dt)
        bias = torch.randn(Co, device="cuda")
        out = torch.empty(n * H * H, Co, device="cuda", dtype=dt)
        med, best = timeit(lambda i: ops.gemm([xs[i % nb]], [w], bias=bias, conv_taps=ops.TAPS_3x3, out=out))
        fl = 2.0 * n * H * H * C * 9 * Co
        res.append({"op": f"conv3x3 n{n} {H}x{H} {C}->{Co}", "ms": med * 1e3, "tflops_med": fl / med / 1e12, "tflops_best": fl / best / 1e12})
    for (M, K, N) in [(32768, 320, 2560), (32768, 1280, 320), (8192, 640, 5120), (8192, 2560, 640), (2048, 1280, 10240), (32768, 320, 320), (4616, 1024, 4096), (8192, 8192, 8192), (32768, 320, 128), (616, 768, 320), (640, 128, 128), (2048, 1280, 128), (2048, 1280, 1280), (8192, 640, 640)]:
        nb = max(2, int(2.6e8 // (M * K * 2)) + 1)
        xs = [torch.randn(M, K, device="cuda").to(dt) for _ in range(nb)]
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dt)
        out = torch.empty(M, N, device="cuda", dtype=dt)
        med, best = timeit(lambda i: ops.gemm([xs[i % nb]], [w], out=out))
        fl = 2.0 * M * K * N
        res.append({"op": f"linear M{M} K{K} N{N}", "ms": med * 1e3, "tflops_med": fl / med / 1e12, "tflops_best": fl / best / 1e12})
        if N % 256 == 0:
            med, best = timeit(lambda i: ops.gemm([xs[i % nb]], [w], out=out, force_bn=256))
            res.append({"op": f"linear M{M} K{K} N{N} bn256", "ms": med * 1e3, "tflops_med": fl / med / 1e12, "tflops_best": fl / best / 1e12})
        t0 = timeit(lambda i: torch.matmul(xs[i % nb], w.t(), out=out))
        res[-1]["cublas_tflops_med"] = fl / t0[0] / 1e12
    for r in res:
        print(json.dumps(r))

if __name__ == "__main__":
    main()
