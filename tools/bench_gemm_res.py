"""K <= 1280 linears of the UNet transformer blocks with and without a residual operand (out-projections, feed-forward outputs and
every gradient accumulation carry one).  Inputs rotate over buffers that together exceed the 126 MB L2; CUDA events around bursts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import ops

def timeit(fn, nb, reps=7, burst=16):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    ts = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda._sleep(1_000_000)
        e0.record()
        for j in range(burst):
            fn(r * burst + j)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / burst)
    ts.sort()
    return ts[len(ts) // 2]

if __name__ == "__main__":
  for M, N, K in [(32768, 320, 320), (8192, 640, 640), (2048, 1280, 1280), (32768, 320, 1280), (8192, 640, 2560), (16384, 320, 320)]:
      nb = max(3, int(3e8 // (M * (K + 2 * N) * 2)) + 1)
      a = [torch.randn(M, K, device="cuda").half() for _ in range(nb)]
      res = [torch.randn(M, N, device="cuda").half() for _ in range(nb)]
      out = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(nb)]
      w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
      bias = torch.randn(N, device="cuda")
      t0 = timeit(lambda i: ops.gemm([a[i % nb]], [w], bias=bias, out=out[i % nb]), nb)
      t1 = timeit(lambda i: ops.gemm([a[i % nb]], [w], bias=bias, residual=res[i % nb], out=out[i % nb]), nb)
      fl = 2.0 * M * N * K
      by0, by1 = 2.0 * (M * K + M * N + N * K), 2.0 * (M * K + 2 * M * N + N * K)
      print(f"({M},{N},{K}) x{nb} buffers: plain {t0:6.1f} us = {fl / t0 / 1e6:5.0f} TFLOP/s, {by0 / t0 / 1e3:5.0f} GB/s | +residual {t1:6.1f} us = {fl / t1 / 1e6:5.0f} TFLOP/s, {by1 / t1 / 1e3:5.0f} GB/s")
