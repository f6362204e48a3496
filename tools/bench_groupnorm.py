"""GroupNorm(+SiLU) forward / backward at the UNet / VAE shapes of BASELINE configs[1]: CUDA-event time per call and the
algorithmic HBM rate (fwd: read x + write y; bwd: read x, dy + write dx).  COMAT_GN=twopass times the two-launch kernels.
  python tools/bench_groupnorm.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import ops

SHAPES = [(8, 4096, 320), (8, 4096, 640), (8, 4096, 960), (8, 1024, 640), (8, 1024, 1280), (8, 1024, 1920), (8, 256, 1280), (8, 256, 2560),
          (8, 64, 1280), (8, 64, 2560), (4, 4096, 320), (4, 4096, 512), (4, 16384, 512), (4, 65536, 256), (4, 262144, 128)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
for n, HW, C in SHAPES:
    x = torch.randn(n, HW, C, device="cuda").half()
    dy = torch.randn(n, HW, C, device="cuda").half()
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    y, mr = ops.groupnorm_fwd(x, g, b, 32, 1e-5, True)
    ops.groupnorm_bwd(x, dy, g, b, mr, 32, True)
    res = {}
    for name, fn, units in (("fwd", lambda: ops.groupnorm_fwd(x, g, b, 32, 1e-5, True), 2), ("bwd", lambda: ops.groupnorm_bwd(x, dy, g, b, mr, 32, True), 3)):
        ts = []
        for _ in range(7):
            flush.zero_()                                  # L2 flush: the tensor is not resident from the previous iteration
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts)[len(ts) // 2]
        res[name] = (t, units * x.numel() * 2 / t / 1e3)
    rows.append(dict(n=n, HW=HW, C=C, fwd_us=res["fwd"][0], fwd_gbs=res["fwd"][1], bwd_us=res["bwd"][0], bwd_gbs=res["bwd"][1]))
    print(f"n={n} HW={HW:6d} C={C:5d}  fwd {res['fwd'][0]:8.1f} us {res['fwd'][1]:7.0f} GB/s   bwd {res['bwd'][0]:8.1f} us {res['bwd'][1]:7.0f} GB/s")
if len(sys.argv) > 1:
    json.dump({"mode": os.environ.get("COMAT_GN", "fused"), "rows": rows}, open(sys.argv[1], "w"), indent=1)
