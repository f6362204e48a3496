"""Micro-benchmark of the attention-map loss kernels at BASELINE config-2 geometry (SD1.5, B=4, 2 attrcon timesteps).
CUDA events on the launching stream; inputs (319 MB) exceed L2 (126 MB) so no explicit flush is needed."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import attn_loss
from oracle import fixtures as FX

def main(B=4, n_t=2, iters=20):
    torch.manual_seed(0)
    H = 8
    spec = {"mid_8": 1, "up_16": 3, "up_32": 3, "up_64": 3}
    layers = list(spec)
    attn = {}
    for t in range(n_t):
        attn[str(951 - 200 * t)] = {k: [torch.softmax(torch.randn(B * H, int(k.split("_")[1]) ** 2, 77, device="cuda") * 2, -1)
                       .reshape(B * H, int(k.split("_")[1]), int(k.split("_")[1]), 77).requires_grad_(True)
                       for _ in range(n)] for k, n in spec.items()}
    g = torch.Generator().manual_seed(1)
    masks = [[FX.random_mask(g, 512).cuda() for _ in range(3)] for _ in range(B)]
    words = [[[3, 4], [10], [12, 13, 14]] for _ in range(B)]
    groups = [{"res": int(l.split("_")[1]), "maps": attn[t][l]} for t in attn for l in layers]
    plan = attn_loss.AttnMapLossPlan(groups, words, masks, torch.device("cuda"))
    flat = [m for gg in groups for m in gg["maps"]]
    byts = plan.algorithmic_bytes
    def fwd():
        return attn_loss.fused_attnmap_loss(plan, flat)
    out = fwd(); out.sum().backward()
    torch.cuda.synchronize()
    res = {"bytes": byts, "n_work": plan.n_work}
    for name in ("fwd", "bwd"):
        ts = []
        for _ in range(iters):
            if name == "fwd":
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record(); out = fwd(); e1.record()
            else:
                out = fwd()
                g2 = torch.ones(2, device="cuda")
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record(); torch.autograd.grad(out, flat, grad_outputs=g2); e1.record()   # as in the step: the map gradients go straight to their consumer (no leaf accumulation)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        ts.sort()
        med = ts[len(ts) // 2]
        res[name] = {"ms_med": med * 1e3, "ms_min": ts[0] * 1e3, "GBps_med": byts / med / 1e9, "GBps_best": byts / ts[0] / 1e9}
    # kernel-only timing: the C-ABI calls on pre-allocated buffers, bursts of 10 launches between two events
    import ctypes as C
    from comat_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda")
    loss2 = torch.empty(2, device=dev); state = torch.empty(plan.state_floats, device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    grads = [torch.empty_like(m) for m in plan.maps]
    gp = torch.tensor([g_.data_ptr() for g_ in grads], dtype=torch.int64, device=dev)
    g2 = torch.ones(2, device=dev)
    st = _lib.stream_ptr()
    def kfwd(): _lib.check(L.comat_attnmap_loss_fwd(C.byref(plan.c), _lib.ptr(loss2), _lib.ptr(state), plan.state_floats, _lib.ptr(counter), st))
    def kbwd(): _lib.check(L.comat_attnmap_loss_bwd(C.byref(plan.c), _lib.ptr(g2), _lib.ptr(state), _lib.ptr(gp), st))
    for name, fn in (("kernel_fwd", kfwd), ("kernel_bwd", kbwd)):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            torch.cuda._sleep(2_000_000)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-4)
        ts.sort()
        res[name] = {"ms_med": ts[len(ts) // 2] * 1e3, "GBps_med": byts / ts[len(ts) // 2] / 1e9, "GBps_best": byts / ts[0] / 1e9,
                     "note": "10 back-to-back launches per sample; maps (319 MB) exceed L2 (126 MB)"}
    print(json.dumps(res))

if __name__ == "__main__":
    main()
