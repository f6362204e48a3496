"""Measure (tile width, kernel, split-K) per GEMM shape of the training step on the GPU at hand and write the winners to
comat_b200/gemm_tuning.json (consulted by comat_b200.ops.gemm).

usage: python tools/tune_gemm.py profiles/r01_gemm_shapes_v5.md [--top 60] [--out comat_b200/gemm_tuning.json]
The shape list is the per-shape table bench.py writes with --gemm_shapes (M, N, K segments, taps).  CUDA events around bursts
of 16 back-to-back launches, inputs rotating over several buffers; an entry is kept only when it beats the built-in
heuristic by more than 3 % and reproduces its result."""
import argparse, json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from comat_b200 import ops
from comat_b200 import unet_weights as UW


def parse_shapes(path, top):
    rows = []
    for l in open(path):
        p = [x.strip() for x in l.strip().strip("|").split("|")]
        if len(p) == 10 and p[0].isdigit():
            M, N, segs, taps, ms = int(p[0]), int(p[1]), tuple(int(s) for s in p[2].split("+")), int(p[3]), float(p[7])
            if taps >= 0 and p[5] == "False":
                rows.append(((M, N, segs, taps), ms))
    agg = {}
    for k, ms in rows:
        agg[k] = agg.get(k, 0.0) + ms
    return [k for k, _ in sorted(agg.items(), key=lambda kv: -kv[1])[:top]]


def build(key, dt=torch.float16):
    M, N, segs, taps = key
    g = torch.Generator(device="cuda").manual_seed(M + N)
    if taps == 0:
        nb = max(2, min(6, int(1.5e8 // (M * sum(segs) * 2)) + 1))
        xs = [[torch.randn(M, k, device="cuda", generator=g).to(dt) for k in segs] for _ in range(nb)]
        w = (torch.randn(N, sum(segs), device="cuda", generator=g) / sum(segs) ** 0.5).to(dt)
        kw = dict(b_koff=[0, segs[0]])
        call = lambda i, **o: ops.gemm(xs[i % nb], [w] * len(segs), **kw, **o)
    else:
        n = 8 if math.isqrt(M // 8) ** 2 * 8 == M and (M // 8) & (M // 8 - 1) == 0 else 4
        H = math.isqrt(M // n)
        assert n * H * H == M, key
        nb = max(2, min(6, int(1.5e8 // (M * sum(segs) * 2)) + 1))
        xs = [[torch.randn(n, H, H, k, device="cuda", generator=g).to(dt) for k in segs] for _ in range(nb)]
        w = (torch.randn(N, taps * sum(segs), device="cuda", generator=g) / (taps * sum(segs)) ** 0.5).to(dt)
        tp = UW.TAPS_3x3 if taps == 9 else UW.TAPS_S2D if taps == 4 else [(0, 0)] * taps
        kw = dict(b_koff=[0, segs[0]], conv_taps=tp, c_total=sum(segs))
        call = lambda i, **o: ops.gemm(xs[i % nb], [w] * len(segs), **kw, **o)
    return call


def timeit(fn, reps=5, burst=16):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shapes")
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "comat_b200", "gemm_tuning.json"))
    a = ap.parse_args()
    ops.TUNING = {}
    entries, log = [], []
    for key in parse_shapes(a.shapes, a.top):
        M, N, segs, taps = key
        call = build(key)
        base_out = call(0).float()
        t_def = timeit(lambda i: call(i))
        kb = sum((k + 63) // 64 for k in segs) * max(taps, 1)
        best = (t_def, None)
        for bn in (64, 128, 160, 256):
            if bn > 64 and bn >= 2 * N:
                continue
            tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
            for kern in ("tile", "persist", "pair"):
                if kern == "pair" and bn == 64:
                    continue
                for sk in (1, 2, 3, 4, 6, 8):
                    if sk > 1 and (tiles * sk > 2 * 148 or kb // sk < 4):
                        continue
                    try:
                        t = timeit(lambda i: call(i, force_bn=bn, kernel=kern, split_k=sk), reps=3)
                    except Exception as ex:       # a combination the kernel rejects
                        continue
                    if t < best[0]:
                        best = (t, (bn, kern, sk))
        line = f"{key}: default {t_def:.1f} us"
        if best[1] is not None and best[0] < 0.97 * t_def:
            bn, kern, sk = best[1]
            out = call(0, force_bn=bn, kernel=kern, split_k=sk).float()
            err = float((out - base_out).abs().max() / base_out.abs().max().clamp_min(1e-6))
            line += f" -> {best[0]:.1f} us with bn={bn} {kern} split_k={sk} (max rel diff {err:.1e})"
            if err < 2e-3:
                entries.append({"key": [M, N, list(segs), taps], "bn": bn, "kernel": kern, "split_k": sk,
                                "us": round(best[0], 2), "default_us": round(t_def, 2)})
        print(line, flush=True)
    with open(a.out, "w") as fh:
        json.dump({"device": torch.cuda.get_device_name(0), "how": "tools/tune_gemm.py (CUDA events, bursts of 16 launches)",
                   "entries": entries}, fh, indent=1)
    print(f"wrote {len(entries)} entries to {a.out}")


if __name__ == "__main__":
    main()
