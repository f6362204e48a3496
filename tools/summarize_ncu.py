"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share.
usage: python tools/summarize_ncu.py gpurun_out/launches.csv > profiles/rNN_launches_summary.md"""
import csv, collections, re, sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"<.*", "", name) if name.startswith("void at::") or "at::native" in name else name
    t = float(r[iv].replace(",", ""))
    c, s = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, s + t)
tot = sum(s for _, s in agg.values())
print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for name, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{name[:90]}` | {c} | {s / 1e6:.2f} | {100 * s / tot:.1f} % |")
print(f"| **total** | {sum(c for c, _ in agg.values())} | {tot / 1e6:.2f} | 100 % |")
