"""Per-kernel SASS evidence that the library is Blackwell-native (profiling recipe, "What proves a Blackwell-native kernel"):
counts of UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA), legacy HMMA, plus registers /
shared memory per kernel from the cubin's resource usage.  CPU only:

    python tools/sass_evidence.py > profiles/r01_sass_evidence.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "comat_b200", "libcomat_b200.so")
PAT = {"UTC*MMA": re.compile(r"\bUTC[A-Z]*MMA\b"), "LDTM": re.compile(r"\bLDTM\b"), "STTM": re.compile(r"\bSTTM\b"),
       "UTMALDG": re.compile(r"\bUTMALDG\b"), "UTMASTG": re.compile(r"\bUTMASTG\b"), "UBLKCP": re.compile(r"\bUBLKCP\b"),
       "UTMAREDG": re.compile(r"\bUTMAREDG\b"), "HMMA": re.compile(r"\bHMMA\b"), "MUFU.EX2": re.compile(r"\bMUFU\.EX2\b"),
       "RED/ATOM": re.compile(r"\b(RED|ATOM|ATOMG|REDG)\b")}


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None:
            continue
        counts[cur]["instructions"] += 1 if re.match(r"\s+/\*[0-9a-f]{4}\*/", line) else 0
        for k, p in PAT.items():
            if p.search(line):
                counts[cur][k] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage, fn = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
    names = demangle(order)
    short = lambda n: re.sub(r"\(.*", "", names.get(n, n))[:78]
    cols = ["UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU.EX2", "RED/ATOM"]
    print("# SASS evidence - `cuobjdump -sass comat_b200/libcomat_b200.so` (sm_100a), per kernel\n")
    print("`UTC*MMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st (TMEM), `UTMALDG`/`UTMASTG`/`UBLKCP` = TMA tensor load / store / bulk copy,")
    print("`HMMA` = legacy mma.sync (none expected).  REG / static SHARED from `cuobjdump -res-usage` (dynamic shared memory is set at launch).\n")
    print("| kernel | SASS instr | " + " | ".join(cols) + " | REG | static smem B |")
    print("|---|---:|" + "---:|" * (len(cols) + 2))
    tot = collections.Counter()
    for fn in order:
        c = counts[fn]
        tot.update(c)
        r = usage.get(fn, ("", ""))
        print(f"| `{short(fn)}` | {c['instructions']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + f" | {r[0]} | {r[1]} |")
    print(f"\n{len(order)} kernels; totals: " + ", ".join(f"{k} {tot[k]}" for k in cols) + ".")
    if tot["HMMA"]:
        print("\nWARNING: legacy HMMA instructions present.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
