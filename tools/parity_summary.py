"""profiles/r02_parity_fullsize.md from the `[measured]` lines of the device-test logs (profiles/r02_gpu_tests_run*.log; the last
occurrence of each line wins).  usage: python tools/parity_summary.py"""
import glob, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
last = {}
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_gpu_tests_run*.log")), key=lambda p: int(re.findall(r"run(\d+)", p)[0])):
    for line in open(f, errors="ignore"):
        for m in re.finditer(r"\[measured\] ([^:]+): (.*?)(?=\[measured\]|$)", line.strip()):
            last[m.group(1).strip()] = (m.group(2).strip().rstrip("."), os.path.basename(f))
out = ["# Round 2 - device parity at BASELINE geometry (B200, fp16 product vs fp32 oracle on the same GPU, TF32 off)\n",
       "Source: `tests/test_parity_fullsize_gpu.py` (+ `tests/parity_world.py`) and the other `-m gpu` tests that print `[measured]` lines, run "
       "under `gpurun`; logs: `profiles/r02_gpu_tests_run*.log` (the last run of each line is kept here).",
       "Frozen weights are rounded to fp16 on both sides (the reference runs `pipeline.unet.to(weight_dtype)`, training_utils/pipeline.py:60-65); "
       "the oracle then computes in fp32.\n",
       "North-star tolerance: 1e-3 relative on the concept-matching loss (`Blip`) and the per-token attention loss (`token_loss`).\n",
       "| test | measured | log |\n|---|---|---|"]
out += [f"| {k} | {v} | {f} |" for k, (v, f) in last.items()]
open(os.path.join(ROOT, "profiles", "r02_parity_fullsize.md"), "w").write("\n".join(out) + "\n")
print(len(last), "lines")
