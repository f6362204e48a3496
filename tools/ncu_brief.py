"""Key metrics of every kernel in an .ncu-rep as a markdown table (the numbers DESIGN.md / profiles/README.md quote).
usage: python tools/ncu_brief.py file.ncu-rep [file2.ncu-rep ...] > profiles/rNN_xxx_ncu.md"""
import csv, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
        ("smsp__sass_inst_executed_op_local_ld.sum", "local loads")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"### {rep.split('/')[-1]}  (`ncu --set full --clock-control none`)\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"**`{name[:110]}`**\n\n| metric | value |\n|---|---:|")
        for k, label in KEYS:
            if k in hdr:
                v = r[hdr.index(k)]
                try:
                    v = f"{float(v.replace(',', '')):.2f}"
                except ValueError:
                    pass
                print(f"| {label} | {v} {units[hdr.index(k)]} |")
        stalls = sorted(((float(r[i] or 0), h.split("issue_stalled_")[1].split("_per_issue")[0]) for i, h in enumerate(hdr)
                         if "issue_stalled" in h and "per_issue_active" in h), reverse=True)[:5]
        print("| top stalls (per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls) + " |\n")
