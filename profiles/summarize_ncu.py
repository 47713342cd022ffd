#!/usr/bin/env python
"""Summarise an ncu report (--set full) into the table committed under profiles/.
    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/ncu_rNN_summary.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 wavefronts %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "warp inst"),
]
stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
print("| kernel | " + " | ".join(c[1] for c in cols) + " | top stalls (warps per issue) |")
print("|---|" + "---|" * (len(cols) + 1))
for r in data:
    name = r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
    vals = []
    for h, _ in cols:
        v = r[idx[h]] if h in idx else "n/a"
        try:
            f = float(v)
            v = f"{f:.4g}"
        except ValueError:
            pass
        vals.append(f"{v} {units[idx[h]]}" if h in idx and units[idx[h]] not in ("", "%") else v)
    st = sorted(((float(r[idx[h]]), h) for h in stall), reverse=True)[:3]
    sts = ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st)
    print(f"| {name} | " + " | ".join(vals) + f" | {sts} |")
