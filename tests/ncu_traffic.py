#!/usr/bin/env python
"""Extract per-launch DRAM traffic of the pass kernels from an `ncu --set full` capture of
tests/profile_run.py (one bench.py step) into profiles/ncu_traffic.json, which bench.py reads for
`roofline.traffic` (so the figure follows the committed capture instead of a literal in bench.py).

    python tests/ncu_traffic.py gpurun_out/prof_r02.ncu-rep cfg3/f32/256x150x256/b2 [profiles/ncu_traffic.json]

Kernels are matched to bench.py's pass labels by launch order within ONE iteration of the step:
forward pass 0 (one launch per segment / block), forward passes 1..nd-1, pair product (= inverse pass nd-1),
inverse passes nd-2..0.
"""
import csv
import io
import json
import os
import subprocess
import sys


def main():
    rep, key = sys.argv[1], sys.argv[2]
    dst = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                              "profiles", "ncu_traffic.json")
    dims = [int(v) for v in key.split("/")[2].split("x")]
    nd = len(dims)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                          "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

    def val(r, h):
        return float(r[ix[h]]) * scale.get(units[ix[h]], 1.0)
    passes = [r for r in data if any(s in r[ix["Kernel Name"]] for s in ("xcb_col_pass", "xcb_row_pass", "xcb_pair_product"))]
    # forward: everything before the first pair-product / PRO row pass launch
    first_inv = next(i for i, r in enumerate(passes)
                     if "pair_product" in r[ix["Kernel Name"]] or ", true>" in r[ix["Kernel Name"]] or ", 1>(" in r[ix["Kernel Name"]] and "row_pass" in r[ix["Kernel Name"]])
    fwd, inv = passes[:first_inv], passes[first_inv:first_inv + nd]
    n_fwd0 = len(fwd) - (nd - 1)
    table = {}
    # the first launch of a kind in a cold process also reads its twiddle tables etc.: take the LAST of equal launches
    table[f"fwd_pass0_N{dims[0]}"] = sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in fwd[:n_fwd0]) / n_fwd0
    for a in range(1, nd):
        r = fwd[n_fwd0 + a - 1]
        table[f"fwd_pass{a}_N{dims[a]}"] = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    for k, r in enumerate(inv):
        a = nd - 1 - k
        table[f"inv_pass{a}_N{dims[a]}"] = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    tab = {}
    if os.path.exists(dst):
        with open(dst) as f:
            tab = json.load(f)
    tab[key] = table
    tab["_source"] = f"ncu --set full capture {os.path.basename(rep)} of tests/profile_run.py (dram__bytes_read.sum + dram__bytes_write.sum per launch; fwd_pass0 = mean over its launches)"
    with open(dst, "w") as f:
        json.dump(tab, f, indent=1)
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
