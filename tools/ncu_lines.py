#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals from an ncu report captured with --import-source on (-lineinfo build):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg = collections.OrderedDict()
fname, hdr, func = None, None, None
first_func = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1]
        first_func = first_func or func
    elif r[0] == "Line No":
        hdr = r
    elif hdr and func == first_func and r[0].isdigit():
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            ex, sm = int(r[iI] or 0), int(r[iS] or 0)
        except ValueError:
            continue
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += ex; a[1] += sm
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"kernel: {first_func}\ntotal warp-instructions {tot_i}, stall samples {tot_s}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:4d} inst={a[0]:>11d} ({100*a[0]/max(tot_i,1):4.1f}%) samples={a[1]:>6d} ({100*a[1]/max(tot_s,1):4.1f}%)  {a[2]}")
