#!/usr/bin/env python
"""Turns the raw ncu artefacts a gpurun call brought back (gpurun_out/) into the small text summaries kept in
profiles/:   python profiles/summarize_ncu.py <launches.csv> <prof.ncu-rep> <out_prefix>
 - <out_prefix>_launches.txt : per-kernel launch counts / mean / min / max device time and each kernel's share
 - <out_prefix>_ncu_full.txt : the metrics quoted in DESIGN.md / bench.py's roofline for the dominant kernel
 - <out_prefix>_hot_sass.txt : the 40 hottest SASS instructions by stall samples (needs -lineinfo)"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[kn], []).append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(out, "w") as f:
        f.write(f"# source: {path}  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
        f.write(f"{'kernel':100s} {'n':>5s} {'mean_ns':>10s} {'min_ns':>10s} {'max_ns':>10s} {'share':>7s}\n")
        for k, v in d.items():
            f.write(f"{k[:100]:100s} {len(v):5d} {sum(v)/len(v):10.0f} {min(v):10.0f} {max(v):10.0f} {sum(v)/tot:7.3f}\n")


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    with open(out, "w") as f:
        f.write(f"# source: {rep}  (ncu --set full --clock-control none --import-source on)\n")
        kn = hdr.index("Kernel Name")
        f.write("kernels: " + " | ".join(r[kn][:90] for r in rows[2:]) + "\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"{m:75s} {rows[1][i]:>12s} " + " ".join(r[i] for r in rows[2:]) + "\n")


def hot(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[1]
    iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
    data = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            break
        try:
            data.append((int(r[iS] or 0), int(r[iI] or 0), r))
        except ValueError:
            break
    tot = collections.Counter()
    for s, _, r in data:
        for c in stall:
            tot[hdr[c]] += int(r[c] or 0)
    with open(out, "w") as f:
        f.write(f"# source: {rep}; kernel: {rows[0][1][:100]}\n")
        f.write(f"instructions: {len(data)} SASS, executed {sum(d[1] for d in data)} warp-instr, {sum(d[0] for d in data)} stall samples\n")
        f.write("stall totals: " + ", ".join(f"{k[6:]}={v}" for k, v in tot.most_common(8)) + "\n")
        for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][0])[:40]):
            s, n, r = data[i]
            st = sorted(((hdr[c][6:], int(r[c] or 0)) for c in stall if (r[c] or "0") != "0"), key=lambda kv: -kv[1])[:2]
            f.write(f"{i:5d} samples={s:4d} exec={n:8d} {r[iSrc].strip()[:70]:70s} {st}\n")


if __name__ == "__main__":
    lcsv, rep, prefix = sys.argv[1:4]
    launches(lcsv, prefix + "_launches.txt")
    full(rep, prefix + "_ncu_full.txt")
    hot(rep, prefix + "_hot_sass.txt")
    print("wrote", prefix + "_{launches,ncu_full,hot_sass}.txt")
