#!/usr/bin/env python
"""Runs ON THE GPU BOX right after an ncu capture: boils a .ncu-rep (tens of MB with --import-source) down to one small text file —
key metrics per captured launch, pipe / opcode shares (the integer / LOP3 issue utilisation BASELINE.json's north_star asks for), the
source lines and SASS instructions with the most issued instructions / stall samples — so that gpurun_out stays under its 64 MiB cap.
    python tools/ncu_summarize_box.py <rep> <out.txt> [kernel launch index for the source pages, default 0]"""
import collections
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum",
    "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_uniform.sum", "smsp__inst_executed_pipe_adu.sum",
    "smsp__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "smsp__sass_inst_executed_op_global_st.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
]
with open(out, "w") as f:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    f.write(f"# source: {rep} (ncu --set full --clock-control none --import-source on)\n")
    kn = hdr.index("Kernel Name")
    f.write("launches: " + " | ".join(r[kn][:70] for r in rows[2:]) + "\n")
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            f.write(f"{m:92s} {rows[1][i]:>10s} " + " ".join(r[i] for r in rows[2:]) + "\n")
    # stall reasons of the first launch (warp state samples)
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            vals = [r[i] for r in rows[2:]]
            if any(v not in ("0", "") for v in vals):
                f.write(f"{h:92s} {'':>10s} " + " ".join(vals) + "\n")
    # source page: opcode histogram + hottest lines / instructions (first captured launch)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    ops = collections.Counter(); lines = collections.OrderedDict(); sass = []
    fname = hdr2 = func = first = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            func = r[1]; first = first or func
        elif r[0] == "Line No":
            hdr2 = r
        elif hdr2 and func == first and r[0].isdigit():
            try:
                iI, iS = hdr2.index("Instructions Executed"), hdr2.index("# Samples")
                ex, sm = int(r[iI] or 0), int(r[iS] or 0)
            except (ValueError, IndexError):
                continue
            a = lines.setdefault((fname, int(r[0])), [0, 0, r[1].strip()[:100]])
            a[0] += ex; a[1] += sm
    # SASS view for opcodes
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h3 = None
    for r in rows:
        if r and r[0] == "Address":
            h3 = r; break
    if h3 is None and len(rows) > 1:
        h3 = rows[1]
    if h3 and "Source" in h3 and "Instructions Executed" in h3:
        iSrc, iI, iS = h3.index("Source"), h3.index("Instructions Executed"), h3.index("# Samples")
        for r in rows[rows.index(h3) + 1:]:
            if len(r) != len(h3):
                break
            try:
                ex, sm = int(r[iI] or 0), int(r[iS] or 0)
            except ValueError:
                break
            ins = r[iSrc].strip()
            op = ins.split()[1] if ins.startswith("@") and len(ins.split()) > 1 else (ins.split()[0] if ins else "?")
            ops[op.split(".")[0]] += ex
            sass.append((sm, ex, ins[:90]))
    tot = sum(ops.values()) or 1
    f.write(f"\nopcode shares of the first captured launch ({tot} warp-instructions): " +
            ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(18)) + "\n")
    ti = sum(a[0] for a in lines.values()) or 1; ts = sum(a[1] for a in lines.values()) or 1
    f.write(f"\nsource lines by instructions executed (all launches of the kernel in the report: {ti} warp-instructions, {ts} stall samples)\n")
    for (fn, ln), a in sorted(lines.items(), key=lambda kv: -kv[1][0])[:36]:
        f.write(f"{fn}:{ln:4d} inst={a[0]:>11d} ({100 * a[0] / ti:4.1f}%) samples={a[1]:>6d} ({100 * a[1] / ts:4.1f}%)  {a[2]}\n")
    f.write("\nhottest SASS by stall samples (first captured launch)\n")
    for sm, ex, ins in sorted(sass, key=lambda t: -t[0])[:24]:
        f.write(f"samples={sm:5d} exec={ex:9d} {ins}\n")
