#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw + source pages) into a short text: key counters, stall mix, and a
per-function breakdown of samples using the -lineinfo file/line of each SASS instruction is not
available in CSV, so functions are recovered from the cuobjdump symbol ranges of the .so."""
import csv, subprocess, sys, collections, re, os

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_inst_executed_op_local_ld.sum',
        'smsp__sass_inst_executed_op_local_st.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__cycles_active.avg']
lines = [f"# {os.path.basename(rep)}  {note}"]
for h, u, v in zip(hdr, units, vals):
    if h in keep or (h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')):
        lines.append(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = rows[2:]
tot = sum(int(r[isamp] or 0) for r in data)
lines.append(f"\n# SASS windows of 400 instructions with >1% of {tot} samples: start, %samples, inst executed, top stalls, top opcodes")
for b in range(0, len(data), 400):
    win = data[b:b + 400]
    s = sum(int(r[isamp] or 0) for r in win)
    if s < 0.01 * tot:
        continue
    ex = sum(int(r[iex] or 0) for r in win)
    st = collections.Counter()
    for r in win:
        for i in stall_cols:
            st[h[i]] += int(r[i] or 0)
    ops = collections.Counter((r[isrc].split()[1] if r[isrc].startswith('@') else r[isrc].split()[0]) for r in win if r[isrc])
    lines.append(f"{b:6d} {100 * s / tot:5.1f}% exec {ex:.2e}  " + " ".join(f"{k[6:]}:{100 * v / max(1, s):.0f}%" for k, v in st.most_common(4))
                 + "  | " + " ".join(f"{k}:{v}" for k, v in ops.most_common(5)))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
