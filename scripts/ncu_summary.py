#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw + source pages) into a short text: key counters, stall mix, and a
per-function breakdown of samples using the -lineinfo file/line of each SASS instruction is not
available in CSV, so functions are recovered from the cuobjdump symbol ranges of the .so."""
import csv, subprocess, sys, collections, re, os

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
json_out = sys.argv[4] if len(sys.argv) > 4 else None      # e.g. profiles/traffic.json: the figures bench.py quotes
chains = int(sys.argv[5]) if len(sys.argv) > 5 else 148
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
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__cycles_active.avg',
        'sm__cycles_elapsed.avg', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_src_fp64.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed']
lines = [f"# {os.path.basename(rep)}  {note}"]
for h, u, v in zip(hdr, units, vals):
    if h in keep or (h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')):
        lines.append(f"{h} [{u}] = {v}")
if json_out:
    import json
    m = {h_: float(v_.replace(",", "")) for h_, v_ in zip(hdr, vals) if h_ in keep and v_ not in ("", "no data")}
    u_ = dict(zip(hdr, units))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd = m['dram__bytes_read.sum'] * scale.get(u_['dram__bytes_read.sum'], 1.0)
    wr = m['dram__bytes_write.sum'] * scale.get(u_['dram__bytes_write.sum'], 1.0)
    cyc = m['sm__cycles_elapsed.avg']
    dfma = m['smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed'] * cyc
    dadd = m['smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed'] * cyc
    dmul = m['smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed'] * cyc
    tens = m['sm__ops_path_tensor_src_fp64.sum']               # flop on the fp64 tensor path (DMMA), 128 per cycle per SM at peak
    json.dump({
        "source": f"{os.path.basename(out)} (ncu --set full, one k_chain launch, {chains} chains, L=32, 24 layers, nstep=10)",
        "chains": chains, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_chain_traj": (rd + wr) / chains,
        "smem_wavefronts_pct_of_peak": m['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'],
        "executed": {"dfma_thread_inst": dfma, "dadd_thread_inst": dadd, "dmul_thread_inst": dmul, "dmma_flop": tens,
                     "flop_per_chain_traj": (2 * dfma + dadd + dmul + tens) / chains,
                     "how": "2*DFMA + DADD + DMUL thread instructions (smsp__sass_thread_inst_executed_op_d*_pred_on) + "
                            "sm__ops_path_tensor_src_fp64.sum, per launch / chains"},
        "pipe_active": {"fp64_pct": m['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'],
                        "dmma_subpipe_pct": m['sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active'],
                        "issue_active_pct": m['smsp__issue_active.avg.pct_of_peak_sustained_active'],
                        "note": "sm__pipe_fp64_cycles_active counts the DFMA/DADD/DMUL issue cycles only; the DMMA.8x8x4 tiles "
                                "are counted by sm__pipe_tensor_subpipe_dmma_cycles_active.  Both run on one fp64 datapath "
                                "(profiles/r1_fp64_pipes_probe.txt: 36.9 / 37.1 alone, 34.7 TFLOP/s together), so the datapath's "
                                "duty is their sum"},
        "kernel_ms": m['gpu__time_duration.sum'] if u_['gpu__time_duration.sum'] == 'ms' else None,
    }, open(json_out, "w"), indent=1)
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
