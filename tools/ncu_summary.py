#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into the text kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>.txt
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active", "sm__pipe_tensor_subpipe_hmma_cycles_active",
        "sm__inst_executed_pipe_uniform", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput",
        "gpu__dram_throughput", "lts__t_bytes.sum", "lts__throughput", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct",
        "sm__cycles_elapsed.avg ", "sm__cycles_active.avg ", "smsp__inst_executed.sum ", "sm__inst_executed_pipe_fma",
        "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_xu", "smsp__inst_executed_pipe_lsu",
        "l1tex__m_xbar2l1tex_read_bytes.sum ", "sm__sass_inst_executed_op_shared", "smsp__cycles_active.avg "]


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
for row in raw[2:]:
    name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"== kernel: {name}")
    for h, u, v in zip(hdr, units, row):
        if any(k.strip() in h for k in KEYS):
            print(f"  {h:95s} {v:>22s} {u}")

src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv"]))))
if len(src) > 2:
    h = src[1]
    ix = {k: i for i, k in enumerate(h)}
    data = [r for r in src[2:] if len(r) == len(h) and r[ix["# Samples"]].strip().isdigit()]
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    print(f"== warp-stall samples: {tot}; top instructions")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:25]:
        st = sorted(((s[6:], int(r[ix[s]])) for s in stalls if int(r[ix[s]]) > 0), key=lambda kv: -kv[1])[:2]
        print(f"  {100 * int(r[ix['# Samples']]) / tot:5.1f}%  thr={r[ix['Avg. Threads Executed']]:>3s}  "
              f"{r[ix['Source']].strip()[:72]:72s} {st}")
    agg = {}
    for r in data:
        for s in stalls:
            agg[s[6:]] = agg.get(s[6:], 0) + int(r[ix[s]])
    print("== stall reasons (all warps):", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
