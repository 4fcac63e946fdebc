#!/usr/bin/env python
"""Condense an ncu report (raw + source pages exported as CSV) into the handful of numbers the
design notes quote: time, DRAM bytes, pipe utilisation, shared wavefronts, instructions per unit,
stall reasons per issue, executed opcode mix.
    ncu -i X.ncu-rep --page raw --csv > X_raw.csv; ncu -i X.ncu-rep --page source --csv > X_source.csv
    python tools/ncu_summary.py X_raw.csv X_source.csv <units per launch>"""
import csv
import sys
from collections import Counter

raw, src, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__icc_request_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__cycles_elapsed.max", "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    print(f"{k:72s} {m.get(k)}")
print("stall cycles per issued instruction:")
for k, v in sorted(((k, float(v)) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")), key=lambda kv: -kv[1])[:9]:
    print(f"  {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
ops, tot = Counter(), 0
for r in data:
    s = r[ix["Source"]].split()
    op = s[0] if not s[0].startswith("@") else s[1]
    op = op.split(".")[0] if not op.startswith(("LDS", "STS", "STG", "LDG")) else op
    ex = int(r[ix["Instructions Executed"]])
    ops[op] += ex
    tot += ex
print(f"warp instructions per unit: {tot / units:.1f}")
print("  " + ", ".join(f"{op} {c / units:.1f}" for op, c in ops.most_common(24)))
