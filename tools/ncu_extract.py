#!/usr/bin/env python
"""Extract the judged metrics of one `ncu --set full` report (first kernel) into a small text file:
   ncu -i rep --page raw --csv | python tools/ncu_extract.py > profiles/xxx.txt"""
import csv
import re
import sys

KEEP = [
    r"^Kernel Name$", r"^launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_.*|waves_per_multiprocessor)$",
    r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_bytes\.sum$", r"^lts__t_sector_hit_rate\.pct$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__pipe_(shared|fp64|tensor)_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
    r"^sm__pipe_tensor_subpipe_dmma_cycles_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_(fp64|lsu|tensor_subpipe_dmma|alu|fma|xu)\.sum$", r"^smsp__inst_executed\.sum$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^l1tex__data_pipe_lsu_wavefronts(_mem_shared)?\.sum\.pct_of_peak_sustained_elapsed$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
    r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$", r"^sm__cycles_elapsed\.max$", r"^smsp__cycles_active\.avg$",
]
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
out = []
for i, h in enumerate(hdr):
    if any(re.search(p, h) for p in KEEP):
        v = vals[i]
        if "stalled" in h:
            try:
                if float(v.replace(",", "")) < 0.3:
                    continue
            except ValueError:
                pass
        out.append((h, v, units[i]))
for h, v, u in out:
    print(f"{h:92s} {v:>22s} {u}")
