#!/usr/bin/env python
"""Cut an `ncu -i X.ncu-rep --page raw --csv` dump down to the metrics the profiles/ summaries keep.
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/ncu_<kernel>_<run>_summary.csv"""
import csv
import re
import sys

KEEP = re.compile(r"^(Kernel Name|dram__bytes(_read|_write)?\.sum(\.per_second)?$|dram__throughput\.avg\.pct|gpu__time_duration\.sum|"
                  r"launch__(block_size|grid_size|registers_per_thread$|shared_mem_per_block_dynamic|occupancy_limit_|waves_per)|"
                  r"lts__t_sector_hit_rate\.pct|l1tex__t_sectors_pipe_lsu_mem_global_op_ld\.sum$|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_(ld|st)\.sum$|"
                  r"sm__cycles_elapsed\.max$|sm__throughput\.avg\.pct|sm__inst_executed_pipe_(alu|fma|lsu|xu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_(alu|fma|fmaheavy)_cycles_active\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct|"
                  r"sm__inst_issued\.avg\.pct|smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$|smsp__pcsamp_warps_issue_stalled_[a-z_]+$)")
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
cols = [i for i, n in enumerate(hdr) if KEEP.search(n)]
w = csv.writer(sys.stdout)
for r in rows:
    w.writerow([r[i] if i < len(r) else "" for i in cols])
