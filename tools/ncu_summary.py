#!/usr/bin/env python
"""Key metrics of one kernel from an .ncu-rep (raw page): time, DRAM bytes, instruction mix, issue, stalls, shared-memory conflicts."""
import csv
import re
import subprocess
import sys

KEYS = r'^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|smsp__inst_executed\.sum$|smsp__thread_inst_executed\.sum$|smsp__issue_active\.avg\.pct|sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__occupancy_limit|launch__grid_size|launch__block_size|sm__cycles_elapsed\.avg$|smsp__thread_inst_executed_per_inst_executed\.ratio|l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st|atom))?\.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_(ld|st|atom))?\.sum$|smsp__inst_executed_op_shared_(ld|st|atom)\.sum$|smsp__average_warps_issue_stalled_.*_per_issue_active|l1tex__data_pipe_lsu_wavefronts\.avg\.pct|gpu__dram_throughput\.avg\.pct|lts__t_sector_hit_rate\.pct|smsp__inst_executed_op_global_(ld|st)\.sum$|sm__inst_executed_pipe_(alu|fma|lsu|cbu|xu|adu|uniform)\.avg\.pct_of_peak_sustained_active)'
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("==", path, vals[hdr.index("Kernel Name")][:80])
        for h, u, v in zip(hdr, units, vals):
            if re.search(KEYS, h):
                if 'stalled' in h and float(v or 0) < 0.05:
                    continue
                print("  %-90s %-8s %s" % (h, u, v))
