#!/usr/bin/env python
"""Print the key metrics of an .ncu-rep (first kernel): python tools/ncu_summary.py rep [title]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','sm__cycles_elapsed.avg','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','launch__shared_mem_per_block_static','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed']
for vals in rows[2:]:
    ix = {h: i for i, h in enumerate(hdr)}
    print("#", vals[ix['Kernel Name']][:110] if 'Kernel Name' in ix else "", *sys.argv[2:])
    for i, h in enumerate(hdr):
        if h in WANT or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            print(f"{h:78s} {vals[i]:>18s} {units[i]}")
