import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
d={h:(u,v) for h,u,v in zip(hdr,units,vals)}
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__warps_eligible.avg.per_cycle_active','dram__bytes_read.sum','dram__bytes_write.sum','sm__cycles_elapsed.avg','gpc__cycles_elapsed.avg.per_second']
keys+= [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h or 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared' in h]
for k in keys:
    if k in d: print(f"{k:95s} {d[k][0]:12s} {d[k][1]}")
