# usage: python tools/ncu_summary.py <report.ncu-rep> <mangled kernel name> "<header line>" [launch index in the report] > profiles/<name>.txt
import csv, subprocess, sys, os
rep, kern, title = sys.argv[1], sys.argv[2], sys.argv[3]
os.system(f"ncu -i {rep} --page raw --csv 2>/dev/null > /tmp/raws.csv")
r=list(csv.reader(open('/tmp/raws.csv'))); h,u,v=r[0],r[1],r[2+(int(sys.argv[4]) if len(sys.argv)>4 else 0)]
want=['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active',
'dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
'smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active',
'smsp__inst_executed_op_shared_atom.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum',
'sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed']
print(title)
d={n:(u[i],v[i]) for i,n in enumerate(h)}
for n in want:
    if n in d: print("%-70s %-16s %s"%(n,d[n][0],d[n][1]))
print("\nwarp-issue stall reasons (warps per issue-active cycle):")
for i,n in enumerate(h):
    if 'issue_stalled' in n and 'per_issue_active' in n and float(v[i] or 0)>0.05:
        print("  %-40s %.2f"%(n.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),float(v[i])))
print("\nwarp instructions by source function (ncu source page, -lineinfo):")
sys.stdout.flush()
os.system(f"python {os.path.dirname(os.path.abspath(__file__))}/hot2.py {rep} {kern} 2>/dev/null | head -18 | sed 's/^/  /'")
