"""Key metrics of an ncu report (`ncu --page raw --csv`): python tools/ncu_rawsum.py <report.ncu-rep>"""
import csv,sys,subprocess
out=subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__warps_eligible.avg.per_cycle_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
for h,u,v in zip(hdr,units,vals):
    try: fv=float(v)
    except: fv=0
    if h in want or ('issue_stalled' in h and 'per_issue' in h and fv>0.05) or ('sm__inst_executed_pipe' in h and 'avg.pct_of_peak_sustained_active' in h and fv>5):
        print(f"{h:90s} {u:12s} {v}")
