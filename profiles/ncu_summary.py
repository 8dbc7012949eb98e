import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','l1tex__throughput.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__average_warp_latency_issue_stalled','smsp__average_warps_issue_stalled','launch__occupancy_limit','sm__maximum_warps_per_active_cycle_pct','achieved_occupancy','smsp__warp_issue_stalled','lts__throughput.avg.pct','l1tex__data_pipe_lsu_wavefronts.sum ','smsp__inst_executed_op_branch','sm__inst_executed_pipe_fp32','smsp__thread_inst_executed_pred_on_per_inst_executed']
for i,h in enumerate(hdr):
    if any(h.startswith(w) for w in want) and not any(x in h for x in ('.max','.min','.sum.p','per_second')):
        print(f'{h:80s} {rows[1][i]:12s}', [r[i][:60] for r in rows[2:]])
