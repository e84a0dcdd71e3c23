import csv,sys,subprocess
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__waves_per_multiprocessor','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
out=subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
if len(sys.argv)>2:
    for h in hdr:
        if sys.argv[2] in h: print(h)
    sys.exit()
idx=[hdr.index(w) for w in want if w in hdr]
seen=set()
for r in rows[2:]:
    name=r[hdr.index('Kernel Name')]
    if name in seen: continue
    seen.add(name)
    print('---')
    for i in idx: print('  ',hdr[i],'=',r[i], rows[1][i])
