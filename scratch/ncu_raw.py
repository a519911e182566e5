import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; units = rows[1]
idx = {h:i for i,h in enumerate(hdr)}
want = ['Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active','launch__shared_mem_per_block_dynamic','sm__inst_executed.sum','smsp__inst_executed_op_shared_ld.sum','sm__cycles_active.avg', 'lts__t_sector_hit_rate.pct','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct']
have = [w for w in want if w in idx]
sel = sys.argv[2:] 
for r in rows[2:]:
    if sel and not any(s in r[idx['Kernel Name']] for s in sel): continue
    print('----')
    for w in have:
        print('  %-75s %s %s' % (w, r[idx[w]][:100], units[idx[w]]))
