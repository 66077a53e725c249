import sys,csv,subprocess,io
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr,units=rows[0],rows[1]
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','lts__t_bytes.sum ','lts__throughput.avg.pct','l1tex__throughput.avg.pct','smsp__average_warps_issue_stalled','smsp__warps_eligible.avg.per_cycle_active','lts__t_sectors_op_red.sum','lts__t_sectors_op_atom.sum','local']
for r in rows[2:]:
    print('KERNEL', r[hdr.index('Kernel Name')][:60])
    for i,h in enumerate(hdr):
        if any(k in h for k in keys) and r[i] not in ('','0'):
            print('  ',h,units[i],r[i])
