import csv, sys, subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__occupancy_limit_warps','launch__occupancy_limit_blocks','sm__inst_issued.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','lts__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed_pipe_lsu.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:70] if 'Kernel Name' in hdr else '')
    for k in keys:
        if k in hdr:
            i=hdr.index(k); print(' ',k, units[i], r[i])
