#!/bin/bash
# Instruction count / duration / L1TEX / DRAM of one snk_tile_kernel launch of a config: bash tools/ncu_quick.sh <config> [ENV=VAL ...]
cfg=$1; shift
for kv in "$@"; do export "$kv"; done
export BENCH_BURN=200 BENCH_STEPS=40
python tools/bench_configs.py $cfg > /dev/null 2>&1 &&
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,launch__registers_per_thread,launch__block_size \
  --clock-control none -k regex:snk_tile -s 215 -c 1 --csv python tools/bench_configs.py $cfg 2>/dev/null | grep -E "snk_tile" | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | paste -sd' ' | sed "s/^/$cfg $* : /"
