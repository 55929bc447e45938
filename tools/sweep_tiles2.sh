#!/bin/bash
export BENCH_STEPS=${BENCH_STEPS:-400}
run() { SNK_COOP=$2 SNK_TILE_ENVS=$3 SNK_THREADS=$4 python tools/bench_configs.py $1 2>&1 | grep '^{'; }
for te in 1 2; do for th in 32 64 128; do run cfg4 0 $te $th; done; done
for th in 96 160 192; do run cfg4 1 1 $th; done
for th in 32 64; do run cfg3 1 1 $th; done
run cfg3 1 2 32
for te in 1 2 4; do for th in 32 64; do run cfg3 0 $te $th; done; done
for te in 1 2; do for th in 32 64 128; do run cfg2 0 $te $th; done; done
run cfg5_shard 0 8 32
for th in 32 64; do run cfg5_shard 0 4 $th; done
for te in 2 4 8; do run cfg5_shard 1 $te 128; done
