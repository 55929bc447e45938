#!/bin/bash
# ncu --set full with source for the cooperative-tile shapes (cfg2 / cfg3 / cfg4 / 131072-env shard).
# Usage (under gpurun): bash tools/prof_coop_shapes.sh <tag> [configs...]
tag=${1:-r02}; shift
cfgs=${@:-cfg2 cfg3 cfg4 cfg5_shard}
export BENCH_BURN=200 BENCH_STEPS=40
for c in $cfgs; do
  python tools/bench_configs.py $c > gpurun_out/${tag}_plain_$c.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:snk_tile -s 215 -c 1 \
      -o gpurun_out/${tag}_$c -f python tools/bench_configs.py $c > gpurun_out/${tag}_ncu_$c.log 2>&1
  echo "$c rc=$?"
done
