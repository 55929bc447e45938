#!/bin/bash
# Tile-shape sweep of the CTA-cooperative mode on the secondary BASELINE shapes (one GPU).
# Usage: tools/sweep_tiles.sh > gpurun_out/sweep_tiles.jsonl
export BENCH_STEPS=${BENCH_STEPS:-400}
for cfg in cfg2 cfg3 cfg4; do
  case $cfg in cfg2) tes="8 4 2 1";; cfg3) tes="8 4 2";; cfg4) tes="2 1";; esac
  for te in $tes; do
    for th in 64 128 256; do
      SNK_TILE_ENVS=$te SNK_THREADS=$th python tools/bench_configs.py $cfg 2>&1 | grep '^{'
    done
  done
done
