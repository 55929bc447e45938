#!/bin/bash
# ncu --set full with source for the config shapes (cfg2 / cfg3 / cfg4 / 131072-env shard / cfg5 full).
# Usage (under gpurun): bash tools/prof_coop_shapes.sh <tag> [configs...]
# gpurun brings back at most 64 MiB: the reports are exported to CSV on the box (raw page = every metric of the launch,
# source page = per-SASS-line samples) and only the reports named in KEEP_REP (default: cfg4) are kept.
tag=${1:-r02}; shift
cfgs=${@:-cfg2 cfg3 cfg4 cfg5_shard}
keep=${KEEP_REP:-cfg4}
export BENCH_BURN=200 BENCH_STEPS=40
for c in $cfgs; do
  python tools/bench_configs.py $c > gpurun_out/${tag}_plain_$c.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:snk_tile -s 215 -c 1 \
      -o gpurun_out/${tag}_$c -f python tools/bench_configs.py $c > gpurun_out/${tag}_ncu_$c.log 2>&1
  rc=$?
  if [ -f gpurun_out/${tag}_$c.ncu-rep ]; then
    ncu -i gpurun_out/${tag}_$c.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_${c}_raw.csv 2>/dev/null
    ncu -i gpurun_out/${tag}_$c.ncu-rep --page source --csv > gpurun_out/${tag}_ncu_${c}_source.csv 2>/dev/null
    case " $keep " in *" $c "*) ;; *) rm -f gpurun_out/${tag}_$c.ncu-rep ;; esac
  fi
  echo "$c rc=$rc"
done
du -sh gpurun_out
