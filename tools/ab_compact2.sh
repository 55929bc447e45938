export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']"; }
for coop in 0 1; do for nc in 0 1; do echo "== shard coop=$coop no_compact=$nc"; SNK_COOP=$coop SNK_NO_COMPACT=$nc run cfg5_shard; done; done
for n in cfg5_256k cfg5_512k; do for coop in 0 1; do for nc in 0 1; do echo "== $n coop=$coop no_compact=$nc"; SNK_COOP=$coop SNK_NO_COMPACT=$nc run $n; done; done; done
for t in 32 64 128; do echo "== cfg5_full compact threads $t"; SNK_THREADS=$t run cfg5_full; done
for t in 64 96 128 160; do echo "== shard coop compact threads $t"; SNK_THREADS=$t run cfg5_shard; done
for e in 2 4 8; do echo "== shard coop compact tile_envs $e"; SNK_TILE_ENVS=$e run cfg5_shard; done
