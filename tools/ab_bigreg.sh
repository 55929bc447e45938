export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for c in cfg5_shard cfg5_256k cfg5_512k cfg5_full; do for b in 1 100000000; do for t in 32 128; do echo "== $c bigreg_tiles $b threads $t"; SNK_BIGREG_TILES=$b SNK_THREADS=$t run $c; done; done; done
