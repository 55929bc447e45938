export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for ct in 0 8192 16384 32768 65536; do for b in 1 100000000; do echo "== cfg5_full chunk $ct bigreg_tiles $b"; SNK_CHUNK_TILES=$ct SNK_BIGREG_TILES=$b run cfg5_full; done; done
for c in cfg5_512k; do for ct in 0 16384 32768; do echo "== $c chunk $ct"; SNK_CHUNK_TILES=$ct run $c; done; done
