export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
echo "== cfg3 default"; run cfg3
for e in 1 2 4; do for t in 32 64 128; do echo "== cfg3 wp tile_envs $e threads $t"; SNK_COOP=0 SNK_TILE_ENVS=$e SNK_THREADS=$t run cfg3; done; done
for e in 2 4; do for t in 64 96 128; do echo "== cfg3 coop tile_envs $e threads $t"; SNK_COOP=1 SNK_TILE_ENVS=$e SNK_THREADS=$t run cfg3; done; done
