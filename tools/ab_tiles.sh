export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for e in 1 2; do for t in 128 192 256; do echo "== cfg4 tile_envs $e threads $t"; SNK_TILE_ENVS=$e SNK_THREADS=$t run cfg4; done; done
for t in 32 64 128; do echo "== shard threads $t"; SNK_THREADS=$t run cfg5_shard; done
for t in 32 64 128; do echo "== N=65536 threads $t"; BENCH_N=65536 SNK_THREADS=$t run cfg5_n; done
for e in 1 2 4; do for t in 64 128; do echo "== cfg3 tile_envs $e threads $t"; SNK_TILE_ENVS=$e SNK_THREADS=$t run cfg3; done; done
