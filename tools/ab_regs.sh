export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for lib in libsnk.so libsnk_mb5.so libsnk_mb6.so; do for t in 96 128 160; do echo "== cfg4 $lib threads $t"; SNK_LIB_PATH=marl-snake_b200/$lib SNK_THREADS=$t run cfg4; done; done
for e in 4 8; do for t in 32 64 128; do echo "== cfg5_full tile_envs $e threads $t"; SNK_TILE_ENVS=$e SNK_THREADS=$t run cfg5_full; done; done
