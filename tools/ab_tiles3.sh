export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for rep in 1 2; do for lib in libsnk.so libsnk_wp8.so; do for t in 32 128; do
  echo "== rep $rep shard $lib threads $t"; SNK_LIB_PATH=marl-snake_b200/$lib SNK_THREADS=$t run cfg5_shard
  echo "== rep $rep full $lib threads $t"; SNK_LIB_PATH=marl-snake_b200/$lib SNK_THREADS=$t run cfg5_full
done; done; done
