export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for lib in libsnk.so libsnk_wp8.so libsnk.so libsnk_wp8.so; do echo "== cfg5_full $lib"; SNK_LIB_PATH=marl-snake_b200/$lib run cfg5_full; done
for lib in libsnk.so libsnk_wp8.so; do echo "== shard $lib"; SNK_LIB_PATH=marl-snake_b200/$lib run cfg5_shard; done
for t in 96 128 160; do echo "== cfg4 default tile threads $t"; SNK_THREADS=$t run cfg4; done
echo "== wide8"; run wide8
echo "== cfg2"; run cfg2
echo "== cfg3"; run cfg3
