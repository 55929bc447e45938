export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for k in "0 100" "1 100" "1 50" "1 30" "0 100" "1 100"; do set -- $k; echo "== cfg5_full l2_keep $1 pct $2"; SNK_L2_KEEP=$1 SNK_L2_KEEP_PCT=$2 run cfg5_full; done
for c in cfg5_512k cfg5_256k cfg5_shard cfg3 cfg4; do for k in 0 1; do echo "== $c l2_keep $k"; SNK_L2_KEEP=$k run $c; done; done
SNK_L2_KEEP=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_modes or full_size" 2>&1 | tail -2
