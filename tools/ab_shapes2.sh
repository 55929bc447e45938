export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for n in 32768 262144; do echo "== wide8_n N=$n auto"; BENCH_N=$n run wide8_n; done
for c in cfg4 cfg5_shard cfg5_full cfg3 cfg2; do echo "== $c"; run $c; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
