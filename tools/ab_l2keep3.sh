export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for rep in 1 2; do for k in 0 1; do echo "== rep $rep cfg3 l2_keep $k"; SNK_L2_KEEP=$k run cfg3; done; done
echo "== cfg3 default"; run cfg3
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
