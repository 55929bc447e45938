export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for rep in 1 2; do for k in "0 100" "1 100" "2 100" "2 50"; do set -- $k; echo "== rep $rep cfg3 l2_keep $1 pct $2"; SNK_L2_KEEP=$1 SNK_L2_KEEP_PCT=$2 run cfg3; done; done
