export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for n in 8192 16384 32768 65536; do for mode in "0 0" "1 0" "1 1"; do set -- $mode; echo "== cfg5 shape N=$n coop=$1 no_compact=$2"; BENCH_N=$n SNK_COOP=$1 SNK_NO_COMPACT=$2 run cfg5_n; done; done
for c in cfg3 cfg4 cfg2 wide8; do for mode in "0 0" "1 0" "1 1"; do set -- $mode; echo "== $c coop=$1 no_compact=$2"; SNK_COOP=$1 SNK_NO_COMPACT=$2 run $c; done; done
