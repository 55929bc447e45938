export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for c in solo_n duo_n wide8_n; do for n in 32768 262144; do for coop in auto 0 1; do
  echo "== $c N=$n coop=$coop"; if [ $coop = auto ]; then BENCH_N=$n run $c; else BENCH_N=$n SNK_COOP=$coop run $c; fi
done; done; done
