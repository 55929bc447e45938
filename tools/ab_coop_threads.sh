export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
for rep in 1 2; do for t in 32 64; do echo "== rep $rep cfg2 threads $t"; SNK_THREADS=$t run cfg2; done; done
for n in 8192 32768; do for t in 32 64; do echo "== cfg5 shape N=$n threads $t"; BENCH_N=$n SNK_THREADS=$t run cfg5_n; done; done
