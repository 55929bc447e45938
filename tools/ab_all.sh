run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), round(r['roofline']['frac_record'],4), r['device_errors']) for r in map(json.loads, sys.stdin)]" 2>&1 | tail -3; }
for c in cfg2 cfg3 cfg4 wide8 cfg5_shard cfg5_full; do echo "== $c"; run $c; done
