export BENCH_MANY=0
for c in cfg5_full cfg5_shard cfg4 cfg3 cfg2; do
  echo "== $c compact"; python tools/bench_configs.py $c 2>&1 | python -c "import sys,json; [print(r['mode'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['roofline']['record_bytes_per_env_step'], r['device_errors']) for r in map(json.loads, sys.stdin)]"
  echo "== $c full"; SNK_NO_COMPACT=1 python tools/bench_configs.py $c 2>&1 | python -c "import sys,json; [print(r['mode'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['roofline']['record_bytes_per_env_step'], r['device_errors']) for r in map(json.loads, sys.stdin)]"
done
for t in 64 128; do echo "== cfg5_full compact threads $t"; SNK_THREADS=$t python tools/bench_configs.py cfg5_full 2>&1 | python -c "import sys,json; [print(r['mode'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4)) for r in map(json.loads, sys.stdin)]"; done
