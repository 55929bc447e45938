# BASELINE shapes at full size against the debug-checks build (device-side range / alignment checks): device_errors must stay 0
export BENCH_MANY=0 BENCH_STEPS=150 BENCH_BURN=100 SNK_LIB_PATH=$PWD/marl-snake_b200/libsnk_dbg.so
for c in cfg2 cfg3 cfg4 cfg5_shard cfg5_full wide8; do python tools/bench_configs.py $c 2>&1 | python -c "import sys,json; [print(r['config'], r['mode'], r['num_envs'], 'steps', r['steps'], 'device_errors', r['device_errors']) for r in map(json.loads, sys.stdin)]"; done
for s in 31415 2718; do SNK_FUZZ_CASES=500 SNK_FUZZ_SEED=$s python -m pytest tests/test_gpu_parity.py -m gpu -q -k fuzz 2>&1 | tail -1; done
unset SNK_LIB_PATH
for s in 1618 99; do SNK_FUZZ_CASES=500 SNK_FUZZ_SEED=$s python -m pytest tests/test_gpu_parity.py -m gpu -q -k fuzz 2>&1 | tail -1; done
