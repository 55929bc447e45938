export BENCH_MANY=0
run() { python tools/bench_configs.py $1 2>&1 | python -c "import sys,json; [print('   ', r['mode'], r['num_envs'], round(r['ms_per_step'],5), round(r['roofline']['frac'],4), r['device_errors']) for r in map(json.loads, sys.stdin) if r['mode']=='eager']" 2>&1 | tail -3; }
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_modes or scenarios or rollout_replay or alternative or masked or fuzz" 2>&1 | tail -3
for rep in 1 2; do for h in 1 0; do for c in cfg4 cfg3 cfg2; do echo "== rep $rep $c help=$h"; SNK_HELP=$h run $c; done; done; done
for n in 8192 32768; do for h in 1 0; do echo "== cfg5 shape N=$n help=$h"; BENCH_N=$n SNK_HELP=$h run cfg5_n; done; done
SNK_LIB_PATH=marl-snake_b200/libsnk_prof.so python tools/phase_timing.py cfg4 2>&1 | cut -c1-700
