timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SNK_NO_DIG=1 timeout 900 python -m pytest tests -m gpu -x -q -k "tile_modes or rollout_replay or scenarios or fuzz" 2>&1 | tail -3
SNK_FUZZ_CASES=300 SNK_FUZZ_SEED=99 timeout 900 python -m pytest tests -m gpu -q -k fuzz 2>&1 | tail -2
python tools/bench_configs.py > gpurun_out/all_configs_dig.jsonl 2>&1
SNK_NO_DIG=1 python tools/bench_configs.py cfg5_full cfg5_shard cfg3 cfg2 >> gpurun_out/all_configs_dig.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/all_configs_dig.jsonl'):
    d=json.loads(l); print(d['config'], 'graph' if d['graph'] else '', '%.4f ms'%d['ms_per_step'], '%.3f'%d['frac_of_measured_peak'], '%.0f B/agent'%d['bytes_per_agent_step'], '%.3f G'%(d['agent_steps_per_sec']/1e9), d['device_errors'])
PY
