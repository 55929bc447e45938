python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for enc in 0 1; do for th in 64 128 192 256; do SNK_ENC_LEGACY=$enc SNK_THREADS=$th python tools/bench_configs.py cfg4 2>&1 | grep '^{' | sed "s/^{/{\"enc_legacy\": $enc, /" >> gpurun_out/wide_cmp.jsonl; done; done
for te in 1 2; do SNK_COOP=0 SNK_TILE_ENVS=$te SNK_THREADS=64 python tools/bench_configs.py cfg4 2>&1 | grep '^{' | sed "s/^{/{\"enc_legacy\": 0, /" >> gpurun_out/wide_cmp.jsonl; done
SNK_TILE_ENVS=2 SNK_THREADS=256 python tools/bench_configs.py cfg4 2>&1 | grep '^{' | sed "s/^{/{\"enc_legacy\": 0, /" >> gpurun_out/wide_cmp.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/wide_cmp.jsonl'):
    d=json.loads(l); print(d['enc_legacy'], d['coop'], d['tile_envs'], d['threads'], '%.4f'%d['ms_per_step'], '%.3f'%d['frac_of_measured_peak'], d['device_errors'])
PY
