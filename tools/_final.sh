timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/r01_bench_n1.err; tail -2 gpurun_out/r01_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r01_bench_reference_arm.json 2>&1
for th in 64 128; do SNK_THREADS=$th python tools/bench_configs.py cfg5_full 2>&1 | grep '^{' >> gpurun_out/private_threads.jsonl; done
python tools/bench_configs.py > gpurun_out/r01_all_configs.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
cut -c1-1200 gpurun_out/r01_bench_n1.json; cut -c1-200 gpurun_out/private_threads.jsonl
