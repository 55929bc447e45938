#!/bin/bash
# One GPU-box call that refreshes everything the round's evidence is built from (run under gpurun):
#   GPU test suite, both bench arms, the ncu launch list of the bench command, ncu --set full per config shape.
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs > gpurun_out/${tag}_gpu_tests.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_gpu_tests.txt
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "ref rc=$?"
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 1 > gpurun_out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
bash tools/prof_coop_shapes.sh $tag cfg2 cfg3 cfg4 cfg5_shard cfg5_full
SNK_LIB_PATH=marl-snake_b200/libsnk_prof.so python tools/phase_timing.py cfg4 cfg5_shard cfg2 > gpurun_out/${tag}_phase_timing.jsonl 2> gpurun_out/${tag}_phase_timing.err; echo "phase rc=$?"
python tools/bench_configs.py n1_rollout > gpurun_out/${tag}_n1_rollout.json 2> gpurun_out/${tag}_n1_rollout.err; echo "n1 rc=$?"
