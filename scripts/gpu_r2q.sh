#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_exactness.jsonl
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log | cut -c1-200
timeout 600 python bench.py --steps 30 --warmup 5 --no-gpu-reference > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -c 300 gpurun_out/bench_q.json; tail -3 gpurun_out/bench_q.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_q.csv python scripts/profile_step.py > gpurun_out/ncu_launch_q.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_q.csv 16
