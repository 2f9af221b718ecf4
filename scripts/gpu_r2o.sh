#!/bin/bash
# round 2: binary MLP with two threads per pixel -- network tests, bench, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_networks_gpu.py -m gpu -q -x > gpurun_out/t_net.log 2>&1; tail -3 gpurun_out/t_net.log | cut -c1-200
timeout 600 python bench.py --steps 30 --warmup 5 --no-gpu-reference > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; tail -c 900 gpurun_out/bench_o.json; tail -3 gpurun_out/bench_o.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_o.csv python scripts/profile_step.py > gpurun_out/ncu_launch_o.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_o.csv 14
