#!/bin/bash
# Final-build check: parity suite, bench line, ncu launch list of the staged step.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; tail -5 gpurun_out/t_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -c 3000 gpurun_out/bench_f.json; tail -3 gpurun_out/bench_f.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_g.csv python scripts/profile_step.py > gpurun_out/ncu_launch_g.log 2>&1; tail -2 gpurun_out/ncu_launch_g.log
python scripts/summarize_launches.py gpurun_out/launches_g.csv 40
