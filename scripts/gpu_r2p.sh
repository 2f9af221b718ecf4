#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_networks_gpu.py -m gpu -q -x -k "mlp or forward" > gpurun_out/t_net.log 2>&1; tail -3 gpurun_out/t_net.log | cut -c1-200
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_p.csv python scripts/profile_step.py > gpurun_out/ncu_launch_p.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_p.csv 10
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:binary_mlp -c 1 \
    -f -o gpurun_out/prof_bmlp python scripts/profile_step.py > gpurun_out/ncu_bmlp.log 2>&1; tail -2 gpurun_out/ncu_bmlp.log
