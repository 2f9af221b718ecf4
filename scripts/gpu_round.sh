#!/bin/bash
# One GPU call: parity tests, bench, kernel timings, ncu launch list + --set full captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; tail -15 gpurun_out/t_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 120 python scripts/time_conv.py > gpurun_out/time_conv.log 2>&1; cat gpurun_out/time_conv.log
timeout 120 python scripts/time_volume.py > gpurun_out/time_volume.log 2>&1; cat gpurun_out/time_volume.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fv_tc_kernel|cv_dot_kernel' -s 12 -c 2 \
    -f -o gpurun_out/prof_volume python scripts/time_volume.py > gpurun_out/ncu_volume.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_' -s 3 -c 1 \
    -f -o gpurun_out/prof_conv python scripts/time_conv.py > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/ncu_*.log
ls -la gpurun_out
