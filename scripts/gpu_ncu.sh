#!/bin/bash
# ncu captures: launch list of one step + --set full of the volume and top conv kernels.
set -x
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fv_tc_kernel|cv_dot_kernel' -s 12 -c 2 \
    -f -o gpurun_out/prof_volume python scripts/time_volume.py > gpurun_out/ncu_volume.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv_' -s 3 -c 1 \
    -f -o gpurun_out/prof_conv python scripts/time_conv.py > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/ncu_*.log
ls -la gpurun_out
