#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/time_conv.py > gpurun_out/time_conv.log 2>&1; cat gpurun_out/time_conv.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 16 -c 1 \
    -f -o gpurun_out/prof_conv64 python scripts/time_conv.py > gpurun_out/ncu_conv64.log 2>&1; tail -2 gpurun_out/ncu_conv64.log
ncu -i gpurun_out/prof_conv64.ncu-rep --page source --csv > gpurun_out/conv64_src.csv 2>/dev/null
ncu -i gpurun_out/prof_conv64.ncu-rep --page raw --csv > gpurun_out/conv64_raw.csv 2>/dev/null
