#!/bin/bash
# round 2, call C: why are the 64->64 3x3 layers at 192x256 3x off their MMA time?  timings + one full ncu capture
set -x
mkdir -p gpurun_out
timeout 300 python scripts/time_conv.py > gpurun_out/time_conv.log 2>&1; cat gpurun_out/time_conv.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 16 -c 1 \
    -f -o gpurun_out/prof_conv64 python scripts/time_conv.py > gpurun_out/ncu_conv64.log 2>&1; tail -3 gpurun_out/ncu_conv64.log
ncu -i gpurun_out/prof_conv64.ncu-rep --page raw --csv > gpurun_out/conv64_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_conv64.ncu-rep --page details 2>/dev/null | grep -E "Duration|Theoretical|Achieved|Stall|stall|Warp Cycles|Issued|Eligible|One or More|No Eligible|L2 Hit|DRAM Throughput|Mem Busy|Max Bandwidth|Tensor" | head -50
timeout 200 python scripts/sm_cap_sweep.py 0 70 80 90 100 > gpurun_out/sm_cap_sweep.log 2>&1; cat gpurun_out/sm_cap_sweep.log
