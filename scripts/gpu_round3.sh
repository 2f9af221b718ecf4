#!/bin/bash
# Plane-sweep split sweep + ncu launch list of the staged step.
set -x
mkdir -p gpurun_out
timeout 200 python scripts/sm_cap_sweep.py "74;74;0;" "74;74;0;2:74,2:148" "74;74;0;1:74,3:148" "74;74;0;3:74,1:148" "74;74;0;2:74,1:110,1:148" "74;74;0;2:74,2:110" "74;74;0;" > gpurun_out/sm_split_sweep.log 2>&1; cat gpurun_out/sm_split_sweep.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_f.csv python scripts/profile_step.py > gpurun_out/ncu_launch_f.log 2>&1; tail -2 gpurun_out/ncu_launch_f.log
python scripts/summarize_launches.py gpurun_out/launches_f.csv 30
