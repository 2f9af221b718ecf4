#!/bin/bash
# round 2, call J: evidence for profiles/ -- bench line, launch list of one step, full ncu captures of the volume kernels
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_launch.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 40 > gpurun_out/launches_summary.txt; cat gpurun_out/launches_summary.txt
NO_SIMT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:cv_dot_kernel -s 8 -c 1 \
    -f -o gpurun_out/prof_cvdot python scripts/time_volume.py > gpurun_out/ncu_cvdot.log 2>&1
NO_SIMT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:cv_dot_band_kernel -s 8 -c 1 \
    -f -o gpurun_out/prof_cvband python scripts/time_volume.py > gpurun_out/ncu_cvband.log 2>&1
NO_SIMT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fv_tc_kernel -s 8 -c 1 \
    -f -o gpurun_out/prof_fvtc python scripts/time_volume.py > gpurun_out/ncu_fvtc.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_cvdot.ncu-rep gpurun_out/prof_cvband.ncu-rep gpurun_out/prof_fvtc.ncu-rep > gpurun_out/ncu_volume_summary.md
head -60 gpurun_out/ncu_volume_summary.md
timeout 300 python scripts/plane_sweep.py > gpurun_out/plane_sweep.log 2>&1; cat gpurun_out/plane_sweep.log
for f in gpurun_out/ncu_*.log; do tail -n 1 $f; done
