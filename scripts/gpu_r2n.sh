#!/bin/bash
# round 2, evidence run: full suite, bench line (with gpu_reference + cpu_baseline), launch list, full ncu captures
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_exactness.jsonl
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log | cut -c1-200
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_launch.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 40 > gpurun_out/launches_summary.txt; cat gpurun_out/launches_summary.txt
NO_SIMT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:cv_dot_kernel -s 8 -c 1 \
    -f -o gpurun_out/prof_cvdot python scripts/time_volume.py > gpurun_out/ncu_cvdot.log 2>&1
NO_SIMT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:cv_dot_band_kernel -s 8 -c 1 \
    -f -o gpurun_out/prof_cvband python scripts/time_volume.py > gpurun_out/ncu_cvband.log 2>&1
NO_SIMT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fv_tc_kernel -s 8 -c 1 \
    -f -o gpurun_out/prof_fvtc python scripts/time_volume.py > gpurun_out/ncu_fvtc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 16 -c 1 \
    -f -o gpurun_out/prof_conv64 python scripts/time_conv.py > gpurun_out/ncu_conv64.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_cvdot.ncu-rep gpurun_out/prof_cvband.ncu-rep gpurun_out/prof_fvtc.ncu-rep gpurun_out/prof_conv64.ncu-rep > gpurun_out/ncu_summary.md
NO_SIMT=1 timeout 200 python scripts/time_volume.py > gpurun_out/time_volume.log 2>&1; cat gpurun_out/time_volume.log
timeout 300 python scripts/plane_sweep.py > gpurun_out/plane_sweep.log 2>&1; tail -12 gpurun_out/plane_sweep.log
timeout 300 python scripts/time_configs.py > gpurun_out/time_configs.log 2>&1; cat gpurun_out/time_configs.log
for f in gpurun_out/ncu_*.log; do tail -n 1 $f; done
