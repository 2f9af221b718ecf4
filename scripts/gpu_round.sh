#!/bin/bash
# One GPU call (run under gpurun from the repo root): parity tests, bench line, ncu launch list of the staged step,
# ncu --set full captures of the volume kernels and the heaviest conv, plane sweep, other configurations, scheduling
# sweeps.  Everything lands in gpurun_out/; copy what should be judged into profiles/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; tail -15 gpurun_out/t_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_launch.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 40
NO_SIMT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:cv_dot_kernel -s 8 -c 1 \
    -f -o gpurun_out/prof_cvdot python scripts/time_volume.py > gpurun_out/ncu_cvdot.log 2>&1
NO_SIMT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fv_tc_kernel -s 8 -c 1 \
    -f -o gpurun_out/prof_fvtc python scripts/time_volume.py > gpurun_out/ncu_fvtc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 \
    -f -o gpurun_out/prof_conv python scripts/time_conv.py > gpurun_out/ncu_conv.log 2>&1
timeout 300 python scripts/plane_sweep.py > gpurun_out/plane_sweep.log 2>&1; cat gpurun_out/plane_sweep.log
timeout 200 python scripts/time_configs.py > gpurun_out/time_configs.log 2>&1; cat gpurun_out/time_configs.log
timeout 100 python scripts/time_native_encoder.py > gpurun_out/time_enc.log 2>&1; cat gpurun_out/time_enc.log
timeout 200 python scripts/sm_cap_sweep.py > gpurun_out/sm_cap_sweep.log 2>&1; cat gpurun_out/sm_cap_sweep.log
timeout 300 python scripts/time_encoder_ahead.py > gpurun_out/time_ahead.log 2>&1; cat gpurun_out/time_ahead.log
for f in gpurun_out/ncu_*.log; do tail -n 2 $f; done
ls -la gpurun_out
