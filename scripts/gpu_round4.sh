#!/bin/bash
# Fused MBConv chain: unit tests, encoder parity, full suite, A/B timing (old chain vs new, with cap/split variants).
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mbconv_gpu.py -q > gpurun_out/t_mbconv.log 2>&1; tail -15 gpurun_out/t_mbconv.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_mbconv_gpu.py > gpurun_out/t_gpu.log 2>&1; tail -8 gpurun_out/t_gpu.log
B200_MBCONV_FUSED=0 timeout 100 python scripts/sm_cap_sweep.py "74;74;0;" > gpurun_out/mb_sweep_old.log 2>&1; cat gpurun_out/mb_sweep_old.log
B200_MBCONV_FUSED=1 timeout 200 python scripts/sm_cap_sweep.py "74;74;0;" "74;74;0;2:74,2:148" "74;74;0;3:74,1:148" "86;86;0;" "86;86;0;3:86,1:148" "64;64;0;" "74;74;0;" > gpurun_out/mb_sweep_new.log 2>&1; cat gpurun_out/mb_sweep_new.log
B200_MBCONV_FUSED=0 timeout 100 python scripts/time_native_encoder.py > gpurun_out/time_enc.log 2>&1
B200_MBCONV_FUSED=1 timeout 100 python scripts/time_native_encoder.py >> gpurun_out/time_enc.log 2>&1; tail -5 gpurun_out/time_enc.log
