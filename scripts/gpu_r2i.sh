#!/bin/bash
# round 2, call I: band kernel correctness (bit-identical to the gather kernel) and timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_volume_gpu.py tests/test_umma_probe_gpu.py -m gpu -q -x > gpurun_out/t_volume.log 2>&1; tail -12 gpurun_out/t_volume.log | cut -c1-300
NO_SIMT=1 timeout 300 python scripts/time_volume.py > gpurun_out/time_volume.log 2>&1; cat gpurun_out/time_volume.log
