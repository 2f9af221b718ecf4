#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q -x > gpurun_out/t_conv.log 2>&1; tail -5 gpurun_out/t_conv.log | cut -c1-300
timeout 300 python scripts/time_conv.py > gpurun_out/time_conv.log 2>&1; cat gpurun_out/time_conv.log
