#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -x -s > gpurun_out/t_parallel.log 2>&1; grep -n "rank \|GATHER\|passed\|failed" gpurun_out/t_parallel.log | head -20
