#!/bin/bash
# round 2, call H: full suite after the pipeline / split-factor changes, then compute-sanitizer runs
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_exactness.jsonl
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; tail -4 gpurun_out/t_gpu.log | cut -c1-300
bash scripts/gpu_sanitize.sh
