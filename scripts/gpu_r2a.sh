#!/bin/bash
# round 2, call A: parity at the benchmarked configs + bench line with the gpu_reference block
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_exactness.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; tail -25 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
