#!/bin/bash
# Full GPU check: parity tests, bench, ncu captures (run under gpurun from the repo root).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
python scripts/time_conv.py > gpurun_out/time_conv.log 2>&1; cat gpurun_out/time_conv.log
python scripts/time_volume.py > gpurun_out/time_volume.log 2>&1; cat gpurun_out/time_volume.log
