#!/bin/bash
# One GPU call: staging tests first (new code), full parity suite, bench, SM-cap sweep.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; tail -8 gpurun_out/t_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 200 python scripts/sm_cap_sweep.py > gpurun_out/sm_cap_sweep.log 2>&1; cat gpurun_out/sm_cap_sweep.log
