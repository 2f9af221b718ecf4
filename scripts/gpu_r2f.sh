#!/bin/bash
# round 2, call F: timm-layout encoder + thresholder + split-K: full GPU suite, bench
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_exactness.jsonl
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; tail -12 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'stage_ms')}, d['e2e']['value'], d['e2e']['pipeline'], d['e2e']['plain_pipeline'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -5 gpurun_out/bench.err
