#!/bin/bash
# round 2, call B: kernel tests of the new conv / pool variants, full parity suite, bench, launch list
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_exactness.jsonl
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_umma_probe_gpu.py -m gpu -q -x > gpurun_out/t_conv.log 2>&1; tail -8 gpurun_out/t_conv.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_conv_gpu.py --deselect tests/test_umma_probe_gpu.py > gpurun_out/t_gpu.log 2>&1; tail -8 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'stage_ms')}, d['e2e']['value'], d['e2e']['pipeline'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -5 gpurun_out/bench.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_launch.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 30
