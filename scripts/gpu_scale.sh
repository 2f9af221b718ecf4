#!/bin/bash
# scaling run on one 8-GPU node: bench.py at N = 8, 4, 2, 1 (weak scaling, 4 frames per GPU), the driver's launch line
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N \
      bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N', 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), d['e2e']['pipeline'],
      'plain', round(d['e2e']['plain_pipeline'], 1), 'd2h', d['e2e']['d2h_bytes_per_step'])
PY
done
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('N=1', 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), d['e2e']['pipeline'],
      'plain', round(d['e2e']['plain_pipeline'], 1))
PY
