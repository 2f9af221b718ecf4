#!/bin/bash
# N=8 (or $1) with different NCCL channel limits for the per-step gather (send/recv to the root): fewer NCCL CTAs on the
# root's SMs while the next forward runs
set -x
N=${1:-8}
mkdir -p gpurun_out
for cfg in default p2p1 p2p2 ctas2; do
  unset NCCL_MAX_P2P_NCHANNELS NCCL_MAX_CTAS
  case $cfg in
    p2p1) export NCCL_MAX_P2P_NCHANNELS=1;;
    p2p2) export NCCL_MAX_P2P_NCHANNELS=2;;
    ctas2) export NCCL_MAX_CTAS=2;;
  esac
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29610 \
      bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_nccl_${cfg}_n$N.json 2> gpurun_out/bench_nccl_${cfg}_n$N.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_nccl_${cfg}_n$N.json').read().strip().splitlines()[-1])
print('$cfg N=$N', 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), d['e2e']['pipeline'],
      'plain', round(d['e2e']['plain_pipeline'], 1))
PY
done
