#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_volume_gpu.py tests/test_networks_gpu.py -m gpu -q -x > gpurun_out/t_net.log 2>&1; tail -3 gpurun_out/t_net.log | cut -c1-200
timeout 600 python scripts/time_encoder_ahead.py "0;0;-" "1;0;-" "1;0;-;132;16" "1;0;-;124;24" "1;0;-;116;32" "1;0;-;0;32" "1;0;-;0;16" "1;0;-;132;0" "1;-1;-;124;24" > gpurun_out/enc_ahead_caps.jsonl 2>&1; cat gpurun_out/enc_ahead_caps.jsonl | cut -c1-200
