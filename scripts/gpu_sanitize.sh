#!/bin/bash
# compute-sanitizer over the CI-sized GPU tests of the mbarrier / TMEM / TMA / cluster kernels ("race
# detection"): memcheck (out-of-bounds / misaligned accesses, also of TMA boxes and DSMEM) and racecheck (shared-memory
# hazards between the warp roles).  Logs -> gpurun_out/sanitize_*.log; summaries are copied to profiles/.
set -x
mkdir -p gpurun_out
SMALL='small or ragged or cfg1 or error or frustum or noncontiguous or per_frame or band'
CONV='not 96-128 and not shape3 and not case13'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 --report-api-errors no \
      python -m pytest tests/test_volume_gpu.py -m gpu -q -x -k "$SMALL" > gpurun_out/sanitize_${tool}_volume.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_${tool}_volume.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 --report-api-errors no \
      python -m pytest tests/test_conv_gpu.py tests/test_mbconv_gpu.py tests/test_umma_probe_gpu.py -m gpu -q -x -k "$CONV" \
      > gpurun_out/sanitize_${tool}_conv.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_${tool}_conv.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 --report-api-errors no \
      python -m pytest tests/test_networks_gpu.py -m gpu -q -x -k "fused_binary_mlp or sample_prior or sigmoid_upsample" \
      > gpurun_out/sanitize_${tool}_heads.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_${tool}_heads.log
done
for f in gpurun_out/sanitize_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=|Error|error" $f | tail -6; done
