#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python scripts/sm_cap_sweep.py "74;74;0;" "80;80;0;" "74;74;0;3:74,1:120" "80;80;0;3:80,1:120" "78;78;0;" "82;82;0;" "80;80;0;" "74;74;0;3:74,1:120" "80;80;0;3:80,1:120" "74;74;0;" > gpurun_out/mb_sweep_new3.log 2>&1; cat gpurun_out/mb_sweep_new3.log
