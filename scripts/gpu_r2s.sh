#!/bin/bash
# per-kernel full captures inside one eager step (profile_step.py), for the per-line instruction breakdown (ncu_lines.py)
set -x
mkdir -p gpurun_out
cap() {  # name regex skip
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 \
      -f -o gpurun_out/prof_$1 python scripts/profile_step.py > gpurun_out/ncu_$1.log 2>&1; tail -1 gpurun_out/ncu_$1.log
}
cap stem stem_tc_kernel 0
cap convtc_match conv_tc_kernel 73
cap convtc_enc0 conv_tc_kernel 0
cap dwconv dwconv3x3_pool_kernel 0
cap upsample upsample2x_kernel 12
cap inapply instnorm_apply_kernel 0
cap mbp maxblurpool_slide_kernel 0
cap sefc1 se_fc1_kernel 0
cap halo16 "conv_halo_kernel<2,.16" 0
cap f32split f32_to_split_kernel 0
