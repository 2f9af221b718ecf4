#!/bin/bash
# round 2, final evidence run: full suite, smoke, bench line (gpu_reference + cpu_baseline), launch list, full ncu captures
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_exactness.jsonl
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log | cut -c1-200
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_launch.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 40 > gpurun_out/launches_summary.txt; head -24 gpurun_out/launches_summary.txt
cap() {  # name regex skip
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 \
      -f -o gpurun_out/prof_x_$1 python scripts/profile_step.py > gpurun_out/ncu_x_$1.log 2>&1; tail -1 gpurun_out/ncu_x_$1.log
}
cap fvtc fv_tc_kernel 0
cap halo64 conv_halo_kernel 8
cap halo128 conv_halo_kernel 2
cap convtc conv_tc_kernel 70
cap bmlp binary_mlp_tc_kernel 0
cap stem stem_tc_kernel 0
python scripts/ncu_summary.py gpurun_out/prof_x_fvtc.ncu-rep gpurun_out/prof_x_halo64.ncu-rep gpurun_out/prof_x_halo128.ncu-rep \
    gpurun_out/prof_x_convtc.ncu-rep gpurun_out/prof_x_bmlp.ncu-rep gpurun_out/prof_x_stem.ncu-rep > gpurun_out/ncu_summary_x.md
timeout 300 python scripts/time_configs.py > gpurun_out/time_configs.log 2>&1; cat gpurun_out/time_configs.log | cut -c1-250
