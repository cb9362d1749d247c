#!/bin/bash
# ncu evidence for round 1 (run under gpurun; one ncu session per gpurun call as the profiling recipe asks)
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_ipa.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_ipa_r1.csv $CMD > gpurun_out/ncu_ipa.log 2>&1
CMD2="python bench.py --batch 4096 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/plain_ipa_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fixed_base_msm -s 5 -c 3 -f -o gpurun_out/prof_ipa_r1 $CMD2 > gpurun_out/ncu_ipa_full.log 2>&1
CMD3="python bench.py --workload msm --steps 2 --warmup 3 --no-cpu-baseline"
$CMD3 > gpurun_out/plain_msm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_msm_bucket -s 3 -c 1 -f -o gpurun_out/prof_msm_r1 $CMD3 > gpurun_out/ncu_msm.log 2>&1
ls -la gpurun_out/ | tail -12
