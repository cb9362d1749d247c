#!/bin/bash
# round-2 GPU evidence in one gpurun call: GPU tests, the default bench line, the ncu launch list of the same
# command and one --set full capture of the dominant kernel (each ncu pass only after the plain command exited 0)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests.log
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-check"
$CMD > gpurun_out/plain_ipa.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench_ipa.csv $CMD > gpurun_out/ncu_ipa.log 2>&1
CMD2="python bench.py --batch 4096 --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-check"
$CMD2 > gpurun_out/plain_ipa_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fixed_base_msm -s 5 -c 3 -f -o gpurun_out/r02_prof_ipa $CMD2 > gpurun_out/ncu_ipa_full.log 2>&1
CMD3="python bench.py --workload msm --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-check"
$CMD3 > gpurun_out/plain_msm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_msm_bucket -s 4 -c 1 -f -o gpurun_out/r02_prof_msm $CMD3 > gpurun_out/ncu_msm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench_msm.csv $CMD3 > gpurun_out/ncu_msm_l.log 2>&1
tail -3 gpurun_out/r02_gpu_tests.log; tail -c 600 gpurun_out/r02_bench_default.err
ls -la gpurun_out/ | tail -20
