set -x
python -m pytest tests/test_gpu_kzg.py tests/test_gpu_crs.py -m gpu -x -q > gpurun_out/r02c_tests.log 2>&1; tail -5 gpurun_out/r02c_tests.log
python tools/setup_bench.py --json gpurun_out/r02_setup_bench.json > gpurun_out/r02_setup_bench.log 2>&1; tail -12 gpurun_out/r02_setup_bench.log
CMD="python bench.py --workload msm --log2n 16 --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-check"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_msm_2p16.csv $CMD > /dev/null 2>&1
CMD="python bench.py --workload msm --log2n 18 --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-check"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_msm_2p18.csv $CMD > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_msm_2p16.csv k_msm_scatter
python tools/launch_summary.py gpurun_out/r02_launches_msm_2p18.csv k_msm_scatter
