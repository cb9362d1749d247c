set -x
python -m pytest tests/test_gpu_crs.py -m gpu -x -q > gpurun_out/r02_crs_tests.log 2>&1; tail -5 gpurun_out/r02_crs_tests.log
python tools/setup_bench.py --json gpurun_out/r02_setup_bench.json > gpurun_out/r02_setup_bench.log 2>&1; tail -12 gpurun_out/r02_setup_bench.log
for l in 16 18; do python bench.py --workload msm --log2n $l --steps 10 --warmup 3 --no-also 2>/dev/null >> gpurun_out/r02_bench_msm_sizes.jsonl; done
python -c "
import json
for l in open('gpurun_out/r02_bench_msm_sizes.jsonl'):
    d=json.loads(l); print(d['config'].get('workload'), d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('kernel_share_of_step'))
"
