python -m pytest tests/test_gpu_commit.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
for l in 16 17 18; do for p in 0 8 16 32; do
  echo -n "log2n=$l P=$p: "
  VKZG_MSM_P=$p python bench.py --workload msm --log2n $l --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'kernel_ms=%.3f' % (r['kernel_ms_total']/r['kernel_launches_timed']), 'frac=%.3f' % r['frac'], d['checked']['ok'])"
done; done
CMD="python bench.py --workload msm --log2n 16 --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-check"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches_msm_2p16.csv $CMD > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02e_launches_msm_2p16.csv k_msm_scatter
