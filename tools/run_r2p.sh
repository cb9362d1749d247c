#!/bin/bash
# round 2: buckets served in order of decreasing list length; threads per group sum
python -m pytest tests/test_gpu_commit.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --workload msm --log2n $1 --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'kernel_ms=%.3f' % (r['kernel_ms_total']/r['kernel_launches_timed']), d['checked']['ok'])"; }
for l in 20 18 17 16; do
  echo -n "log2n=$l default: "; run $l
  echo -n "log2n=$l no order: "; VKZG_MSM_NO_ORDER=1 run $l
  for g in 32 64; do echo -n "log2n=$l gs=$g: "; VKZG_MSM_GS=$g run $l; done
done
for p in 4 16; do echo -n "log2n=20 P=$p: "; VKZG_MSM_P=$p run 20; done
