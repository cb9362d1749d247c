#!/bin/bash
# round 2: split MSM (second scatter under the first bucket pass): correctness, then split denominators and lanes per bucket
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_commit.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --workload msm --log2n $1 --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'kernel_ms_total/step=%.3f' % (r['kernel_ms_total']/d['steps']), d['checked']['ok'])"; }
for l in 20 19; do
  for sp in 0 2 3 4 6 8; do echo -n "log2n=$l split=$sp: "; VKZG_MSM_SPLIT=$sp run $l; done
done
for p in 2 8; do echo -n "log2n=20 split=0 P=$p: "; VKZG_MSM_SPLIT=0 VKZG_MSM_P=$p run 20; done
for p in 2 4 8; do echo -n "log2n=20 split=4 P=$p: "; VKZG_MSM_SPLIT=4 VKZG_MSM_P=$p run 20; done
