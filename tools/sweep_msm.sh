#!/bin/bash
python -m pytest tests/test_gpu_commit.py -m gpu -x -q 2>&1 | tail -2
for p in 0 4 8 16 32; do
  echo "P=$p"
  VKZG_MSM_P=$p python bench.py --workload msm --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  msm value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'kernel_ms=%.3f' % (r['kernel_ms_total']/r['kernel_launches_timed']), 'frac=%.3f' % r['frac'])"
done
echo two-pass; VKZG_MSM_TWO_PASS=1 python bench.py --workload msm --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  msm value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'])"
for l in 16 18; do python bench.py --workload msm --log2n $l --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  log2n=$l value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'])"; done
