#!/bin/bash
# round 2: full GPU suite on the build with the dedicated square / prove_all / two-level weighted bucket sum, then the MSM
# sweep: tail form (VKZG_MSM_TAIL 1 = bit-parallel, 2 = running sums, 3 = two-level) x window width
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --workload msm --log2n $1 --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'kernel_ms=%.3f' % (r['kernel_ms_total']/r['kernel_launches_timed']), d['checked']['ok'])"; }
for l in 16 17 18 19 20; do
  case $l in 16) cs="12 13";; 17) cs="13 15";; 18|19) cs="15 16 17";; 20) cs="16 17 18 19 20";; esac
  for c in $cs; do for t in 1 2 3; do
    if [ $t = 1 ] && [ $c -gt 17 ]; then continue; fi
    echo -n "log2n=$l c=$c tail=$t: "; VKZG_MSM_C=$c VKZG_MSM_TAIL=$t run $l
  done; done
done
