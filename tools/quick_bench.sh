#!/bin/bash
# quick perf + parity loop used during kernel work
python -m pytest tests/test_gpu_commit.py tests/test_gpu_ipa.py -m gpu -x -q 2>&1 | tail -3
for w in commit msm ipa; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/qb_$w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w', 'value=%.4g %s' % (d['value'], d['unit']), 'ms/step=%.3f' % d['ms_per_step'], 'e2e=%.4g' % d['e2e']['value'], 'frac=%.3f' % r['frac'], 'kernel_ms=%.3f' % (r['kernel_ms_total']/max(1,r['kernel_launches_timed'])), 'share=%.3f' % r['kernel_share_of_step'], 'launches', d['gpu_launches'])
" || tail -5 gpurun_out/qb_$w.err
done
