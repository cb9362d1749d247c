#!/bin/bash
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
for c in 16 18 19 20; do
  echo "c=$c"
  python bench.py --workload commit --window-bits $c --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/w$c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  commit value=%.4g' % d['value'], 'e2e=%.4g' % d['e2e']['value'], 'frac=%.3f' % r['frac'])" || tail -3 gpurun_out/w$c.err
done
python bench.py --workload ipa --window-bits 20 --steps 2 --warmup 2 --no-cpu-baseline 2>gpurun_out/w20i.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  ipa c=20 value=%.4g' % d['value'], 'e2e=%.4g' % d['e2e']['value'], 'frac=%.3f' % r['frac'])" || tail -3 gpurun_out/w20i.err
