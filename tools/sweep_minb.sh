#!/bin/bash
for m in 3 4 5 6; do
  echo "MINB=$m"
  VKZG_FB_MINB=$m python bench.py --workload commit --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  commit value=%.4g' % d['value'], 'frac=%.3f' % r['frac'])"
  VKZG_FB_MINB=$m python bench.py --workload ipa --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  ipa value=%.4g' % d['value'], 'frac=%.3f' % r['frac'])"
done
