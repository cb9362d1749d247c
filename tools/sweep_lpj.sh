#!/bin/bash
for l in 32 16 8; do
  echo "LPJ=$l"
  for w in commit ipa; do
  VKZG_FB_LPJ=$l python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  $w value=%.4g' % d['value'], 'frac=%.3f' % r['frac'])"
  done
done
