#!/bin/bash
for f in 6 3 2; do
  echo "FILL_X2=$f"
  VKZG_FB_FILL_X2=$f python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ipa value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'])"
  VKZG_FB_FILL_X2=$f python bench.py --workload commit --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  commit value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'])"
done
