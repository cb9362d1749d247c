#!/bin/bash
python -m pytest tests/test_gpu_tree.py tests/test_gpu_commit.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
for l in 4 2 1; do
  echo "TREE_LPJ=$l"
  VKZG_TREE_LPJ=$l python bench.py --workload tree --batch 262144 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  tree value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'e2e=%.4g' % d['e2e']['value'])"
done
