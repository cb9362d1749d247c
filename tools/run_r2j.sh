python -m pytest tests/test_gpu_tree.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
for th in 1 4 8 16; do echo "threads=$th"; VKZG_TREE_THREADS=$th VKZG_TREE_TIMING=1 python bench.py --workload tree --steps 5 --warmup 3 --no-cpu-baseline --no-also 2> gpurun_out/tree_$th.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'e2e=%.4g' % d['e2e']['value'], d['checked'])"; grep "vkzg_tree" gpurun_out/tree_$th.err | tail -2; done
