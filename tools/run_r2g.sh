python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for l in 16 17 18 20; do echo -n "log2n=$l: "; python bench.py --workload msm --log2n $l --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'e2e=%.4g' % d['e2e']['value'], 'frac=%.3f' % r['frac'], 'share=%.2f' % r['kernel_share_of_step'], d['checked']['ok'])"; done
