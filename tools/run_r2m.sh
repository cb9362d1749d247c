python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for l in 16 17 20; do echo -n "msm log2n=$l: "; python bench.py --workload msm --log2n $l --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], d['checked']['ok'])"; done
for lpj in 0 8 16 32; do echo -n "ipa LPJ=$lpj: "; VKZG_FB_LPJ=$lpj python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value=%.6g' % d['value'], 'ms=%.3f' % d['ms_per_step'], d['checked']['ok'])"; done
python tools/latency.py 2>/dev/null | tail -12
