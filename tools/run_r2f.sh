python -m pytest tests/test_gpu_commit.py tests/test_gpu_ipa.py tests/test_gpu_kzg.py -m gpu -x -q 2>&1 | tail -2
run() { python bench.py --workload msm --log2n $1 --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'kernel_ms=%.3f' % (r['kernel_ms_total']/r['kernel_launches_timed']), d['checked']['ok'])"; }
for l in 17 18 19 20; do for c in 13 14 15 16 17; do for bp in 8192 65536; do
  echo -n "log2n=$l c=$c bitpar_max=$bp: "; VKZG_MSM_C=$c VKZG_MSM_BITPAR_MAX=$bp run $l
done; done; done
