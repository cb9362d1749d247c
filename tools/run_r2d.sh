set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_tests.log 2>&1; tail -3 gpurun_out/r02d_tests.log
python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 300 gpurun_out/r02d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['checked'])
for k,v in d['also'].items():
    print(k, v.get('value'), v.get('ms_per_step'), 'e2e', v.get('e2e',{}).get('value'), 'frac', v.get('roofline',{}).get('frac'), v.get('checked'))
PY
