set -x
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02_bench_default_4gpu.json 2> gpurun_out/r02_bench_default_4gpu.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default_4gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d['checked'])
for k,v in d['also'].items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('scaling'), 'e2e', v.get('e2e',{}).get('value'), v.get('checked'))
PY
tail -c 300 gpurun_out/r02_bench_default_4gpu.err
