set -x
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_default_2gpu.json 2> gpurun_out/r02_bench_default_2gpu.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_2gpu.json 2>/dev/null; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d['checked'])
for k,v in d['also'].items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('scaling'), 'e2e', v.get('e2e',{}).get('value'), v.get('checked'))
PY
tail -c 400 gpurun_out/r02_bench_default_2gpu.err
cat gpurun_out/r02_bench_reference_2gpu.json | head -c 400
