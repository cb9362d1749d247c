#!/bin/bash
# every workload once (1 GPU); JSON lines go to gpurun_out/bench_all_r1.jsonl
out=gpurun_out/bench_all_r1.jsonl; : > $out
for w in ipa commit kzg multiproof tree; do python bench.py --workload $w --steps 5 --warmup 3 2>gpurun_out/ba_$w.err >> $out || tail -3 gpurun_out/ba_$w.err; done
for l in 16 18 20; do python bench.py --workload msm --log2n $l --steps 5 --warmup 3 2>gpurun_out/ba_msm$l.err >> $out || tail -3 gpurun_out/ba_msm$l.err; done
python - <<'PY'
import json
for l in open('gpurun_out/bench_all_r1.jsonl'):
    d=json.loads(l); r=d['roofline']; c=d.get('cpu_baseline') or {}
    print('%-34s value=%-10.4g e2e=%-10.4g ms/step=%-8.3f frac=%.3f cpu=%s launches=%s %s' % (d['metric'], d['value'], d['e2e']['value'], d['ms_per_step'], r['frac'] or 0, c.get('value'), d['gpu_launches'], d['config'].get('workload','')[:40]))
PY
